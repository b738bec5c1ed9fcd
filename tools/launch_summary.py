#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = {}
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("vsp::<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/tot:5.1f}%  x{n:4d}  {name[:110]}")
