#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel name -> launches, total us, DRAM read/write MB.  launch_traffic.py file.csv [--json out.json]"""
import csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = {}
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]); name = re.sub(r"^void ", "", name).replace("vsp::<unnamed>::", "")
    v = float(r["Metric Value"].replace(",", "")); unit = r["Metric Unit"]; m = r["Metric Name"]
    a = per.setdefault((r["ID"], name), {})
    if m == "gpu__time_duration.sum":
        a["us"] = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        a["rd" if "read" in m else "wr"] = v * scale
agg = {}
for (_, name), a in per.items():
    g = agg.setdefault(name, {"launches": 0, "us": 0.0, "dram_read_mb": 0.0, "dram_write_mb": 0.0})
    g["launches"] += 1; g["us"] += a.get("us", 0); g["dram_read_mb"] += a.get("rd", 0) / 1e6; g["dram_write_mb"] += a.get("wr", 0) / 1e6
tot = sum(g["us"] for g in agg.values())
print(f"total {tot:.1f} us over {sum(g['launches'] for g in agg.values())} launches, "
      f"DRAM {sum(g['dram_read_mb'] for g in agg.values())/1e3:.2f} GB read + {sum(g['dram_write_mb'] for g in agg.values())/1e3:.2f} GB written")
for name, g in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    bw = (g["dram_read_mb"] + g["dram_write_mb"]) / max(g["us"], 1e-9) * 1e3 / 1e3
    print(f"{g['us']:10.1f} us {100*g['us']/tot:5.1f}%  x{g['launches']:4d}  rd {g['dram_read_mb']:9.1f} MB  wr {g['dram_write_mb']:9.1f} MB  {bw:6.2f} TB/s  {name[:90]}")
if "--json" in sys.argv:
    conv = {k: v for k, v in agg.items() if k.startswith("conv_")}
    out = {"total_us": tot, "kernels": agg,
           "conv_kernels": {"launches": sum(v["launches"] for v in conv.values()), "us": sum(v["us"] for v in conv.values()),
                            "dram_bytes": 1e6 * sum(v["dram_read_mb"] + v["dram_write_mb"] for v in conv.values())}}
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
