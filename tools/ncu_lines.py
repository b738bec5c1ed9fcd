#!/usr/bin/env python
"""Stall samples aggregated per CUDA source line from `ncu --page source --print-source cuda,sass --csv`:
ncu_lines.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if not hdr or len(r) != len(hdr):
        continue
    ix_s = hdr.index("# Samples")
    line, src = r[0], r[1]
    if not line:      # SASS row under a source line: has its own samples
        continue
    try:
        s = int(r[ix_s] or 0)
    except ValueError:
        s = 0
    st = {}
    for k, v in zip(hdr, r):
        if k.startswith("stall_") and "Not Issued" not in k:
            try: st[k] = int(v or 0)
            except ValueError: pass
    a = agg.setdefault((cur, int(line)), [0, src.strip(), {}])
    a[0] += s
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, l), (s, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = sorted(((v, k) for k, v in st.items()), reverse=True)[:2]
    print(f"{s:6d} {100*s/max(tot,1):5.1f}%  {f}:{l:<4d} {src[:80]:80s} {top}")
