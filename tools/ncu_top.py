#!/usr/bin/env python
"""Top stall sites from `ncu --page source --csv` output: ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[ix['# Samples']] or 0)
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:7d} {100*s/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top}")
