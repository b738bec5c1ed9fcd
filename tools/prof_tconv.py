#!/usr/bin/env python
"""Stride-2 transposed convolution (the up-convolution's 1x-FLOPs form) a few times, for ncu and quick timing:
prof_tconv.py b cin cout h [groups: 1 = shared weights, 0 = per-sample]"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc
b, cin, cout, h = [int(v) for v in sys.argv[1:5]]
shared = len(sys.argv) > 5 and sys.argv[5] == "1"
x = torch.randn(b, cin, h, h, device="cuda")
w = torch.randn(cout, cin, 3, 3, device="cuda")
s = torch.randn(b, cin, device="cuda") * 0.3 + 1
xq = mc.nchw_to_nhwc_bf16(x)
wq, d = mc.pack_weights(w, None if shared else s, wscale=1 / math.sqrt(cin * 9), want_demod=not shared)
epi = mc.make_epilogue(row_scale=d) if d is not None else None
for _ in range(3):
    out = mc.conv_transpose_s2(xq, wq, cout, 3, 3, epi=epi, out_nhwc=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    out = mc.conv_transpose_s2(xq, wq, cout, 3, 3, epi=epi, out_nhwc=True)
ev[1].record()
torch.cuda.synchronize()
us = ev[0].elapsed_time(ev[1]) * 100
print(f"conv_transpose_s2 b{b} {cin}->{cout} {h}x{h}: {us:.1f} us  {2.0 * b * h * h * cin * cout * 9 / us / 1e6:.0f} TF/s", out.shape)
