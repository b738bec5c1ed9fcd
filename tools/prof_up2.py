#!/usr/bin/env python
"""Fused up-convolution with / without the two skip residuals (for ncu and quick timing): prof_up2.py b cin cout h"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc
b, cin, cout, h = [int(v) for v in sys.argv[1:5]]
x = mc.nchw_to_nhwc_bf16(torch.randn(b, cin, h, h, device="cuda"))
w = torch.randn(cout, cin, 3, 3, device="cuda")
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
blur = torch.outer(k1, k1) / 16
s = torch.randn(b, cin, device="cuda") * 0.3 + 1
w3 = mc.compose_up2_weights(w, blur)
wq, _ = mc.pack_weights(w3, s, wscale=1 / math.sqrt(cin * 9))
d = torch.rand(b, cout, device="cuda") + 0.5
noise = torch.randn(b, 1, 2 * h, 2 * h, device="cuda")
bias = torch.randn(cout, device="cuda")
r1 = torch.randn(b, 2 * h, 2 * h, cout, device="cuda").to(torch.bfloat16)
r2 = torch.randn(b, 2 * h, 2 * h, cout, device="cuda").to(torch.bfloat16)
nwd = torch.full((1,), 0.05, device="cuda")
for res in (False, True):
    epi = mc.make_epilogue(row_scale=d, noise=noise, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2), noise_weight_dev=nwd,
                           residual=r1 if res else None, residual2=r2 if res else None)
    for _ in range(2):
        out = mc.conv_up2_fused(x, wq, cout, epi=epi)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(3):
        out = mc.conv_up2_fused(x, wq, cout, epi=epi)
    en.record()
    torch.cuda.synchronize()
    print(f"up2 b{b} {cin}->{cout} {h}x{h} residuals={res}: {st.elapsed_time(en) / 3 * 1e3:.1f} us")
