#!/usr/bin/env python
"""Channels-last bf16 blur with the noise + bias + lrelu + two-residual epilogue at the model's shape
([32, 129, 129, 256] -> [32, 128, 128, 256], after the 512->256 transposed conv): us and algorithmic TB/s.
bench_blur_epi.py [batch]    (VSP_BLUR_EPI_AHEAD=0..3 selects the look-ahead of the epilogue operand loads)"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import fastpath as fp
from vspbfr_b200.op import modconv as mc

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prof = len(sys.argv) > 2 and sys.argv[2] == "prof"      # three epilogue launches of the first shape only (for ncu)
dev = torch.device("cuda", 0)
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
k = (k1[None] * k1[:, None] / 64 * 4).to(dev)
for (c, h) in ((256, 129), (512, 65)):
    x = torch.randn(b, h, h, c, device=dev).to(torch.bfloat16)
    r1 = torch.randn(b, h - 1, h - 1, c, device=dev).to(torch.bfloat16)
    r2 = torch.randn(b, h - 1, h - 1, c, device=dev).to(torch.bfloat16)
    noise = torch.randn(b, 1, h - 1, h - 1, device=dev)
    bias = torch.randn(c, device=dev)
    for name, e in (("plain", None), ("epilogue", mc.make_epilogue(noise=noise, noise_weight=0.1, bias=bias, act=3, alpha=0.2,
                                                                    scale=math.sqrt(2), residual=r1, residual2=r2))):
        if prof and e is None:
            continue
        for _ in range(3):
            fp.upfirdn_nhwc(x, k, pad=(1, 1), epi=e)
        if prof:
            torch.cuda.synchronize()
            sys.exit(0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = 1e9
        for _ in range(5):
            ev[0].record()
            for _ in range(10):
                fp.upfirdn_nhwc(x, k, pad=(1, 1), epi=e)
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]) / 10 * 1e3)
        nbytes = x.numel() * 2 + r1.numel() * 2 * (1 if e is None else 3) + (0 if e is None else noise.numel() * 4)
        print(f"blur_nhwc b{b} c{c} {h}x{h} {name:9s} ahead={os.environ.get('VSP_BLUR_EPI_AHEAD', 'default')}: "
              f"{best:7.1f} us  {nbytes / best / 1e6:5.2f} TB/s")
