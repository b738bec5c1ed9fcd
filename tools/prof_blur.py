#!/usr/bin/env python
"""A few launches of the fp32 drop-in blur modes for ncu (tools/prof_blur.py [case]); no timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op.upfirdn2d import upfirdn2d_raw

case = sys.argv[1] if len(sys.argv) > 1 else "1025"
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
k = (k1[None] * k1[:, None] / 64 * 4).cuda().contiguous()
shape, pad = {"1025": ((4, 32, 1025, 1025), (1, 1, 1, 1)), "65": ((4, 512, 65, 65), (1, 1, 1, 1)),
              "512": ((4, 64, 512, 512), (2, 2, 2, 2)), "down2": ((4, 512, 128, 128), (1, 1, 1, 1)),
              "up2": ((4, 512, 64, 64), (2, 1, 2, 1)), "up2f": ((4, 512, 64, 64), (2, 1, 2, 1))}[case]
xs = [torch.randn(*shape, device="cuda") for _ in range(3)]
for x in xs:
    if case == "down2":
        upfirdn2d_raw(x, k, (1, 1), (2, 2), pad)
    elif case == "up2":
        upfirdn2d_raw(x, k, (2, 2), (1, 1), pad)
    elif case == "up2f":
        upfirdn2d_raw(x, k, (2, 2), (1, 1), pad, bias=torch.randn(512, device="cuda"), act=3, alpha=0.2, scale=2 ** 0.5)
    else:
        upfirdn2d_raw(x, k, (1, 1), (1, 1), pad)
torch.cuda.synchronize()
