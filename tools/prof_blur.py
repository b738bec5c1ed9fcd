import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import fastpath as fp
b, c, h, pad = [int(v) for v in sys.argv[1:5]]
x = torch.randn(b, h, h, c, device="cuda").to(torch.bfloat16)
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
k = torch.outer(k1, k1) / 64
for _ in range(3):
    y = fp.upfirdn_nhwc(x, k, pad=(pad, pad))
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    y = fp.upfirdn_nhwc(x, k, pad=(pad, pad))
e.record(); torch.cuda.synchronize()
print(f"R={os.environ.get('VSP_BLUR_R','4')} b{b} c{c} {h} pad{pad}: {s.elapsed_time(e)*100:.1f} us", y.shape)
