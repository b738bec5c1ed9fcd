import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import fastpath as fp
b, c, h = [int(v) for v in sys.argv[1:4]]
x = torch.randn(b, h + 1, h + 1, c, device="cuda").to(torch.bfloat16)
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
k = torch.outer(k1, k1) / 16
for _ in range(3):
    y = fp.upfirdn_nhwc(x, k, pad=(1, 1))
torch.cuda.synchronize()
print("done", y.shape)
