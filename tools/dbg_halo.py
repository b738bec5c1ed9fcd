import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from vspbfr_b200.op import modconv as mc
torch.manual_seed(0)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randn(1, 64, 6, 128, device="cuda")
xq = mc.nchw_to_nhwc_bf16(x)
xr = xq.permute(0, 3, 1, 2).float()
for tap in range(9):
    w = torch.zeros(64, 64, 3, 3, device="cuda")
    w[:, :, tap // 3, tap % 3] = torch.randn(64, 64, device="cuda") / 8
    wq, _ = mc.pack_weights(w)
    out = mc.conv_fprop(xq, wq, 64, 3, 3, 1, d, d)
    want = F.conv2d(xr, w.to(torch.bfloat16).float(), None, 1, d, d)
    err = (out - want).abs()
    print("tap", tap, "kh,kw", tap // 3, tap % 3, "maxerr %.4f" % float(err.max()), "bad cols", sorted(set(err.amax(dim=(0, 1, 2)).gt(0.05).nonzero().flatten().tolist()))[:12])
