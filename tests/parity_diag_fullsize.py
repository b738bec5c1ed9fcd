"""Full-size parity diagnostics: fused pipeline vs CPU oracle, error statistics (env flags select kernel paths)."""
import math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from vspbfr_b200 import fastpath as fp
from vspbfr_b200.restorenet import Restoration_net
from vspbfr_b200.stylegan2 import Generator
torch.manual_seed(11)
net = Restoration_net(512, 512, 8, channel_multiplier=2).eval()
dec = Generator(1024, 512, 8, channel_multiplier=2).eval()
g = torch.Generator().manual_seed(12)
low = torch.rand(2, 3, 512, 512, generator=g) * 2 - 1
codes = torch.randn(2, 18, 512, generator=g)
z = torch.randn(2, 512, generator=g)
cache = "/tmp/fullsize_ref.pt"
if os.path.exists(cache):
    want, want_img = torch.load(cache)
else:
    with torch.no_grad():
        want, want_img = oracle.restore_faces_ref(net.state_dict(), dec.state_dict(), low, codes, z, 512, 1024, 8)
    torch.save((want, want_img), cache)
net, dec = net.cuda(), dec.cuda()
got, got_img = fp.restore_faces(net, dec, low.cuda(), codes.cuda(), [z.cuda()])
got = got.cpu()
err = (got - want).abs()
peak = float(want.max() - want.min())
mse = float(((got - want) ** 2).mean())
print(f"flags={ {k: v for k, v in os.environ.items() if k.startswith('VSP_')} }")
print(f"range {peak:.1f}  max-abs {float(err.max()):.3f} ({100*float(err.max())/peak:.2f}% of range)  psnr {10*math.log10(peak*peak/mse):.1f} dB  "
      f"mean-abs {float(err.mean()):.4f}  p99.9 {float(err.flatten().kthvalue(int(0.999*err.numel())).values):.3f}")
idx = np.unravel_index(int(err.argmax()), err.shape)
print("argmax", idx, "got", float(got[idx]), "want", float(want[idx]))
# structure: mean error by column mod 128 / row mod 4 (tile artefacts would show up here)
e2 = err.mean(dim=(0, 1))
print("col%128 max/mean of column-mean error:", float(e2.mean(0).view(4, 128).mean(0).max()), float(e2.mean()))
print("row%8 means:", [round(float(v), 4) for v in e2.mean(1).view(64, 8).mean(0)])
