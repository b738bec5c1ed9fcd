"""Diagnostics (not a test): decoder-image error of the fused pipeline vs the fp32 CPU oracle with the half-composed
up-convolution on / off, over a few seeds.  python tests/parity_diag_up2h.py [seeds...]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from vspbfr_b200 import fastpath as fp  # noqa: E402
from vspbfr_b200.restorenet import Restoration_net  # noqa: E402
from vspbfr_b200.stylegan2 import Generator  # noqa: E402


def stats(got, want):
    peak = float(want.max() - want.min())
    d = (got - want).double()
    return float(d.abs().max()) / peak, math.sqrt(float((d * d).mean())) / peak


def main():
    seeds = [int(s) for s in sys.argv[1:]] or [11, 5, 7]
    for seed in seeds:
        torch.manual_seed(seed)
        net = Restoration_net(512, 512, 8, channel_multiplier=2).eval()
        dec = Generator(1024, 512, 8, channel_multiplier=2).eval()
        g = torch.Generator().manual_seed(seed + 1)
        low = torch.rand(2, 3, 512, 512, generator=g) * 2 - 1
        codes = torch.randn(2, 18, 512, generator=g)
        z = torch.randn(2, 512, generator=g)
        with torch.no_grad():
            want, want_img = oracle.restore_faces_ref(net.state_dict(), dec.state_dict(), low, codes, z, 512, 1024, 8)
        want_img = torch.nn.functional.adaptive_avg_pool2d(want_img, (512, 512))
        net, dec = net.cuda(), dec.cuda()
        for flag in (True, False):
            fp._UP2H = flag
            fp.clear_cache()
            got, got_img = fp.restore_faces(net, dec, low.cuda(), codes.cuda(), [z.cuda()])
            mi, ri = stats(got_img.cpu(), want_img)
            mr, rr = stats(got.cpu(), want)
            print(f"seed {seed} up2h={int(flag)}: decoder image max {mi:.3e} rms {ri:.3e} | restored max {mr:.3e} rms {rr:.3e}", flush=True)


if __name__ == "__main__":
    main()
