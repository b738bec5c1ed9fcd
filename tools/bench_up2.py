#!/usr/bin/env python
"""Up-sampling modulated convolutions of the hot path at micro-batch 32: dense composite form (vsp_conv2d_up2_fused_bf16,
4x the algorithmic FLOPs) vs the half-composed form (vsp_conv2d_up2h_bf16, 2x), with and without the two skip residuals.
One JSON row per measurement; algorithmic TFLOP/s = 2*B*H*W*Cout*Cin*9 / time, GB/s = (x + out [+ residuals]) / time."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=5, inner=4):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner):
            fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / inner)
    return best * 1e-3


def main():
    b = int(os.environ.get("B", "32"))
    rows = []
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k4 = (torch.outer(k1, k1) / 64 * 4).to(DEV)
    fx = (k1 / 4).tolist()
    for cin, cout, h in ((64, 32, 512), (128, 64, 256)):
        w = h
        g = torch.Generator(device="cpu").manual_seed(1)
        x = torch.randn(b, h, w, cin, device=DEV).to(torch.bfloat16)
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
        s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV)
        noise = torch.randn(b, 1, 2 * h, 2 * w, device=DEV)
        bias = torch.randn(cout, device=DEV)
        rs = torch.rand(b, cout, device=DEV) + 0.5
        wq4, _ = mc.pack_weights(mc.compose_up2_weights(wt, k4), s)
        wq2, _ = mc.pack_weights(mc.compose_up2h_weights(wt, fx), s)
        out = torch.empty(b, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=DEV)
        out2 = torch.empty_like(out)
        for with_res in (False, True):
            kw = {}
            if with_res:
                kw = dict(residual=torch.randn(b, 2 * h, 2 * w, cout, device=DEV).to(torch.bfloat16),
                          residual2=torch.randn(b, 2 * h, 2 * w, cout, device=DEV).to(torch.bfloat16))
            epi = mc.make_epilogue(row_scale=rs, noise=noise, noise_weight=0.05, bias=bias, act=3, alpha=0.2,
                                   scale=math.sqrt(2), **kw)
            flops = 2.0 * b * h * w * cout * cin * 9
            nbytes = 2.0 * b * (h * w * cin + 4 * h * w * cout * (3 if with_res else 1)) + 4.0 * b * 4 * h * w
            t4 = timeit(lambda: mc.conv_up2_fused(x, wq4, cout, epi=epi, out=out))
            t2 = timeit(lambda: mc.conv_up2h(x, wq2, cout, fx[::-1], epi=epi, out=out2))
            err = float((out.float() - out2.float()).abs().max() / out.float().abs().max())
            for name, t in (("dense 4x (up2_fused)", t4), ("half-composed 2x (up2h)", t2)):
                rows.append({"layer": f"b{b} {cin}->{cout} {h}x{w}->{2 * h}x{2 * w}" + (" +2 residuals" if with_res else ""),
                             "impl": name, "us": t * 1e6, "alg_TFLOPs": flops / t / 1e12, "GBps": nbytes / t / 1e9,
                             "max_rel_diff_between_impls": err})
                print(json.dumps(rows[-1]), flush=True)
            del kw, epi
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_up2.json", "w"), indent=1)


if __name__ == "__main__":
    main()
