#!/usr/bin/env python
"""Op-level timing of the memory-bound kernels (BASELINE config 1 and the model's other modes).
CUDA events on the launching stream, L2 flushed between timed launches. Prints one JSON per op."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import op  # noqa: E402
from vspbfr_b200.op.upfirdn2d import upfirdn2d_bias_act, upfirdn2d_raw  # noqa: E402
from vspbfr_b200.op.fused_act import bias_act_raw, bias_act_bwd_raw  # noqa: E402

PEAKS = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAKS = json.load(open(p))
HBM = PEAKS.get("hbm_gbs", 6650.0)


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = (k1[None] * k1[:, None] / 64).to(dev)
    rows = []

    def run(name, fn, bytes_):
        med, mn = timeit(fn, flush=flush)
        rows.append({"op": name, "us_median": med * 1e6, "us_min": mn * 1e6, "alg_MB": bytes_ / 1e6,
                     "GBps": bytes_ / med / 1e9, "frac_hbm": bytes_ / med / 1e9 / HBM})
        print(json.dumps(rows[-1]), flush=True)

    x = torch.randn(4, 512, 64, 64, device=dev)
    y = torch.randn(4, 512, 128, 128, device=dev)
    b = torch.randn(512, device=dev)
    up2 = ((2, 2), (1, 1), (2, 1, 2, 1))
    run("upfirdn2d up2 [4,512,64,64]", lambda: upfirdn2d_raw(x, k * 4, *up2), x.numel() * 4 + y.numel() * 4)
    run("upfirdn2d down2 (bwd of up2) [4,512,128,128]", lambda: upfirdn2d_raw(y, k * 4, (1, 1), (2, 2), (1, 1, 1, 1)),
        x.numel() * 4 + y.numel() * 4)
    run("bias_act fwd [4,512,128,128]", lambda: bias_act_raw(y, b, None, 3, 0, 0.2, 2 ** 0.5), 2 * y.numel() * 4)
    run("bias_act fwd [4,512,64,64]", lambda: bias_act_raw(x, b, None, 3, 0, 0.2, 2 ** 0.5), 2 * x.numel() * 4)
    run("bias_act bwd+dbias [4,512,128,128]", lambda: bias_act_bwd_raw(y, y, True, 0.2, 2 ** 0.5), 3 * y.numel() * 4)
    run("fused upfirdn2d+bias+lrelu [4,512,64,64]",
        lambda: upfirdn2d_raw(x, k * 4, *up2, bias=b, act=3, alpha=0.2, scale=2 ** 0.5), x.numel() * 4 + y.numel() * 4)
    xb = torch.randn(4, 512, 65, 65, device=dev)
    run("blur pad(1,1) [4,512,65,65]", lambda: upfirdn2d_raw(xb, k * 4, (1, 1), (1, 1), (1, 1, 1, 1)),
        xb.numel() * 4 + 4 * 512 * 64 * 64 * 4)
    xl = torch.randn(4, 32, 1025, 1025, device=dev)
    run("blur pad(1,1) [4,32,1025,1025]", lambda: upfirdn2d_raw(xl, k * 4, (1, 1), (1, 1), (1, 1, 1, 1)),
        xl.numel() * 4 + 4 * 32 * 1024 * 1024 * 4)
    xd = torch.randn(4, 64, 512, 512, device=dev)
    run("blur pad(2,2) [4,64,512,512]", lambda: upfirdn2d_raw(xd, k, (1, 1), (1, 1), (2, 2, 2, 2)),
        xd.numel() * 4 + 4 * 64 * 513 * 513 * 4)
    # torch copy as the in-run HBM yardstick
    src = torch.randn(64 * 1024 * 1024, device=dev)
    dst = torch.empty_like(src)
    run("torch copy 256MB (yardstick)", lambda: dst.copy_(src), 2 * src.numel() * 4)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
