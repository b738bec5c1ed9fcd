#!/usr/bin/env python
"""Op-level timing of the memory-bound kernels (BASELINE config 1 and the model's other modes).
CUDA events on the launching stream; every launch of a timed series reads operands no earlier launch touched (rotating
buffer sets > 1.2 GB per round), plus lone launches after an L2 flush. Prints one JSON per op."""
import json
import math
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


def timeit(fns, rounds=5):
    """``fns``: the same op on DISTINCT buffer sets whose total footprint exceeds L2 (126 MB) several times, so no launch finds
    its operands cached by an earlier one.  Two figures: (a) all launches back to back between one pair of CUDA events,
    best of ``rounds`` (the event clock is ~2 us coarse and a lone 30-100 us launch also pays its ramp-up); (b) the median
    of lone launches, each between its own events after a 256 MB L2 flush."""
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(rounds):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for fn in fns:
            fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3 / len(fns))
    lone = []
    for fn in fns:
        FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        lone.append(s.elapsed_time(e) * 1e-3)
    return best, float(np.median(lone))


FLUSH = None


def main():
    global FLUSH
    dev = "cuda"
    FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = (k1[None] * k1[:, None] / 64).to(dev)
    k4 = (k * 4).contiguous()
    rows = []

    def run(name, make, bytes_, sets=None):
        """``make()`` -> a closure launching the op on its OWN freshly allocated operands (outputs are allocated by the op)."""
        n = sets or max(4, int(math.ceil(1.2e9 / bytes_)))           # >= 1.2 GB of distinct operands per round
        fns = [make() for _ in range(n)]
        b2b, lone = timeit(fns)
        rows.append({"op": name, "us_back_to_back": b2b * 1e6, "us_lone_median": lone * 1e6, "alg_MB": bytes_ / 1e6,
                     "GBps": bytes_ / b2b / 1e9, "frac_hbm": bytes_ / b2b / 1e9 / HBM,
                     "frac_hbm_lone": bytes_ / lone / 1e9 / HBM, "buffer_sets": n})
        print(json.dumps(rows[-1]), flush=True)
        del fns
        torch.cuda.empty_cache()

    R = lambda *shape: torch.randn(*shape, device=dev)
    b = torch.randn(512, device=dev)
    up2 = ((2, 2), (1, 1), (2, 1, 2, 1))
    nx, ny = 4 * 512 * 64 * 64, 4 * 512 * 128 * 128

    def mk_up2():
        x = R(4, 512, 64, 64)
        return lambda: upfirdn2d_raw(x, k4, *up2)
    run("upfirdn2d up2 [4,512,64,64]", mk_up2, (nx + ny) * 4)

    def mk_down2():
        y = R(4, 512, 128, 128)
        return lambda: upfirdn2d_raw(y, k4, (1, 1), (2, 2), (1, 1, 1, 1))
    run("upfirdn2d down2 (bwd of up2) [4,512,128,128]", mk_down2, (nx + ny) * 4)

    def mk_ba(shape):
        def mk():
            y = R(*shape)
            return lambda: bias_act_raw(y, b, None, 3, 0, 0.2, 2 ** 0.5)
        return mk
    run("bias_act fwd [4,512,128,128]", mk_ba((4, 512, 128, 128)), 2 * ny * 4)
    run("bias_act fwd [4,512,64,64]", mk_ba((4, 512, 64, 64)), 2 * nx * 4)

    def mk_bab():
        dy, ref = R(4, 512, 128, 128), R(4, 512, 128, 128)          # distinct gradient and reference tensors
        return lambda: bias_act_bwd_raw(dy, ref, True, 0.2, 2 ** 0.5)
    run("bias_act bwd+dbias [4,512,128,128]", mk_bab, 3 * ny * 4)

    def mk_fused():
        x = R(4, 512, 64, 64)
        return lambda: upfirdn2d_raw(x, k4, *up2, bias=b, act=3, alpha=0.2, scale=2 ** 0.5)
    run("fused upfirdn2d+bias+lrelu [4,512,64,64]", mk_fused, (nx + ny) * 4)

    def mk_blur(shape, gain, pad):
        def mk():
            x = R(*shape)
            kk = (k * gain).contiguous()
            return lambda: upfirdn2d_raw(x, kk, (1, 1), (1, 1), pad)
        return mk
    run("blur pad(1,1) [4,512,65,65]", mk_blur((4, 512, 65, 65), 4, (1, 1, 1, 1)), (4 * 512 * 65 * 65 + nx) * 4)
    run("blur pad(1,1) [4,32,1025,1025]", mk_blur((4, 32, 1025, 1025), 4, (1, 1, 1, 1)),
        (4 * 32 * 1025 * 1025 + 4 * 32 * 1024 * 1024) * 4)
    run("blur pad(2,2) [4,64,512,512]", mk_blur((4, 64, 512, 512), 1, (2, 2, 2, 2)),
        (4 * 64 * 512 * 512 + 4 * 64 * 513 * 513) * 4)

    def mk_copy():
        src = R(64 * 1024 * 1024)
        dst = torch.empty_like(src)
        return lambda: dst.copy_(src)
    run("torch copy 256MB (yardstick)", mk_copy, 2 * 64 * 1024 * 1024 * 4, sets=4)

    # same-size yardsticks: what a plain device copy moving the SAME number of bytes reaches (launch ramp + tail of a 15-40 us
    # kernel are a fixed cost that the 256 MB copy amortises)
    def mk_copy_n(nfloat):
        def mk():
            src = R(nfloat)
            dst = torch.empty_like(src)
            return lambda: dst.copy_(src)
        return mk
    run("torch copy, 168 MB of traffic (= config 1 / down2 / fused)", mk_copy_n((nx + ny) // 2), (nx + ny) * 4)
    run("torch copy, 67 MB of traffic (= bias_act [4,512,64,64] / blur [4,512,65,65])", mk_copy_n(nx), 2 * nx * 4)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
