#!/usr/bin/env python
"""Timing of the tcgen05 convolution on BASELINE config 2 (B=8, 512->512, 3x3, 64x64) and the
other representative layer shapes; CUDA events, L2 flushed between launches."""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
TF = PEAKS.get("bf16_tflops", 1590.0)


def timeit(fn, iters=10, warmup=3, flush=None):
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
    return float(np.median(ts))


def main():
    dev = "cuda"
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []

    def rec(name, t, flops):
        rows.append({"op": name, "us": t * 1e6, "TFLOPs": flops / t / 1e12, "frac_bf16_peak": flops / t / 1e12 / TF})
        print(json.dumps(rows[-1]), flush=True)

    shapes = [  # b, cin, cout, h, k, dil
        (8, 512, 512, 64, 3, 1),
        (8, 256, 256, 128, 3, 1),
        (8, 128, 128, 256, 3, 1),
        (4, 64, 64, 512, 3, 1),
        (4, 64, 16, 512, 3, 2),
        (8, 512, 128, 64, 3, 4),
        (4, 32, 32, 1024, 3, 1),
        (8, 512, 512, 16, 3, 1),
        (8, 64, 3, 512, 1, 1),
    ]
    for b, cin, cout, h, k, dil in shapes:
        torch.manual_seed(0)
        x = torch.randn(b, cin, h, h, device=dev)
        w = torch.randn(cout, cin, k, k, device=dev)
        s = torch.randn(b, cin, device=dev) * 0.3 + 1
        scale = 1 / math.sqrt(cin * k * k)
        xq = mc.nchw_to_nhwc_bf16(x)
        wq, d = mc.pack_weights(w, s, wscale=scale, want_demod=True)
        epi = mc.make_epilogue(row_scale=d)
        pad = (k - 1) * dil // 2
        flops = 2.0 * b * cout * cin * k * k * h * h
        tag = f"B{b} {cin}->{cout} {h}x{h} k{k} d{dil}"
        out32 = torch.empty(b, cout, h, h, device=dev)
        t = timeit(lambda: mc.conv_fprop(xq, wq, cout, k, k, 1, pad, dil, epi=epi, out=out32), flush=flush)
        rec(f"fprop kernel (NCHW f32 out) {tag}", t, flops)
        outq = torch.empty(b, h, h, mc._round_up(cout, 8), dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: mc.conv_fprop(xq, wq, cout, k, k, 1, pad, dil, epi=epi, out=outq, out_nhwc=True), flush=flush)
        rec(f"fprop kernel (NHWC bf16 out) {tag}", t, flops)
        if h <= 256:
            wq1, _ = mc.pack_weights(w)
            t = timeit(lambda: mc.conv_fprop(xq, wq1, cout, k, k, 1, pad, dil, out=outq, out_nhwc=True), flush=flush)
            rec(f"fprop kernel shared weights {tag}", t, flops)
        del out32, outq
    # config 2 end-to-end pieces
    b, c, h = 8, 512, 64
    x = torch.randn(b, c, h, h, device=dev, requires_grad=True)
    w = torch.randn(1, c, c, 3, 3, device=dev, requires_grad=True)
    s = (torch.randn(b, c, device=dev) * 0.3 + 1).requires_grad_(True)
    flops = 2.0 * b * c * c * 9 * h * h
    t = timeit(lambda: mc.nchw_to_nhwc_bf16(x), flush=flush); rec("cfg2 nchw->nhwc bf16", t, 0.0)
    t = timeit(lambda: mc.pack_weights(w.reshape(c, c, 3, 3), s, wscale=1 / 67.88, want_demod=True), flush=flush); rec("cfg2 weight prologue", t, 0.0)
    with torch.no_grad():
        t = timeit(lambda: mc.modulated_conv2d(x, w, s, True, "same", 1), flush=flush); rec("cfg2 modulated_conv2d fwd (op-level, fp32 NCHW in/out)", t, flops)
    dy = torch.randn(b, c, h, h, device=dev)

    def fwdbwd():
        y = mc.modulated_conv2d(x, w, s, True, "same", 1)
        torch.autograd.grad(y, [x, w, s], dy)

    t = timeit(fwdbwd, flush=flush); rec("cfg2 modulated_conv2d fwd+bwd (op-level)", t, 3 * flops)
    dzq = mc.nchw_to_nhwc_bf16(dy)
    xq = mc.nchw_to_nhwc_bf16(x.detach())
    t = timeit(lambda: mc.conv_wgrad(dzq, xq, b, 3, 3, 1, 1, 1), flush=flush); rec("cfg2 wgrad kernel (per-sample)", t, flops)
    t = timeit(lambda: mc.conv_wgrad(dzq, xq, 1, 3, 3, 1, 1, 1), flush=flush); rec("cfg2 wgrad kernel (shared)", t, flops)
    # incumbents: cuDNN grouped fp32 (what the reference runs), TF32 and bf16 channels_last
    wm = torch.randn(b * c, c, 3, 3, device=dev)
    xin = x.detach().reshape(1, b * c, h, h)
    torch.backends.cudnn.allow_tf32 = False
    t = timeit(lambda: F.conv2d(xin, wm, padding=1, groups=b), flush=flush); rec("cfg2 cuDNN grouped fp32 (reference path)", t, flops)
    torch.backends.cudnn.allow_tf32 = True
    t = timeit(lambda: F.conv2d(xin, wm, padding=1, groups=b), flush=flush); rec("cfg2 cuDNN grouped tf32", t, flops)
    xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wb = torch.randn(c, c, 3, 3, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t = timeit(lambda: F.conv2d(xb, wb, padding=1), flush=flush); rec("cfg2 cuDNN bf16 channels_last shared weights", t, flops)
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    bm = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: a @ bm); rec("cuBLAS bf16 8192^3 (yardstick)", t, 2.0 * 8192 ** 3)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_conv.json", "w"), indent=1)


if __name__ == "__main__":
    main()
