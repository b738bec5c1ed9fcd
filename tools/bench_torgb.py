#!/usr/bin/env python
"""ToRGB kernel in isolation: us and GB/s (algorithmic: feature map read + RGB out [+ skip read]) per model shape.
bench_torgb.py [batch]   (VSP_NO_TORGB_LANES=1 selects the pixel-per-thread kernels)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import _lib
from vspbfr_b200._lib import ptr, stream_ptr
from vspbfr_b200.op.upfirdn2d import upfirdn2d_raw

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
lib = _lib.load()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / n


for c, h in ((64, 512), (128, 256), (256, 128), (512, 64), (32, 1024)):
    x = torch.randn(b, h, h, c, device=dev).to(torch.bfloat16)
    w = torch.randn(3, c, device=dev)
    s = torch.randn(b, c, device=dev)
    bias = torch.randn(3, device=dev)
    skip = torch.randn(b, 3, h, h, device=dev)
    low = torch.randn(b, 3, h // 2, h // 2, device=dev)
    out = torch.empty(b, 3, h, h, device=dev)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device=dev)
    k = torch.outer(k1, k1) / 16
    for name, sk in (("no-skip", None), ("skip", skip)):
        t = timeit(lambda: lib.vsp_torgb_nhwc_bf16(ptr(x), ptr(w), ptr(s), ptr(bias), ptr(sk), ptr(out), b, h * h, c, 0.125,
                                                    stream_ptr()))
        nbytes = x.numel() * 2 + out.numel() * 4 * (2 if sk is not None else 1)
        print(f"torgb c{c:<4d} {h}x{h} b{b} {name:8s} {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s")
    if c != 32:
        t = timeit(lambda: upfirdn2d_raw(low, k, (2, 2), (1, 1), (2, 1, 2, 1)))
        print(f"  skip upsample [{b},3,{h // 2},{h // 2}] -> {h}: {t * 1e6:8.1f} us  {(low.numel() + out.numel()) * 4 / t / 1e9:7.0f} GB/s")
    else:
        taps = (__import__("ctypes").c_float * 9)(*[1 / 9.0] * 9)
        o2 = torch.empty(b, 3, h // 2, h // 2, device=dev)
        for name, tp in (("3x3 skip taps in-kernel", taps), ("pre-pooled skip", None)):
            t = timeit(lambda: lib.vsp_torgb_pool2_nhwc_bf16(ptr(x), ptr(w), ptr(s), ptr(bias), ptr(low), tp, ptr(o2), b, h // 2,
                                                             h // 2, c, 0.125, stream_ptr()))
            print(f"torgb_pool2 c{c} {h}x{h} b{b} {name:24s} {t * 1e6:8.1f} us  {(x.numel() * 2 + o2.numel() * 8) / t / 1e9:7.0f} GB/s")
        k3 = torch.full((3, 3), 1 / 9.0, device=dev)
        t = timeit(lambda: upfirdn2d_raw(low, k3, (1, 1), (1, 1), (1, 1, 1, 1)))
        print(f"  skip 3x3 pass [{b},3,{h // 2},{h // 2}]: {t * 1e6:8.1f} us")
    del x, skip, out
