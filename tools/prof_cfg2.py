#!/usr/bin/env python
"""BASELINE configs[1] through the module API: ModulatedConv2d(512, 512, 3, style_dim=512) forward and forward+backward on
[8,512,64,64] fp32 NCHW — eager and CUDA-graph timings plus the per-kernel table of one iteration (torch profiler)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.layers import ModulatedConv2d  # noqa: E402


def main():
    dev = "cuda"
    torch.manual_seed(0)
    b, c, h = 8, 512, 64
    m = ModulatedConv2d(c, c, 3, 512).to(dev)
    x = torch.randn(b, c, h, h, device=dev, requires_grad=True)
    style = torch.randn(b, 512, device=dev, requires_grad=True)
    dy = torch.randn(b, c, h, h, device=dev)
    params = [x, style] + list(m.parameters())

    def fwd():
        with torch.no_grad():
            return m(x, style)

    def fwdbwd():
        y = m(x, style)
        return torch.autograd.grad(y, params, dy)

    flops = 2.0 * b * c * c * 9 * h * h
    for name, fn, mult in (("fwd", fwd, 1), ("fwd+bwd", fwdbwd, 3)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) / 20 * 1e-3
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        s.record()
        for _ in range(20):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        tg = s.elapsed_time(e) / 20 * 1e-3
        print(f"{name}: eager {t * 1e6:.1f} us ({mult * flops / t / 1e12:.0f} TFLOP/s), graph {tg * 1e6:.1f} us "
              f"({mult * flops / tg / 1e12:.0f} TFLOP/s)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fwdbwd()
        torch.cuda.synchronize()
    rows = [(ev.key, ev.device_time_total, ev.count) for ev in prof.key_averages() if ev.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"kernel time of one fwd+bwd: {tot:.1f} us")
    for k, t, n in rows[:30]:
        print(f"{t:8.1f} us x{n:<3d} {k[:110]}")


if __name__ == "__main__":
    main()
