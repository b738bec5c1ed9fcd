#!/usr/bin/env python
"""Grouped style linears (vsp_grouped_linear_f32) at the shapes of a 32-face micro-batch: the restorer decoder's bank (17
modulations of a [B, 2048] style, 8960 rows), a [B, 512] bank and one 512 x 512 style-MLP layer.  us per launch, TB/s on
the weights.  VSP_LINEAR_GROUPS=1|2|4 forces the 8-row groups per block (default: by problem size)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import fastpath as fp, layers as L

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cases = {"restorer bank 2048": (2048, [512] * 11 + [256, 256, 128, 128, 64, 64, 32, 32]),
         "decoder bank 512": (512, [512] * 15 + [256, 256, 256, 128, 128, 128, 64, 64, 64, 32, 32, 32]),
         "style MLP layer": (512, [512])}
for name, (d, outs) in cases.items():
    lins = [L.EqualLinear(d, o, bias_init=1).to(dev) for o in outs]
    bank = fp.ModulationBank([(m, i % 4) for i, m in enumerate(lins)])
    styles = torch.randn(b, 4, d, device=dev)
    for _ in range(3):
        bank(styles)
    # the call is host-bound in eager mode (descriptor key check, torch.empty, ctypes): time a captured replay instead
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        bank(styles)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            bank(styles)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(5):
        ev[0].record()
        graph.replay()
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]) / 10 * 1e3)
    nbytes = sum(m.weight.numel() for m in lins) * 4
    print(f"{name:20s} b{b} rows {sum(outs):5d} groups={os.environ.get('VSP_LINEAR_GROUPS', 'auto')}: {best:7.1f} us  "
          f"{nbytes / best / 1e6:5.2f} TB/s")
