#!/usr/bin/env python
"""Per-launch table of one hot-path micro-batch (CUDA events around each conv / blur / ToRGB / prologue call):
prof_layers.py [micro]   -> stdout table: name, shape, us, TFLOP/s, GB/s (algorithmic)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vspbfr_b200 import fastpath
from vspbfr_b200.op import modconv as mc

micro = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
net, dec = bench.build_models(dev)
low, codes, z = (t.to(dev) for t in bench.synth_inputs(micro, 1))

_orig_up = fastpath.upfirdn_nhwc
_orig_rgb = fastpath.to_rgb
_orig_pack = mc.pack_weights
_orig_lin = fastpath._linear


def up(x, kernel, up=1, down=1, pad=(0, 0), epi=None):
    n, h, w, c = x.shape
    return mc._prof("blur_nhwc", 0.0, lambda: _orig_up(x, kernel, up, down, pad, epi),
                    detail=f"b{n} c{c} {h}x{w} pad{pad} epi={epi is not None}", nbytes=2.0 * 2 * n * h * w * c)


def rgb(m, x, style, skip=None):
    n, h, w, c = x.shape
    return mc._prof("to_rgb(+linear+upsample)", 0.0, lambda: _orig_rgb(m, x, style, skip), detail=f"b{n} c{c} {h}x{w}",
                    nbytes=2.0 * n * h * w * c + 4.0 * 2 * 3 * n * h * w)


def pack(weight, style=None, **kw):
    return mc._prof("pack_weights", 0.0, lambda: _orig_pack(weight, style, **kw),
                    detail=f"{tuple(weight.shape)} g{style.shape[0] if style is not None else 1}",
                    nbytes=weight.numel() * (4.0 + 2.0 * (style.shape[0] if style is not None else 1)))


def lin(l, x):
    return mc._prof("modulation_linear", 0.0, lambda: _orig_lin(l, x), detail=f"{tuple(l.weight.shape)}")


fastpath.upfirdn_nhwc = up
fastpath.to_rgb = rgb
mc.pack_weights = pack
fastpath._linear = lin

for _ in range(2):
    fastpath.restore_faces(net, dec, low, codes, [z])
torch.cuda.synchronize()
prof = mc.KernelProfiler()
mc.set_profiler(prof)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
fastpath.restore_faces(net, dec, low, codes, [z])
e.record()
rows = prof.table()
mc.set_profiler(None)
tot = s.elapsed_time(e) * 1e3
print(f"micro-batch {micro}: {tot:.0f} us wall (instrumented)")
agg = {}
inner = 0.0
for name, detail, flops, nbytes, sec in rows:
    us = sec * 1e6
    print(f"{name:26s} {detail:44s} {us:9.1f} us  {flops / sec / 1e12 if flops else 0:7.1f} TF/s  {nbytes / sec / 1e9:7.0f} GB/s")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] += flops
for k, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"SUM {k:26s} x{n:3d} {us:9.1f} us  {fl / us / 1e6 if fl else 0:7.1f} TF/s")
