#!/usr/bin/env python
"""GPU incumbents on the same B200: the REFERENCE's own CUDA operators (op/upfirdn2d_kernel.cu, op/fused_bias_act_kernel.cu,
rebuilt unmodified for sm_100a by tools/make_incumbent.py) and its own model files on cuDNN (fp32 and TF32), timed beside
this package's kernels on the same inputs: BASELINE configs[0] and the model's other upfirdn2d modes, configs[1]
(ModulatedConv2d 512->512 @64^2, batch 8, fwd and fwd+bwd) and the hot path of configs[2] (style decoder @1024 +
Restoration_net @512, batch 4).  Prints one JSON row per measurement; gpurun_out/incumbent.json holds them all."""
import json
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "VSPBFR")
if not os.path.isdir(REF):
    sys.exit("baseline/_ref/VSPBFR missing: run tools/make_incumbent.py in the build container first")
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
m = types.ModuleType("matplotlib"); m.use = lambda *a, **k: None; sys.modules["matplotlib"] = m   # e4e/models/psp.py:1-3
import op as ref_op                                              # noqa: E402  the reference's package (its own CUDA kernels)
from models import RestoreNet as ref_R                           # noqa: E402
from e4e.models.stylegan2 import model as ref_S                  # noqa: E402
from vspbfr_b200 import fastpath, layers, op as our_op            # noqa: E402
from vspbfr_b200.restorenet import Restoration_net               # noqa: E402
from vspbfr_b200.stylegan2 import Generator                      # noqa: E402

assert ref_op.__file__.startswith(REF), ref_op.__file__
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM, TF = PEAKS.get("hbm_gbs", 6650.0), PEAKS.get("bf16_tflops", 1590.0)
DEV = "cuda"
rows = []


def rotate(fns, rounds=5):
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
    return best


def rec(name, impl, t, nbytes=0.0, flops=0.0):
    row = {"case": name, "impl": impl, "us": t * 1e6}
    if nbytes:
        row.update(GBps=nbytes / t / 1e9, frac_hbm=nbytes / t / 1e9 / HBM)
    if flops:
        row.update(TFLOPs=flops / t / 1e12, frac_bf16_burst=flops / t / 1e12 / TF)
    rows.append(row)
    print(json.dumps(row), flush=True)


@torch.no_grad()
def ops():
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device=DEV)
    k = torch.outer(k1, k1) / 64
    k4 = (k * 4).contiguous()
    b = torch.randn(512, device=DEV)
    cases = [
        ("upfirdn2d up2 [4,512,64,64] (configs[0])", (4, 512, 64, 64), lambda f, x: f(x, k4, up=2, down=1, pad=(2, 1)), 5 * 4 * 512 * 64 * 64 * 4),
        ("upfirdn2d down2 [4,512,128,128]", (4, 512, 128, 128), lambda f, x: f(x, k4, up=1, down=2, pad=(1, 1)), 5 * 4 * 512 * 64 * 64 * 4),
        ("blur pad(1,1) [4,512,65,65]", (4, 512, 65, 65), lambda f, x: f(x, k4, pad=(1, 1)), (4 * 512 * 65 * 65 + 4 * 512 * 64 * 64) * 4),
        ("blur pad(1,1) [4,32,1025,1025]", (4, 32, 1025, 1025), lambda f, x: f(x, k4, pad=(1, 1)), (4 * 32 * 1025 * 1025 + 4 * 32 * 1024 * 1024) * 4),
        ("blur pad(2,2) [4,64,512,512]", (4, 64, 512, 512), lambda f, x: f(x, k, pad=(2, 2)), (4 * 64 * 512 * 512 + 4 * 64 * 513 * 513) * 4),
    ]
    for name, shape, call, nbytes in cases:
        n = max(4, int(math.ceil(1.2e9 / nbytes)))
        xs = [torch.randn(*shape, device=DEV) for _ in range(n)]
        for impl, f in (("reference CUDA op (sm_100a rebuild)", ref_op.upfirdn2d), ("vspbfr_b200", our_op.upfirdn2d)):
            rec(name, impl, rotate([(lambda x=x: call(f, x)) for x in xs]), nbytes)
        del xs
        torch.cuda.empty_cache()
    for shape in ((4, 512, 128, 128), (4, 512, 64, 64)):
        nbytes = 2 * 4 * int(np.prod(shape))
        xs = [torch.randn(*shape, device=DEV) for _ in range(max(4, int(math.ceil(1.2e9 / nbytes))))]
        for impl, f in (("reference CUDA op (sm_100a rebuild)", ref_op.fused_leaky_relu), ("vspbfr_b200", our_op.fused_leaky_relu)):
            rec(f"fused_leaky_relu {list(shape)}", impl, rotate([(lambda x=x: f(x, b)) for x in xs]), nbytes)
        del xs
        torch.cuda.empty_cache()


def config2():
    bsz, c, h = 8, 512, 64
    flops = 2.0 * bsz * c * c * 9 * h * h
    torch.manual_seed(0)
    ref = ref_R.ModulatedConv2d(c, c, 3, 512).to(DEV)
    ours = layers.ModulatedConv2d(c, c, 3, 512).to(DEV)
    ours.load_state_dict(ref.state_dict())
    xs = [torch.randn(bsz, c, h, h, device=DEV, requires_grad=True) for _ in range(6)]
    sts = [torch.randn(bsz, 512, device=DEV, requires_grad=True) for _ in range(6)]
    dys = [torch.randn(bsz, c, h, h, device=DEV) for _ in range(6)]
    for impl, mod, tf32 in (("reference module, cuDNN grouped conv fp32", ref, False),
                            ("reference module, cuDNN grouped conv TF32", ref, True), ("vspbfr_b200 module (eager)", ours, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        params = list(mod.parameters())
        with torch.no_grad():
            rec("configs[1] ModulatedConv2d fwd", impl, rotate([(lambda x=x, s=s: mod(x, s)) for x, s in zip(xs, sts)], rounds=3), flops=flops)
        rec("configs[1] ModulatedConv2d fwd+bwd", impl,
            rotate([(lambda x=x, s=s, dy=dy: torch.autograd.grad(mod(x, s), [x, s] + params, dy)) for x, s, dy in zip(xs, sts, dys)],
                   rounds=3), flops=3 * flops)


@torch.no_grad()
def config3_hot_path():
    bsz = 4
    torch.manual_seed(0)
    rnet = ref_R.Restoration_net(512, 512, 8, channel_multiplier=2).to(DEV).eval()
    rdec = ref_S.Generator(1024, 512, 8, channel_multiplier=2).to(DEV).eval()
    net = Restoration_net(512, 512, 8, channel_multiplier=2).to(DEV).eval()
    dec = Generator(1024, 512, 8, channel_multiplier=2).to(DEV).eval()
    net.load_state_dict(rnet.state_dict())
    dec.load_state_dict(rdec.state_dict())
    g = torch.Generator().manual_seed(3)
    low = (torch.rand(bsz, 3, 512, 512, generator=g) * 2 - 1).to(DEV)
    codes = torch.randn(bsz, 18, 512, generator=g).to(DEV)
    z = torch.randn(bsz, 512, generator=g).to(DEV)
    pool = torch.nn.AdaptiveAvgPool2d((512, 512))

    def ref_run():
        img, feats = rdec([codes], input_is_latent=True, randomize_noise=True, return_features=True)   # psp.py:235-248
        pool(img)
        return rnet(low, feats[:16], codes, [z])

    def time_it(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) * 1e-3 / n

    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        t = time_it(ref_run)
        rec("configs[2] hot path, batch 4 (decoder@1024 + Restoration_net@512)",
            f"reference modules + reference CUDA ops + cuDNN {'TF32' if tf32 else 'fp32'}", t)
        rows[-1]["faces_per_s"] = bsz / t
    t = time_it(lambda: fastpath.restore_faces(net, dec, low, codes, [z]))
    rec("configs[2] hot path, batch 4 (decoder@1024 + Restoration_net@512)", "vspbfr_b200 fused path (eager)", t)
    rows[-1]["faces_per_s"] = bsz / t
    gr = fastpath.GraphedRestorer(net, dec, bsz, device=DEV)
    t = time_it(lambda: gr(low, codes, z, clone=False), n=10)
    rec("configs[2] hot path, batch 4 (decoder@1024 + Restoration_net@512)", "vspbfr_b200 fused path (CUDA graph)", t)
    rows[-1]["faces_per_s"] = bsz / t
    # the reference pipeline at the benchmark's micro-batch (its int32 indexing allows <= 32 at 1024^2)
    for bsz2 in (16,):
        low2, codes2, z2 = low.repeat(bsz2 // bsz, 1, 1, 1), codes.repeat(bsz2 // bsz, 1, 1), z.repeat(bsz2 // bsz, 1)

        def ref_run2():
            img, feats = rdec([codes2], input_is_latent=True, randomize_noise=True, return_features=True)
            pool(img)
            return rnet(low2, feats[:16], codes2, [z2])
        torch.backends.cudnn.allow_tf32 = True
        t = time_it(ref_run2, n=3)
        rec(f"hot path, batch {bsz2}", "reference modules + reference CUDA ops + cuDNN TF32", t)
        rows[-1]["faces_per_s"] = bsz2 / t


if __name__ == "__main__":
    ops()
    config2()
    config3_hot_path()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "incumbent.json"), "w"), indent=1)
