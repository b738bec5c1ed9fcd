#!/usr/bin/env python
"""Golden vectors for the W+ front end (vspbfr_b200/frontend.py) from the UNMODIFIED reference, CPU path:

    cp -r /root/reference /tmp/refprobe && chmod -R u+w /tmp/refprobe
    TORCH_CUDA_ARCH_LIST=10.0a VSP_REF=/tmp/refprobe python tests/golden/make_golden_frontend.py

The IR-SE50 encoder has ~200 M parameters, far too many to commit: reference and rebuild are constructed under the same
`torch.manual_seed`, so only the inputs, outputs and the (name, shape) manifest of the `state_dict`s are stored.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("VSP_REF", "/tmp/refprobe")
OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
m = types.ModuleType("matplotlib")
m.use = lambda *a, **k: None
sys.modules["matplotlib"] = m
import op  # noqa: E402,F401  (JIT-builds the reference extensions; the encoder module imports the stylegan2 model)

sys.modules["op.fused_act_cpu"] = sys.modules["op.fused_act"]
sys.modules["op.upfirdn2d_cpu"] = sys.modules["op.upfirdn2d"]
from e4e.models.encoders import psp_encoders  # noqa: E402
from models.CodeDiffuser import Code_diffuser  # noqa: E402
from ldm.ddpm import My_DDPM  # noqa: E402


def np_(t):
    return t.detach().cpu().numpy()


def manifest(sd):
    return np.array([f"{k}|{'x'.join(str(int(s)) for s in v.shape)}" for k, v in sd.items()])


out = {}
opts = types.SimpleNamespace(input_channel=3, stylegan_size=1024)
torch.manual_seed(77)
enc = psp_encoders.Encoder4Editing(50, "ir_se", opts).eval()
# BatchNorm running statistics / PReLU slopes away from their init so that eval-mode arithmetic is exercised
g = torch.Generator().manual_seed(78)
with torch.no_grad():
    for mod in enc.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
            mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
x = torch.randn(2, 3, 64, 64, generator=g)
with torch.no_grad():
    w = enc(x)
out["enc_x"], out["enc_w"], out["enc_manifest"] = np_(x), np_(w), manifest(enc.state_dict())
print("encoder", tuple(w.shape), float(w.abs().mean()), len(enc.state_dict()))

torch.manual_seed(79)
den = Code_diffuser(timesteps=4).eval()
ddpm = My_DDPM(denoise=den, timesteps=4, linear_start=0.1, linear_end=0.99).eval()
cond = torch.randn(3, 18, 512, generator=g)
x_t = torch.randn(3, 18, 512, generator=g)
t = torch.tensor([0, 2, 3])
with torch.no_grad():
    out["den_out"] = np_(den(x_t, cond, t))
    # deterministic replay of the inference branch (ldm/ddpm.py:421-429) from a fixed x_T
    cur = x_t
    for i in reversed(range(ddpm.num_timesteps)):
        cur, _ = ddpm.p_sample(cur, torch.full((3,), i, dtype=torch.long), cond, clip_denoised=ddpm.clip_denoised)
    out["ddpm_out"] = np_(cur)
out["den_cond"], out["den_xt"], out["den_t"] = np_(cond), np_(x_t), np_(t)
out["ddpm_manifest"] = manifest(ddpm.state_dict())
for name in ("betas", "posterior_mean_coef1", "posterior_mean_coef2"):
    out["ddpm_" + name] = np_(getattr(ddpm, name))
print("ddpm", float(cur.abs().mean()), len(ddpm.state_dict()))
np.savez_compressed(os.path.join(OUT, "frontend.npz"), **out)
print("wrote frontend.npz")
