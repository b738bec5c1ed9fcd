"""CPU oracle for the whole hot path: style decoder + Restoration_net forward (test infrastructure).

A functional torch-CPU restatement, driven by a plain ``state_dict``, of
/root/reference/e4e/models/stylegan2/model.py:475-552 (``Generator.forward`` with
``input_is_latent=True, return_features=True``) and /root/reference/models/RestoreNet.py:915-1046
(``encoder_forward`` + ``Restoration_net.forward``), built only from the oracle ops of this package
(``modulated_conv2d_ref``, ``upfirdn2d_native_port``, leaky-ReLU bias-act).  It is what
``bench.py`` times as the CPU baseline (``cpu_baseline.kind = "port"``, and ``--impl reference``)
and what the GPU tests compare the fused pipeline with at sizes other than the golden one.
Pinned by tests/test_oracle_golden.py::test_network_oracle_matches_reference against
tests/golden/networks.npz (outputs of the real reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .modconv_ref import modulated_conv2d_ref
from .upfirdn2d_ref import upfirdn2d_native_port

SQRT2 = math.sqrt(2.0)


def _blur_kernel(gain=1.0, dtype=torch.float32):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=dtype)
    k = torch.outer(k, k)
    return k / k.sum() * gain


def _lrelu(x, bias=None):
    """fused_leaky_relu (op/fused_act.py:216-228): slope 0.2, gain sqrt(2), bias on dim 1."""
    if bias is not None:
        x = x + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x, 0.2) * SQRT2


def _equal_linear(sd, p, x, lr_mul=1.0, act=False):
    """EqualLinear (models/RestoreNet.py:142-169)."""
    w = sd[p + "weight"]
    scale = lr_mul / math.sqrt(w.shape[1])
    b = sd[p + "bias"] * lr_mul
    if act:
        return _lrelu(F.linear(x, w * scale), b)
    return F.linear(x, w * scale, b)


def _noise(sd, p, x, noise):
    """NoiseInjection (models/RestoreNet.py:558-569)."""
    if noise is None:
        noise = torch.randn(x.shape[0], 1, x.shape[2], x.shape[3], dtype=x.dtype)
    return x + sd[p + "weight"] * noise


def _modconv(sd, p, x, style, demod=True, up=False, down=False, dilation=1, modulated=False):
    s = style if modulated else _equal_linear(sd, p + "modulation.", style)
    return modulated_conv2d_ref(x, sd[p + "weight"], s, demodulate=demod, upsample=up, downsample=down,
                                dilation=dilation)


def _styled_conv(sd, p, x, style, noise, up=False, down=False):
    """StyledConv / StyledConv_down (models/RestoreNet.py:571-643)."""
    y = _modconv(sd, p + "conv.", x, style, up=up, down=down)
    return _lrelu(_noise(sd, p + "noise.", y, noise), sd[p + "activate.bias"])


def _to_rgb(sd, p, x, style, skip=None):
    """ToRGB (models/RestoreNet.py:647-666)."""
    y = _modconv(sd, p + "conv.", x, style, demod=False) + sd[p + "bias"]
    if skip is not None:
        y = y + upfirdn2d_native_port(skip, _blur_kernel(4.0, x.dtype), 2, 1, (2, 1))
    return y


def _smart(sd, p, x, style, noise, rates=(1, 2, 4, 8)):
    """SMART_layer (models/RestoreNet.py:225-244)."""
    s = _equal_linear(sd, p + "modulation.", style)
    outs = [_modconv(sd, f"{p}ModulatedConv2ds.{j}.", x, s, dilation=r, modulated=True) for j, r in enumerate(rates)]
    w = sd[p + "fusion.0.weight"]
    y = F.conv2d(torch.cat(outs, 1), w / math.sqrt(w.shape[1] * 9), padding=1)
    y = _lrelu(y, sd[p + "fusion.1.bias"])
    return _lrelu(_noise(sd, p + "noise.", y, noise), sd[p + "activate.bias"])


def _large_conv(sd, p, x, k, rates=(1, 2, 4, 8)):
    """LargeConvLayer without downsampling (models/RestoreNet.py:725-787)."""
    outs = []
    for j, r in enumerate(rates):
        w = sd[f"{p}dilated_convs.{j}.weight"]
        outs.append(F.conv2d(x, w / math.sqrt(w.shape[1] * k * k), padding=((k - 1) * r) // 2, dilation=r))
    w = sd[p + "fusion.0.weight"]
    y = _lrelu(F.conv2d(torch.cat(outs, 1), w / math.sqrt(w.shape[1])), sd[p + "fusion.1.bias"])
    return _lrelu(y, sd[p + "activate.bias"])


def _style_mlp(sd, z, n_mlp, lr_mlp=0.01):
    x = z * torch.rsqrt(z.square().mean(dim=1, keepdim=True) + 1e-8)
    for i in range(1, n_mlp + 1):
        x = _equal_linear(sd, f"style.{i}.", x, lr_mul=lr_mlp, act=True)
    return x


@torch.no_grad()
def generator_ref(sd, codes, size, noise=None):
    """Style decoder with ``input_is_latent=True``; returns (image, [features])."""
    log_size = int(math.log2(size))
    n_layers = (log_size - 2) * 2 + 1
    noise = noise or [None] * n_layers
    b = codes.shape[0]
    out = sd["input.input"].repeat(b, 1, 1, 1)
    out = _styled_conv(sd, "conv1.", out, codes[:, 0], noise[0])
    skip = _to_rgb(sd, "to_rgb1.", out, codes[:, 1])
    feats = [out]
    i = 1
    for lvl in range(log_size - 2):
        out = _styled_conv(sd, f"convs.{2 * lvl}.", out, codes[:, i], noise[2 * lvl + 1], up=True)
        feats.append(out)
        out = _styled_conv(sd, f"convs.{2 * lvl + 1}.", out, codes[:, i + 1], noise[2 * lvl + 2])
        skip = _to_rgb(sd, f"to_rgbs.{lvl}.", out, codes[:, i + 2], skip)
        i += 2
    return skip, feats


@torch.no_grad()
def restoration_ref(sd, images, de_feats, pre_styles, z, size, n_mlp, noise=None):
    """Restoration_net.forward with one noise style ``z`` (models/RestoreNet.py:968-1046)."""
    log_size = int(math.log2(size))
    n_latent = log_size * 2 - 2
    n_layers = (log_size - 2) * 2 + 1
    if isinstance(noise, dict):
        # explicit per-layer noise for BOTH halves (test form): the reference's single list cannot express it, because its
        # encoder indexes the reversed decoder list and the down-convs' shapes do not match (:924-927)
        noise_rev, noise = list(noise["encoder"]), list(noise["decoder"])
    else:
        noise = noise or [None] * n_layers
        noise_rev = noise[::-1]
    b = images.shape[0]
    w_noise = _style_mlp(sd, z, n_mlp).unsqueeze(1).repeat(1, n_latent, 1)
    latent = torch.cat([pre_styles[:, :n_latent], w_noise], dim=-1)
    lat_rev = torch.flip(latent, dims=[1])
    # encoder (:915-942): SMART and its down-conv share the latent index
    out = _large_conv(sd, "down_from_big.", images, 1)
    features = []
    for lvl in range(log_size - 2):
        ii = 2 * lvl
        out = _smart(sd, f"encoder_convs.{ii}.", out, lat_rev[:, ii], noise_rev[ii])
        features.append(out)
        out = _styled_conv(sd, f"encoder_convs.{ii + 1}.", out, lat_rev[:, ii], noise_rev[ii + 1], down=True)
    out = _large_conv(sd, "final_layer.", out, 3)
    x_global = _equal_linear(sd, "final_linear.0.", out.reshape(b, -1), act=True)   # Dropout2d: eval -> identity
    early = _equal_linear(sd, "final_transfer.", x_global, act=True).reshape(b, -1, 4, 4)
    features.append(out + early)
    features = features[::-1]

    def sty(i):
        return torch.cat([latent[:, i], x_global], dim=1)

    out = _smart(sd, "conv1.", features[0], sty(0), noise[0])
    skip = _to_rgb(sd, "to_rgb1.", out, sty(1))
    i = 1
    for lvl in range(log_size - 2):
        out = _styled_conv(sd, f"convs.{2 * lvl}.", out, sty(i), noise[2 * lvl + 1], up=True)
        level = (i + 1) // 2
        out = out + features[level] + de_feats[level]
        out = _smart(sd, f"convs.{2 * lvl + 1}.", out, sty(i + 1), noise[2 * lvl + 2])
        skip = _to_rgb(sd, f"to_rgbs.{lvl}.", out, sty(i + 2), skip)
        i += 2
    return skip


@torch.no_grad()
def restore_faces_ref(net_sd, dec_sd, low, codes, z, size, dec_size, n_mlp, dec_noise=None, net_noise=None):
    """Hot path of one batch on the CPU: decoder features -> restoration network.  ``dec_noise``: per-layer list for the
    style decoder; ``net_noise``: list (reference semantics) or {"encoder": [...], "decoder": [...]} for the restorer."""
    image, feats = generator_ref(dec_sd, codes, dec_size, noise=dec_noise)
    restored = restoration_ref(net_sd, low, feats, codes, z, size, n_mlp, noise=net_noise)
    return restored, image
