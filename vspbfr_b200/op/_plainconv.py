"""The three convolution primitives behind ``conv2d_gradfix`` on the tcgen05 kernels: forward conv, transposed
conv (= input gradient) and weight gradient.  NCHW fp32 at the boundary (what the reference's layers pass,
models/RestoreNet.py:125-131, :547-553); operands are rounded to bf16, products accumulate in fp32.

Every shape the reference's networks produce goes through ``vsp_conv2d_gather_bf16`` / ``vsp_conv2d_wgrad_bf16``:
  * any kernel with <= 16 taps, per-axis padding and dilation, one stride (1 or 2) for both axes;
  * groups == 1, or the grouped form of the modulated convolution (input [1, G*Cin, H, W], groups == G, per-group
    weights: models/RestoreNet.py:547-553) — any other (N > 1, groups > 1) input is served sample by sample;
  * channel counts below / off a multiple of 8 (RGB stems, the 513-channel Discriminator head) are zero-padded to the
    16-byte channel vector the TMA boxes need.
There is no library (ATen / cuDNN) convolution here and no CPU path: an unsupported configuration raises.
"""
from __future__ import annotations

from typing import NamedTuple, Tuple

import torch

from .. import _lib
from . import modconv as mc


class ConvSpec(NamedTuple):
    """Static description of one convolution of the closed set (op/conv2d_gradfix.py:104-116 caches on the same fields)."""
    transpose: bool
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    output_padding: Tuple[int, int]
    dilation: Tuple[int, int]
    groups: int


def _check(spec: ConvSpec, kh: int, kw: int):
    if spec.stride[0] != spec.stride[1] or spec.stride[0] not in (1, 2):
        raise RuntimeError(f"conv2d_gradfix: stride {spec.stride} not supported (one stride of 1 or 2 for both axes)")
    if kh * kw > 16:
        raise RuntimeError(f"conv2d_gradfix: {kh}x{kw} kernel not supported (at most 16 taps)")
    if min(spec.dilation) < 1 or min(spec.padding) < 0 or min(spec.output_padding) < 0:
        raise RuntimeError("conv2d_gradfix: bad dilation / padding")


def _as_groups(t, groups):
    """[N, G*C, H, W] -> list of per-launch tensors shaped [G or N, C, H, W] (see the module docstring)."""
    n, ct, h, w = t.shape
    if groups == 1:
        return [t]
    c = ct // groups
    return [t[i].reshape(groups, c, h, w) for i in range(n)]


def _pack_group_weights(weight, groups, transpose):
    """weight [G*A, Bc, kh, kw] (per-group [A, Bc, kh, kw]) -> bf16 [G, taps, n, k_pad]: n = A, k = Bc for a forward
    conv; n = Bc, k = A for the transposed op (its weight is laid out [in, out/G, kh, kw])."""
    at, bc, kh, kw = weight.shape
    a = at // groups
    if groups == 1:
        return mc.pack_weights(weight, transpose=transpose)[0]
    wg = weight.reshape(groups, a, bc, kh * kw)
    wg = wg.permute(0, 3, 2, 1) if transpose else wg.permute(0, 3, 1, 2)          # [G, taps, n, k]
    k = wg.shape[3]
    out = torch.zeros((groups, kh * kw, wg.shape[2], mc._round_up(k, 8)), dtype=torch.bfloat16, device=weight.device)
    out[..., :k] = wg.to(torch.bfloat16)
    return out


def conv_forward(input, weight, bias, spec: ConvSpec):
    """``F.conv2d`` / ``F.conv_transpose2d`` semantics (op/conv2d_gradfix.py:138-147) on tcgen05."""
    if input.device.type != "cuda":
        raise RuntimeError("conv2d_gradfix: input must be a CUDA tensor (vspbfr_b200 has no CPU path)")
    if input.dtype != torch.float32 or input.ndim != 4 or weight.ndim != 4:
        raise RuntimeError("conv2d_gradfix: fp32 NCHW input and a 4-D weight are required")
    kh, kw = weight.shape[2], weight.shape[3]
    _check(spec, kh, kw)
    out = _transposed(input, weight, spec) if spec.transpose else _direct(input, weight, bias, spec)
    if spec.transpose and bias is not None:
        out = out + bias.reshape(1, -1, 1, 1)
    return out


def _direct(input, weight, bias, spec):
    n, c_total, h, w = input.shape
    g = spec.groups
    cout_total, cin, kh, kw = weight.shape
    if c_total != cin * g or cout_total % g:
        raise RuntimeError(f"conv2d: input {tuple(input.shape)} does not match weight {tuple(weight.shape)} (groups {g})")
    cout = cout_total // g
    s = spec.stride[0]
    (ph, pw), (dh, dw) = spec.padding, spec.dilation
    oh = (h + 2 * ph - dh * (kh - 1) - 1) // s + 1
    ow = (w + 2 * pw - dw * (kw - 1) - 1) // s + 1
    out = torch.empty((n, cout_total, max(oh, 0), max(ow, 0)), dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    wq = _pack_group_weights(weight.contiguous(), g, False)
    taps = [(i, j) for i in range(kh) for j in range(kw)]
    tap_w = [i * kw + j for i, j in taps]
    tap_dy = [i * dh - ph for i, j in taps]
    tap_dx = [j * dw - pw for i, j in taps]
    epi = mc.make_epilogue(bias=bias.contiguous()) if (bias is not None and g == 1) else None
    outs = out if g == 1 else None
    for idx, xg in enumerate(_as_groups(input, g)):
        xq = mc.nchw_to_nhwc_bf16(xg, c_pad=wq.shape[3])
        dst = outs if g == 1 else out[idx].view(g, cout, oh, ow)
        mc.conv_gather(xq, wq, cout, tap_w, tap_dy, tap_dx, s, (oh, ow), epi=epi, out=dst)
    if bias is not None and g != 1:
        out = out + bias.reshape(1, -1, 1, 1)
    return out


def _transposed(input, weight, spec):
    """``F.conv_transpose2d`` through ``modconv.conv_transposed`` (one gather launch per output parity class)."""
    n, c_total, ih, iw = input.shape
    g = spec.groups
    cin_total, cout, kh, kw = weight.shape           # transposed op: weight [in, out/G, kh, kw]
    if c_total != cin_total or cin_total % g:
        raise RuntimeError(f"conv_transpose2d: input {tuple(input.shape)} does not match weight {tuple(weight.shape)}")
    s = spec.stride[0]
    (ph, pw), (dh, dw), (oph, opw) = spec.padding, spec.dilation, spec.output_padding
    fh = (ih - 1) * s - 2 * ph + dh * (kh - 1) + oph + 1
    fw = (iw - 1) * s - 2 * pw + dw * (kw - 1) + opw + 1
    out = torch.empty((n, cout * g, max(fh, 0), max(fw, 0)), dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    wq = _pack_group_weights(weight.contiguous(), g, True)          # n = out channels, k = in channels
    for idx, xg in enumerate(_as_groups(input, g)):
        xq = mc.nchw_to_nhwc_bf16(xg, c_pad=wq.shape[3])
        dst = out if g == 1 else out[idx].view(g, cout, fh, fw)
        mc.conv_transposed(xq, wq, cout, kh, kw, s, pad=(ph, pw), dil=(dh, dw), out_pad=(oph, opw), out=dst)
    return out


def conv_weight_grad(grad_output, input, weight_shape, spec: ConvSpec):
    """Weight gradient of the op described by ``spec`` (replaces aten::cudnn_convolution(_transpose)_backward_weight,
    op/conv2d_gradfix.py:180-199): a pixel-K GEMM on tcgen05.  For the transposed op the roles of the two activations
    swap: gw[i, o, t] = sum_a input[a, i] * grad_output[a*s - p + t*d, o]."""
    if spec.transpose:
        sampled, pixels = grad_output, input        # `pixels` is indexed by a, `sampled` at a*s + t*d - p
    else:
        sampled, pixels = input, grad_output
    kh, kw = weight_shape[2], weight_shape[3]
    _check(spec, kh, kw)
    g, s, p, d = spec.groups, spec.stride[0], tuple(spec.padding), tuple(spec.dilation)
    n = input.shape[0]
    c_pix, c_smp = pixels.shape[1] // g, sampled.shape[1] // g      # gw rows / columns per group
    gw_total = torch.zeros(weight_shape, dtype=torch.float32, device=input.device) if (g > 1 and n > 1) else None
    if grad_output.numel() == 0 or input.numel() == 0:
        return torch.zeros(weight_shape, dtype=torch.float32, device=input.device)
    small = (not spec.transpose and g == 1 and kh == 1 and kw == 1 and c_smp <= 8 and s == 1 and p == (0, 0)
             and (input.shape[2] * input.shape[3]) % 4 == 0 and n <= 65535)
    if small:
        # RGB-side 1x1 layers (3 -> 16, 3 -> 64): a streaming reduction beats a GEMM whose N is 8 padded columns
        go, x = grad_output.contiguous(), input.contiguous()
        gw = torch.empty(weight_shape, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            rc = _lib.load().vsp_conv1x1_wgrad_small_f32(_lib.ptr(go), _lib.ptr(x), _lib.ptr(gw), n, c_pix, c_smp,
                                                         x.shape[2] * x.shape[3], _lib.stream_ptr())
        _lib.check(rc, "conv1x1_wgrad_small_f32")
        return gw
    result = None
    for pix_g, smp_g in zip(_as_groups(pixels, g), _as_groups(sampled, g)):
        pq = mc.nchw_to_nhwc_bf16(pix_g, c_pad=mc._round_up(c_pix, 8))
        sq = mc.nchw_to_nhwc_bf16(smp_g, c_pad=mc._round_up(c_smp, 8))
        gw = mc.conv_wgrad(pq, sq, g if g > 1 else 1, kh, kw, s, p, d)           # [G, taps, c_pix_pad, c_smp_pad]
        gw = gw[:, :, :c_pix, :c_smp].permute(0, 2, 3, 1).reshape(weight_shape)  # [G*c_pix, c_smp, kh, kw]
        if gw_total is None:
            result = gw.contiguous()
        else:
            gw_total += gw
    return result if gw_total is None else gw_total
