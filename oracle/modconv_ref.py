"""Oracle for the (modulated) convolution layers (test infrastructure).

torch-CPU restatement of /root/reference/models/RestoreNet.py:478-555
(``ModulatedConv2d.forward``, fused branch), :334-418 (``Dilated_ModulatedConv2d``)
and the e4e twin e4e/models/stylegan2/model.py:237-278.  Works in fp32 or fp64 and
is differentiable by autograd to any order, which is what the gradient parity
tests compare the CUDA backward against.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .upfirdn2d_ref import upfirdn2d_native_port


def conv2d_ref(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Plain convolution = what conv2d_gradfix.conv2d reduces to (op/conv2d_gradfix.py:34-42)."""
    return F.conv2d(x, w, bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def _blur(x, blur_kernel, pad, gain=1.0):
    k = torch.as_tensor(blur_kernel, dtype=x.dtype)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k = k / k.sum() * gain
    return upfirdn2d_native_port(x, k, 1, 1, pad)


def modulated_conv2d_ref(x, weight, mod_style, demodulate=True, upsample=False, downsample=False,
                         dilation=1, blur_kernel=(1, 3, 3, 1), eps=1e-8):
    """x [B,Cin,H,W]; weight [1,Cout,Cin,k,k]; mod_style [B,Cin] = modulation(style), already applied.

    Per sample b: w_b = scale*W*s_b ; demod_b[o] = rsqrt(sum w_b^2 + eps) ; y_b = conv(x_b, w_b*demod_b).
    upsample: transposed conv stride 2 then blur(pad from RestoreNet.py:443-449, gain 4);
    downsample: blur(pad :451-457) then stride-2 conv.
    """
    b, cin, h, w_ = x.shape
    _, cout, _, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    wmod = scale * weight * mod_style.reshape(b, 1, cin, 1, 1)           # [B,Cout,Cin,k,k]
    if demodulate:
        d = torch.rsqrt(wmod.pow(2).sum(dim=(2, 3, 4)) + eps)
        wmod = wmod * d.reshape(b, cout, 1, 1, 1)
    outs = []
    for i in range(b):                                                   # one conv per sample
        xi = x[i:i + 1]
        wi = wmod[i]
        if upsample:
            p = (len(blur_kernel) - 2) - (k - 1) * dilation
            pad = ((p + 1) // 2 + 1, p // 2 + 1)
            yi = F.conv_transpose2d(xi, wi.transpose(0, 1), stride=2, padding=0, dilation=dilation)
            yi = _blur(yi, blur_kernel, pad, gain=4.0)
        elif downsample:
            p = (len(blur_kernel) - 2) + (k - 1)
            pad = ((p + 1) // 2, p // 2)
            yi = F.conv2d(_blur(xi, blur_kernel, pad), wi, stride=2, padding=0, dilation=dilation)
        else:
            yi = F.conv2d(xi, wi, padding=((k - 1) * dilation) // 2, dilation=dilation)
        outs.append(yi)
    return torch.cat(outs, 0)
