"""Oracle for upfirdn2d (test infrastructure; see oracle/__init__.py).

Semantics follow the reference's CPU path ``upfirdn2d_native``
(/root/reference/op/upfirdn2d.py:365-406) and its dispatcher (:346-362):
zero-stuff by ``up``, pad (negative pad = crop), TRUE convolution with ``kernel``
(= correlation with the flipped kernel, :388), keep every ``down``-th sample.

Two independent restatements are provided:
  * ``upfirdn2d_ref``        — numpy, direct polyphase gather (no zero-stuffed
                               intermediate), any float dtype (use float64 for a
                               high-precision check);
  * ``upfirdn2d_native_port`` — torch-CPU, the reference's own sequence of steps
                               (stuff -> pad/crop -> conv2d -> stride), used as
                               the timed CPU baseline because it is what the
                               reference executes on a CPU tensor.
"""
from __future__ import annotations

import numpy as np


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


def _pad4(pad):
    pad = tuple(int(p) for p in pad)
    return (pad[0], pad[1], pad[0], pad[1]) if len(pad) == 2 else pad  # op/upfirdn2d.py:353-354


def upfirdn2d_out_size(n_in, k, up, down, pad0, pad1):
    """op/upfirdn2d.py:301-302."""
    return (n_in * up + pad0 + pad1 - k + down) // down


def upfirdn2d_grad_pads(in_hw, out_hw, k_hw, up, down, pad):
    """Backward pads of op/upfirdn2d.py:309-312 -> (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)."""
    (in_h, in_w), (out_h, out_w), (kh, kw) = in_hw, out_hw, k_hw
    up_x, up_y = _pair(up)
    down_x, down_y = _pair(down)
    pad_x0, pad_x1, pad_y0, pad_y1 = _pad4(pad)
    return (kw - pad_x0 - 1,
            in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
            kh - pad_y0 - 1,
            in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)


def _tap_index(n_out, n_in, j, up, down, pad0):
    """For tap j: which outputs o read which input i (padded/stuffed index o*down + j)."""
    o = np.arange(n_out)
    u = o * down - pad0 + j
    ok = (u >= 0) & (u % up == 0) & (u // up < n_in)
    return o[ok], (u[ok] // up)


def upfirdn2d_ref(x, kernel, up=1, down=1, pad=(0, 0)):
    """x [N,C,H,W] ndarray, kernel [kh,kw] -> [N,C,H',W'] (same dtype as x)."""
    x = np.asarray(x)
    k = np.asarray(kernel, dtype=x.dtype)
    up_x, up_y = _pair(up)
    down_x, down_y = _pair(down)
    pad_x0, pad_x1, pad_y0, pad_y1 = _pad4(pad)
    n, c, in_h, in_w = x.shape
    kh, kw = k.shape
    out_h = upfirdn2d_out_size(in_h, kh, up_y, down_y, pad_y0, pad_y1)
    out_w = upfirdn2d_out_size(in_w, kw, up_x, down_x, pad_x0, pad_x1)
    out = np.zeros((n, c, max(out_h, 0), max(out_w, 0)), dtype=x.dtype)
    if out.size == 0:
        return out
    # out[o] = sum_j kflip[j] * padded[o*down + j],  kflip[j] = k[K-1-j]; y outer, x inner
    for jy in range(kh):
        oy, iy = _tap_index(out_h, in_h, jy, up_y, down_y, pad_y0)
        if oy.size == 0:
            continue
        for jx in range(kw):
            ox, ix = _tap_index(out_w, in_w, jx, up_x, down_x, pad_x0)
            if ox.size == 0:
                continue
            out[:, :, oy[:, None], ox[None, :]] += k[kh - 1 - jy, kw - 1 - jx] * x[:, :, iy[:, None], ix[None, :]]
    return out


def upfirdn2d_native_port(x, kernel, up=1, down=1, pad=(0, 0)):
    """torch-CPU port of the reference's CPU algorithm (op/upfirdn2d.py:365-406)."""
    import torch
    import torch.nn.functional as F

    up_x, up_y = _pair(up)
    down_x, down_y = _pair(down)
    pad_x0, pad_x1, pad_y0, pad_y1 = _pad4(pad)
    n, c, in_h, in_w = x.shape
    kh, kw = kernel.shape
    planes = x.reshape(n * c, 1, in_h, in_w)
    # 1) zero-stuffing: sample (y, x) lands at (y*up_y, x*up_x)
    stuffed = planes.new_zeros(n * c, 1, in_h * up_y, in_w * up_x)
    stuffed[:, :, ::up_y, ::up_x] = planes
    # 2) pad, then crop what negative pads remove
    stuffed = F.pad(stuffed, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    h, w = stuffed.shape[2], stuffed.shape[3]
    stuffed = stuffed[:, :, max(-pad_y0, 0): h - max(-pad_y1, 0), max(-pad_x0, 0): w - max(-pad_x1, 0)]
    # 3) true convolution == cross-correlation with the flipped filter
    filt = torch.flip(kernel.to(x.dtype), [0, 1]).reshape(1, 1, kh, kw)
    full = F.conv2d(stuffed, filt)
    # 4) decimate
    out = full[:, :, ::down_y, ::down_x]
    return out.reshape(n, c, out.shape[2], out.shape[3])
