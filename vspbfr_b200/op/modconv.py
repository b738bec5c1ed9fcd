"""Host side of the tcgen05 convolution: layout conversion, weight prologue, conv launches and the
autograd Functions that implement the modulated convolution and the plain-conv primitives of
conv2d_gradfix on top of them.

Reference semantics: models/RestoreNet.py:478-555 (ModulatedConv2d.forward, fused branch),
:334-418 (Dilated_ModulatedConv2d), op/conv2d_gradfix.py:134-223 (closed gradient set).
Precision contract (north_star): operands are rounded to bf16, products accumulate in fp32 on
the tensor cores, demodulation is applied in fp32 in the epilogue.
"""
from __future__ import annotations

import ctypes
import math

import torch
from torch.autograd import Function

from .. import _lib
from .._lib import ConvEpilogue, ptr, stream_ptr


def _round_up(v, m):
    return (v + m - 1) // m * m


class KernelProfiler:
    """Optional per-launch accounting used by bench.py for the roofline numbers: CUDA events around
    every tcgen05 conv launch on the launching stream plus the launch's ALGORITHMIC flops
    (2 * pixels * Cout * Cin * taps).  Off (None) by default: no overhead on the product path."""

    def __init__(self):
        self.records = []
        self.details = []

    def wrap(self, name, flops, fn, detail="", nbytes=0.0):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        self.records.append((name, float(flops), s, e))
        self.details.append((detail, float(nbytes)))
        return out

    def table(self):
        """Per-launch rows (name, detail, flops, bytes, seconds) in launch order."""
        torch.cuda.synchronize()
        return [(n, d, f, nb, s.elapsed_time(e) * 1e-3) for (n, f, s, e), (d, nb) in zip(self.records, self.details)]

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for (name, flops, s, e), (_, nbytes) in zip(self.records, self.details):
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += flops
            a[2] += s.elapsed_time(e) * 1e-3
            a[3] += nbytes
        return {k: {"launches": v[0], "flops": v[1], "seconds": v[2], "bytes": v[3]} for k, v in agg.items()}


_profiler = None


def set_profiler(p):
    global _profiler
    _profiler = p


def _prof(name, flops, fn, detail="", nbytes=0.0):
    return fn() if _profiler is None else _profiler.wrap(name, flops, fn, detail, nbytes)


# ----------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ----------------------------------------------------------------------------------------------

def nchw_to_nhwc_bf16(x, scale_nc=None, c_pad=None, dot_with=None):
    """[N,C,H,W] fp32 -> [N,H,W,c_pad] bf16 (optionally * scale_nc[n,c]); extra channels are zero.
    ``dot_with`` (same shape as x): additionally returns dot[n,c] = sum_p x[n,c,p] * dot_with[n,c,p] (x unscaled) from the
    same pass — the demodulation gradient sum_p dy*y rides on the conversion of dy."""
    n, c, h, w = x.shape
    c_pad = c_pad or _round_up(c, 8)
    xc = x.contiguous()
    y = torch.empty((n, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    dot = None
    if dot_with is not None:
        if (h * w) % 4 or c_pad == 8 or (xc.data_ptr() | dot_with.data_ptr()) & 15:     # shapes the fused reduction does not cover
            dot, dot_with = (xc * dot_with).sum((2, 3)), None
        else:
            dot_with = dot_with.contiguous()
            dot = torch.empty((n, c), dtype=torch.float32, device=x.device)
    if y.numel():
        with _lib.device_guard(x.device):
            rc = _lib.load().vsp_nchw_f32_to_nhwc_bf16_dot(ptr(xc), ptr(scale_nc), ptr(dot_with), ptr(y),
                                                           ptr(dot) if dot_with is not None else None, n, c, h * w, c_pad,
                                                           stream_ptr())
        _lib.check(rc, "nchw_f32_to_nhwc_bf16")
    return y if dot is None else (y, dot)


def nhwc_bf16_to_nchw(x, c=None, scale_nc=None, dot_with=None):
    """[N,H,W,c_pad] bf16 -> [N,c,H,W] fp32, optionally * scale_nc[n,c]; ``dot_with`` [N,c,H,W] fp32 additionally
    returns dot[n,c] = sum_p x[n,p,c] * dot_with[n,c,p] (x unscaled) — the style gradient sum_p x*d(x*s)."""
    n, h, w, c_pad = x.shape
    c = c or c_pad
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    fused = (h * w) % 4 == 0 and c_pad % 8 == 0
    if not fused and (scale_nc is not None or dot_with is not None):
        y = nhwc_bf16_to_nchw(x, c)
        dot = (y * dot_with).sum((2, 3)) if dot_with is not None else None
        if scale_nc is not None:
            y = y * scale_nc.reshape(n, c, 1, 1)
        return y if dot_with is None else (y, dot)
    dot = None
    if dot_with is not None:
        dot_with = dot_with.contiguous()
        dot = torch.empty((n, c), dtype=torch.float32, device=x.device)
    if scale_nc is not None:
        scale_nc = scale_nc.contiguous()
    if y.numel():
        with _lib.device_guard(x.device):
            rc = _lib.load().vsp_nhwc_bf16_to_nchw_f32_dot(ptr(x.contiguous()), ptr(scale_nc), ptr(dot_with), ptr(y), ptr(dot),
                                                           n, c, h * w, c_pad, stream_ptr())
        _lib.check(rc, "nhwc_bf16_to_nchw_f32")
    return y if dot_with is None else (y, dot)


def nchw_to_bf16(x, scale_nc=None):
    """[N,C,H,W] fp32 -> same layout bf16, optionally * scale_nc[n,c]."""
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty_like(x, dtype=torch.bfloat16)
    if y.numel():
        with _lib.device_guard(x.device):
            rc = _lib.load().vsp_nchw_f32_to_bf16(ptr(x), ptr(scale_nc), ptr(y), n * c, h * w, stream_ptr())
        _lib.check(rc, "nchw_f32_to_bf16")
    return y


def weight_sumsq(weight):
    """wsq[o,i] = sum over taps of W^2 — the style-independent half of the demodulation sum."""
    cout, cin, kh, kw = weight.shape
    weight = weight.contiguous()
    wsq = torch.empty((cout, cin), dtype=torch.float32, device=weight.device)
    with _lib.device_guard(weight.device):
        rc = _lib.load().vsp_weight_sumsq_f32(ptr(weight), ptr(wsq), cout, cin, kh * kw, stream_ptr())
    _lib.check(rc, "weight_sumsq_f32")
    return wsq


def pack_weights(weight, style=None, wscale=1.0, eps=1e-8, transpose=False, want_demod=False, fold_demod=False,
                 batch=None, wsq=None):
    """weight [Cout,Cin,kh,kw] fp32 (+ style [B,Cin]) -> (wq bf16 [G,taps,n_pad,k_pad], demod [G,Cout] | None).

    G = B when a style is given (per-sample modulated weights), else 1.
    """
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    g = style.shape[0] if style is not None else 1
    n_real, k_real = (cin, cout) if transpose else (cout, cin)
    n_pad, k_pad = n_real, _round_up(k_real, 8)
    weight = weight.contiguous()
    if style is not None:
        style = style.contiguous()
    wq = torch.empty((g, taps, n_pad, k_pad), dtype=torch.bfloat16, device=weight.device)
    demod = torch.empty((g, cout), dtype=torch.float32, device=weight.device) if (want_demod or fold_demod) else None
    with _lib.device_guard(weight.device):
        rc = _lib.load().vsp_modulate_weights_bf16(ptr(weight), ptr(style), ptr(demod), ptr(wq), g, cout, cin, taps,
                                                   wscale, eps, int(transpose), int(fold_demod), n_pad, k_pad,
                                                   ptr(wsq), stream_ptr())
    _lib.check(rc, "modulate_weights_bf16")
    return wq, demod


def make_epilogue(row_scale=None, noise=None, noise_weight=0.0, noise_weight_dev=None, bias=None, act=0, alpha=0.2,
                  scale=1.0, residual=None, residual2=None, pre_bias=None, pre_act=0, alpha_vec=None):
    """Build a ``vsp_conv_epilogue``; returns (struct, keepalive tuple)."""
    e = ConvEpilogue()
    e.row_scale = row_scale.data_ptr() if row_scale is not None else None
    e.noise = noise.data_ptr() if noise is not None else None
    if noise is not None:
        e.noise_bstride = 0 if noise.shape[0] == 1 else noise[0].numel()
    e.noise_weight = float(noise_weight)
    e.noise_weight_dev = noise_weight_dev.data_ptr() if noise_weight_dev is not None else None
    e.bias = bias.data_ptr() if bias is not None else None
    e.act, e.alpha, e.scale = int(act), float(alpha), float(scale)
    e.pre_bias = pre_bias.data_ptr() if pre_bias is not None else None
    e.pre_act = int(pre_act)
    e.residual = residual.data_ptr() if residual is not None else None
    e.residual2 = residual2.data_ptr() if residual2 is not None else None
    e.alpha_vec = alpha_vec.data_ptr() if alpha_vec is not None else None
    return e, (row_scale, noise, noise_weight_dev, bias, residual, residual2, pre_bias, alpha_vec)


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def conv_out_size(n, k, stride, pad, dil):
    return (n + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _rows_per_tap(wq):
    """Row pitch of a packed weight tensor [G,taps,n,k] — also valid for a channel-slice VIEW of it."""
    assert wq.stride(3) == 1 and wq.stride(2) == wq.shape[3], "packed weights must be dense in (n, k)"
    if wq.shape[1] > 1:
        return wq.stride(1) // wq.stride(2)
    if wq.shape[0] > 1:
        return wq.stride(0) // wq.stride(2)
    return wq.shape[2]


def conv_fprop(x_nhwc, wq, cout, kh, kw, stride=1, pad=0, dil=1, epi=None, out=None, out_nhwc=False, co_off=0,
               algo_cin=None):
    """out = conv(x_nhwc, wq) (+ epilogue).  x_nhwc [B,H,W,Cin_pad] bf16, wq [G,taps,cout_pad,Cin_pad] bf16."""
    b, h, w, cin = x_nhwc.shape
    g, taps, _, k_pad = wq.shape
    assert taps == kh * kw and k_pad == cin, (wq.shape, x_nhwc.shape, kh, kw)
    cout_pad = _rows_per_tap(wq)
    oh, ow = conv_out_size(h, kh, stride, pad, dil), conv_out_size(w, kw, stride, pad, dil)
    if out is None:
        if out_nhwc:
            out = torch.empty((b, oh, ow, _round_up(cout, 8)), dtype=torch.bfloat16, device=x_nhwc.device)
            if out.shape[3] != cout:
                out.zero_()
        else:
            out = torch.empty((b, cout, oh, ow), dtype=torch.float32, device=x_nhwc.device)
    ldo = out.shape[3] if out_nhwc else cout
    e, keep = epi if epi is not None else (None, None)
    with _lib.device_guard(x_nhwc.device):
        # algo_cin: channels of the layer this launch stands for when the operand carries extra terms (two-term operands)
        rc = _prof("conv_fprop", 2.0 * b * oh * ow * cout * (algo_cin or cin) * kh * kw, lambda: _lib.load().vsp_conv2d_fprop_bf16(
            ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, cout_pad, kh, kw, stride, pad, dil, int(out_nhwc),
            ldo, co_off, ctypes.byref(e) if e is not None else None, stream_ptr()),
            detail=f"b{b} {cin}->{cout} k{kh} s{stride} d{dil} {h}x{w} g{g}",
            nbytes=2.0 * b * (h * w * cin + oh * ow * cout * (2 if not out_nhwc else 1)))
    _lib.check(rc, "conv2d_fprop_bf16")
    return out


def conv_gather(x_nhwc, wq, cout, tap_w, tap_dy, tap_dx, stride, out_hw, full_hw=None, os_=1, oo=(0, 0), epi=None,
                out=None, out_nhwc=False, co_off=0):
    """General tap-list convolution (input gradients, parity classes); see vsp_conv2d_gather_bf16."""
    b, h, w, cin = x_nhwc.shape
    g, taps_total, _, k_pad = wq.shape
    assert k_pad == cin
    cout_pad = _rows_per_tap(wq)
    oh, ow = out_hw
    fh, fw = full_hw or out_hw
    if out is None:
        if out_nhwc:
            out = torch.zeros((b, fh, fw, _round_up(cout, 8)), dtype=torch.bfloat16, device=x_nhwc.device)
        else:
            out = torch.zeros((b, cout, fh, fw), dtype=torch.float32, device=x_nhwc.device)
    ldo = out.shape[3] if out_nhwc else cout
    e, keep = epi if epi is not None else (None, None)
    with _lib.device_guard(x_nhwc.device):
        rc = _lib.load().vsp_conv2d_gather_bf16(ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, cout_pad,
                                                taps_total, len(tap_w), _int_array(tap_w), _int_array(tap_dy),
                                                _int_array(tap_dx), stride, oh, ow, int(out_nhwc), fh, fw, os_,
                                                oo[0], oo[1], ldo, co_off,
                                                ctypes.byref(e) if e is not None else None, stream_ptr())
    _lib.check(rc, "conv2d_gather_bf16")
    return out


def conv_transpose_s2(x_nhwc, wq, cout, kh, kw, epi=None, out_nhwc=False):
    """Stride-2, padding-0 transposed convolution -> extent ((H-1)*2+kh, (W-1)*2+kw)."""
    b, h, w, cin = x_nhwc.shape
    g, taps, _, k_pad = wq.shape
    assert taps == kh * kw and k_pad == cin
    cout_pad = _rows_per_tap(wq)
    fh, fw = (h - 1) * 2 + kh, (w - 1) * 2 + kw
    if out_nhwc:
        out = torch.empty((b, fh, fw, _round_up(cout, 8)), dtype=torch.bfloat16, device=x_nhwc.device)
        if out.shape[3] != cout or kh == 1 or kw == 1:
            out.zero_()
    else:
        out = torch.empty((b, cout, fh, fw), dtype=torch.float32, device=x_nhwc.device)
        if kh == 1 or kw == 1:
            out.zero_()
    ldo = out.shape[3] if out_nhwc else cout
    e, keep = epi if epi is not None else (None, None)
    with _lib.device_guard(x_nhwc.device):
        rc = _prof("conv_transpose_s2", 2.0 * b * h * w * cout * cin * kh * kw,
                   lambda: _lib.load().vsp_conv_transpose2d_s2_bf16(
                       ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, cout_pad, kh, kw, int(out_nhwc), ldo, 0,
                       ctypes.byref(e) if e is not None else None, stream_ptr()),
                   detail=f"b{b} {cin}->{cout} k{kh} up2 {h}x{w} g{g}",
                   nbytes=2.0 * b * (h * w * cin + (2 * h + 1) * (2 * w + 1) * cout))
    _lib.check(rc, "conv_transpose2d_s2_bf16")
    return out


def scale_nhwc(x_nhwc, s):
    """bf16(x[b,h,w,c] * s[b,c]) — input-side style modulation (models/RestoreNet.py:481-508, `fused=False`)."""
    b, h, w, c = x_nhwc.shape
    if s.shape[1] != c:
        s = torch.nn.functional.pad(s, (0, c - s.shape[1]))
    s = s.contiguous()
    y = torch.empty_like(x_nhwc)
    with _lib.device_guard(x_nhwc.device):
        rc = _lib.load().vsp_scale_nhwc_bf16(ptr(x_nhwc), ptr(s), ptr(y), b, h * w, c, stream_ptr())
    _lib.check(rc, "scale_nhwc_bf16")
    return y


def demod_from_wsq(s, wsq, wscale, eps=1e-8):
    """d[b,o] = rsqrt(wscale^2 * sum_i s[b,i]^2 * wsq[o,i] + eps) (models/RestoreNet.py:513-516 with cached sum_t W^2)."""
    b, cin = s.shape
    cout = wsq.shape[0]
    d = torch.empty((b, cout), dtype=torch.float32, device=s.device)
    with _lib.device_guard(s.device):
        rc = _lib.load().vsp_modulate_weights_bf16(ptr(wsq), ptr(s.contiguous()), ptr(d), None, b, cout, cin, 1, wscale, eps,
                                                   0, 0, cout, cin, ptr(wsq), stream_ptr())
    _lib.check(rc, "modulate_weights_bf16(demod)")
    return d


def conv_branches(x_nhwc, wq, cout, dils, epi=None, out=None, out_nhwc=True, co_off=0, algo_cin=None):
    """The dilated 3x3 branches of a SMART layer in one launch (vsp_conv2d_branches_bf16)."""
    b, h, w, cin = x_nhwc.shape
    g, taps, rows, k_pad = wq.shape
    assert taps == 9 and k_pad == cin and rows == cout, (wq.shape, cout, cin)
    if out is None:
        out = (torch.empty((b, h, w, cout), dtype=torch.bfloat16, device=x_nhwc.device) if out_nhwc
               else torch.empty((b, cout, h, w), dtype=torch.float32, device=x_nhwc.device))
    ldo = out.shape[3] if out_nhwc else cout
    e, keep = epi if epi is not None else (None, None)
    dl = _int_array(dils)
    with _lib.device_guard(x_nhwc.device):
        rc = _prof("conv_branches", 2.0 * b * h * w * cout * (algo_cin or cin) * 9, lambda: _lib.load().vsp_conv2d_branches_bf16(
            ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, len(dils), dl, int(out_nhwc), ldo, co_off,
            ctypes.byref(e) if e is not None else None, stream_ptr()),
            detail=f"b{b} {cin}->{cout} k3 x{len(dils)} branches {h}x{w} g{g}", nbytes=2.0 * b * h * w * (cin + cout))
    _lib.check(rc, "conv2d_branches_bf16")
    return out


def compose_up2_weights(weight, blur_kernel):
    """Composite weights of `conv_transpose2d(stride=2, 3x3)` followed by `Blur(4x4, pad=(1,1))`
    (models/RestoreNet.py:522-535), split into the four output-parity classes of the 6x6 stride-2 kernel.

    weight [Cout, Cin, 3, 3] (ModulatedConv2d.weight[0]), blur_kernel [4, 4] (already x upsample_factor^2)
    -> [4*Cout, Cin, 3, 3] with row (pa*2+pb)*Cout + o and tap (dy+1, dx+1) on the low-res grid:
    out[2A+pa, 2B+pb] = sum_{dy,dx} x[A+dy, B+dx] * W6[pa - 2*dy + 2, pb - 2*dx + 2],  W6 = weight (*) blur (full conv).
    """
    cout, cin, kh, kw = weight.shape
    assert kh == 3 and kw == 3 and tuple(blur_kernel.shape) == (4, 4)
    w6 = torch.nn.functional.conv2d(weight.reshape(cout * cin, 1, 3, 3).double(),
                                    torch.flip(blur_kernel, [0, 1]).reshape(1, 1, 4, 4).double(), padding=3)
    w6 = w6.reshape(cout, cin, 6, 6)
    out = torch.empty((4, cout, cin, 3, 3), dtype=torch.float64, device=weight.device)
    for pa in range(2):
        for pb in range(2):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    out[pa * 2 + pb, :, :, dy + 1, dx + 1] = w6[:, :, pa - 2 * dy + 2, pb - 2 * dx + 2]
    return out.reshape(4 * cout, cin, 3, 3).float().contiguous()


def conv_up2_fused(x_nhwc, wq, cout, epi=None, out=None, co_off=0):
    """Fused transposed-conv + blur (see vsp_conv2d_up2_fused_bf16): x [B,H,W,Cin] -> [B,2H,2W,ldo] NHWC bf16."""
    b, h, w, cin = x_nhwc.shape
    g, taps, rows, k_pad = wq.shape
    assert taps == 9 and k_pad == cin and rows == 4 * cout, (wq.shape, cout)
    if out is None:
        out = torch.empty((b, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    e, keep = epi if epi is not None else (None, None)
    with _lib.device_guard(x_nhwc.device):
        # algorithmic FLOPs of the layer it replaces (the transposed conv); the dense form executes 4x as many
        rc = _prof("conv_up2_fused", 2.0 * b * h * w * cout * cin * 9,
                   lambda: _lib.load().vsp_conv2d_up2_fused_bf16(
                       ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, out.shape[3], co_off,
                       ctypes.byref(e) if e is not None else None, stream_ptr()),
                   detail=f"b{b} {cin}->{cout} up2-fused {h}x{w} g{g}",
                   nbytes=2.0 * b * (h * w * cin + 4 * h * w * cout))
    _lib.check(rc, "conv2d_up2_fused_bf16")
    return out


def compose_up2h_weights(weight, fx):
    """Weights of `conv_transpose2d(stride=2, 3x3)` composed with the HORIZONTAL factor of the blur only (the vertical factor
    is applied by the kernel's epilogue, see vsp_conv2d_up2h_bf16).

    weight [Cout, Cin, 3, 3] (ModulatedConv2d.weight[0]), fx: the 4 horizontal taps (blur == outer(fy, fx))
    -> [2*Cout, Cin, 3, 3] with row q*Cout + o (q = output column parity) and tap (kh, dx + 1), dx = -1..1 on the low-res grid:
    hz[2a+pa][2j+q] = sum_{kh = pa (mod 2)} sum_dx x[a - (kh-pa)/2][j + dx] * Wc[q*Cout + o, :, kh, dx + 1],
    Wc[.., kh, dx + 1] = sum_{v, kw : q + v - 1 - kw = 2 dx} fx[3 - v] * W[.., kh, kw]   (upfirdn2d correlates with the flip).
    """
    cout, cin, kh, kw = weight.shape
    assert kh == 3 and kw == 3 and len(fx) == 4
    w = weight.double()
    out = torch.zeros((2, cout, cin, 3, 3), dtype=torch.float64, device=weight.device)
    for q in range(2):
        for v in range(4):
            for kwi in range(3):
                t = q + v - 1 - kwi
                if t % 2 == 0 and -1 <= t // 2 <= 1:
                    out[q, :, :, :, t // 2 + 1] += float(fx[3 - v]) * w[:, :, :, kwi]
    return out.reshape(2 * cout, cin, 3, 3).float().contiguous()


def up2h_supported(cin, cout, h, w):
    """Shapes vsp_conv2d_up2h_bf16 takes (the wide levels of both networks)."""
    return cin in (64, 128) and cout % 32 == 0 and w >= 128 and w % 32 == 0 and h >= 4


def conv_up2h(x_nhwc, wq, cout, ky, epi=None, out=None, co_off=0):
    """Up-sampling conv with the horizontal blur in the weights and the vertical blur in the epilogue (see
    vsp_conv2d_up2h_bf16): x [B,H,W,Cin] -> [B,2H,2W,ldo] NHWC bf16.  ``ky``: 4 vertical taps as applied (host floats)."""
    b, h, w, cin = x_nhwc.shape
    g, taps, rows, k_pad = wq.shape
    assert taps == 9 and k_pad == cin and rows == 2 * cout, (wq.shape, cout)
    if out is None:
        out = torch.empty((b, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    e, keep = epi if epi is not None else (None, None)
    kyc = ky if isinstance(ky, ctypes.Array) else (ctypes.c_float * 4)(*[float(v) for v in ky])
    with _lib.device_guard(x_nhwc.device):
        rc = _prof("conv_up2_fused", 2.0 * b * h * w * cout * cin * 9,
                   lambda: _lib.load().vsp_conv2d_up2h_bf16(
                       ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, out.shape[3], co_off, kyc,
                       ctypes.byref(e) if e is not None else None, stream_ptr()),
                   detail=f"b{b} {cin}->{cout} up2-hfold {h}x{w} g{g}",
                   nbytes=2.0 * b * (h * w * cin + 4 * h * w * cout))
    _lib.check(rc, "conv2d_up2h_bf16")
    return out


def conv_wgrad(dy_nhwc, x_nhwc, groups, kh, kw, stride, pad, dil):
    """gw[g,t,o,i] = sum_p dy[b,p,o] x[b, p*stride + t*dil - pad, i] (tap-major); groups == batch or 1; ``pad`` / ``dil``
    are ints or per-axis (h, w) pairs.
    dy_nhwc [B,OH,OW,Cout_pad] bf16, x_nhwc [B,H,W,Cin_pad] bf16 -> [groups, kh*kw, Cout_pad, Cin_pad] fp32."""
    b, oh, ow, cout = dy_nhwc.shape
    _, h, w, cin = x_nhwc.shape
    gw = torch.empty((groups, kh * kw, cout, cin), dtype=torch.float32, device=dy_nhwc.device)
    (ph, pw), (dh, dw) = (pad if isinstance(pad, tuple) else (pad, pad)), (dil if isinstance(dil, tuple) else (dil, dil))
    with _lib.device_guard(dy_nhwc.device):
        rc = _lib.load().vsp_conv2d_wgrad_bf16(ptr(dy_nhwc), ptr(x_nhwc), ptr(gw), b, groups, h, w, cin, cout,
                                               oh, ow, kh, kw, stride, ph, pw, dh, dw, stream_ptr())
    _lib.check(rc, "conv2d_wgrad_bf16")
    return gw


def conv_transposed(x_nhwc, wq, cout, kh, kw, stride, pad=(0, 0), dil=(1, 1), out_pad=(0, 0), out=None, epi=None,
                    out_nhwc=False):
    """General transposed convolution (``F.conv_transpose2d`` semantics) without a zero-stuffed intermediate:
    out[y, x] = sum over taps (i, j) and input pixels (a, b) with y = a*stride - pad + i*dil of x[a, b] * w[i, j].
    One gather launch per output parity class (stride^2 classes); class (pa, pb) owns the output pixels
    (A*stride + pa, B*stride + pb) and reads x[A + (pa + pad - i*dil)/stride] for the taps where that is an integer.
    x_nhwc [B,H,W,K] bf16, wq [G,kh*kw,n_pad,K] (n = output channels) -> [B,cout,FH,FW] fp32 NCHW (or NHWC bf16) with
    FH = (H-1)*stride - 2*pad + dil*(kh-1) + out_pad + 1."""
    b, h, w, _ = x_nhwc.shape
    s = stride
    (ph, pw), (dh, dw), (oph, opw) = pad, dil, out_pad
    fh = (h - 1) * s - 2 * ph + dh * (kh - 1) + oph + 1
    fw = (w - 1) * s - 2 * pw + dw * (kw - 1) + opw + 1
    if out is None:
        out = (torch.zeros((b, max(fh, 0), max(fw, 0), _round_up(cout, 8)), dtype=torch.bfloat16, device=x_nhwc.device)
               if out_nhwc else torch.empty((b, cout, max(fh, 0), max(fw, 0)), dtype=torch.float32, device=x_nhwc.device))
    if out.numel() == 0:
        return out
    classes = []
    for pa in range(s):
        for pb in range(s):
            taps = [(i, j) for i in range(kh) for j in range(kw)
                    if (pa + ph - i * dh) % s == 0 and (pb + pw - j * dw) % s == 0]
            rows, cols = (fh - pa + s - 1) // s, (fw - pb + s - 1) // s
            if rows > 0 and cols > 0:
                classes.append((pa, pb, taps, rows, cols))
    if any(not c[2] for c in classes):
        out.zero_()                                  # a class without taps (e.g. 1x1 stride 2) stays zero
    for pa, pb, taps, rows, cols in classes:
        if taps:
            conv_gather(x_nhwc, wq, cout, [i * kw + j for i, j in taps], [(pa + ph - i * dh) // s for i, j in taps],
                        [(pb + pw - j * dw) // s for i, j in taps], 1, (rows, cols), full_hw=(fh, fw), os_=s,
                        oo=(pa, pb), out=out, epi=epi, out_nhwc=out_nhwc)
    return out


# ----------------------------------------------------------------------------------------------
# the modulated convolution (NCHW fp32 boundary): modulation in the producer, demodulation in the epilogue
# ----------------------------------------------------------------------------------------------
#
#   y[b] = d[b,o] * conv(x[b] * s[b,i], wscale * W)         d[b,o] = rsqrt(wscale^2 * sum_i s[b,i]^2 * sum_t W[o,i,t]^2 + eps)
#
# is the reference's un-fused branch (models/RestoreNet.py:481-508) and algebraically its fused one (:510-553).  Nothing
# per-sample is ever materialised: the style scales the ACTIVATION inside the NCHW fp32 -> NHWC bf16 conversion the
# tensor-core path needs anyway (`ModulateInput`), the GEMM runs on ONE shared bf16 weight tensor (`SharedWeightConv`:
# samples may share an M tile, the weight gradient is one batch-summed GEMM), demodulation is a row scale in the epilogue,
# and `demod_coefs` is plain (twice differentiable) tensor algebra on [B,Cin] x [Cin,Cout], so autograd supplies the
# demodulation terms of dW and ds.  The style gradient needs no per-sample weight gradient either:
# ds[b,i] = sum_p x[b,i,p] * dxs[b,i,p], a reduction fused into the conversion of the incoming gradient.

def _taps(k):
    return [(i, j) for i in range(k) for j in range(k)]


def _demod_coefs_algebra(weight, s, wscale, eps=1e-8):
    """d[b,o] as plain tensor algebra — differentiable to any order (the double-backward route)."""
    wsq = weight.reshape(weight.shape[-4], weight.shape[-3], -1).square().sum(-1)          # [Cout, Cin]
    return torch.rsqrt((wscale * wscale) * (s.square() @ wsq.t()) + eps)


def _weight_derived(cache, weight, tag, build):
    """A tensor derived from ``weight`` alone (sum_t W^2, packed bf16 forms), rebuilt only when the parameter changes
    (``_version`` moves on every in-place update, i.e. every optimizer step; a new storage moves ``data_ptr``).
    ``cache`` is a dict OWNED BY THE LAYER MODULE (it dies with the parameter, so a recycled address can never alias a stale
    entry); None = no caching.  While a CUDA graph is being captured the tensor is always rebuilt, so the packing kernels
    are part of the graph and a replay that follows an in-graph optimizer step never sees weights packed at capture time."""
    if cache is None or torch.cuda.is_current_stream_capturing():
        return build()
    ver = (weight.data_ptr(), weight._version, tuple(weight.shape))
    hit = cache.get(tag)
    if hit is not None and hit[0] == ver:
        return hit[1]
    val = build()
    cache[tag] = (ver, val)
    return val


class DemodCoefs(Function):
    """d[b,o] = rsqrt(wscale^2 * sum_i s[b,i]^2 * sum_t W[o,i,t]^2 + eps) (models/RestoreNet.py:513-516): one launch over the
    cached sum_t W^2 instead of five library passes over the 9-tap weight.  First-order gradients in closed form; under
    ``create_graph`` the backward re-derives them from the differentiable algebra."""

    @staticmethod
    def forward(ctx, weight, s, wscale, eps, cache):
        w3 = weight.detach()
        wsq = _weight_derived(cache, w3, "wsq", lambda: weight_sumsq(w3.contiguous()))
        d = demod_from_wsq(s.detach(), wsq, wscale, eps)
        ctx.save_for_backward(weight, s, d, wsq)
        ctx.cfg = (wscale, eps)
        return d

    @staticmethod
    def backward(ctx, gd):
        weight, s, d, wsq = ctx.saved_tensors
        wscale, eps = ctx.cfg
        if torch.is_grad_enabled():
            with torch.enable_grad():
                dd = _demod_coefs_algebra(weight, s, wscale, eps)
                wanted = [t for t, n in zip((weight, s), ctx.needs_input_grad[:2]) if n]
                grads = list(torch.autograd.grad(dd, wanted, gd, create_graph=True, allow_unused=True))
            return tuple(grads.pop(0) if n else None for n in ctx.needs_input_grad[:2]) + (None, None, None)
        gq = (-0.5 * wscale * wscale) * gd * d * d * d                    # dL/d(sum_i s^2 wsq) [B, Cout]
        gw = gs = None
        if ctx.needs_input_grad[0]:
            gw = (2.0 * (gq.t() @ s.square())).unsqueeze(-1).unsqueeze(-1) * weight        # [Cout, Cin, 1, 1] * W
        if ctx.needs_input_grad[1]:
            gs = 2.0 * s * (gq @ wsq)
        return gw, gs, None, None, None


def demod_coefs(weight, s, wscale, eps=1e-8, cache=None):
    """d[b,o] (models/RestoreNet.py:513-516) from the style-independent sum_t W^2; differentiable in W and s."""
    if weight.device.type != "cuda":
        return _demod_coefs_algebra(weight, s, wscale, eps)
    return DemodCoefs.apply(weight, s, wscale, eps, cache)


class ModulateInput(Function):
    """xs = bf16(x * s[b,i]) in NHWC (s None: plain layout conversion).  Backward turns the NHWC bf16 gradient of xs
    into dx = dxs * s (NCHW fp32) and ds = sum_p x * dxs in one pass."""

    @staticmethod
    def forward(ctx, x, s):
        if x.device.type != "cuda":
            raise RuntimeError("modulated_conv2d: input must be a CUDA tensor (vspbfr_b200 has no CPU path)")
        ctx.save_for_backward(x, s)
        return nchw_to_nhwc_bf16(x, scale_nc=s.contiguous() if s is not None else None)

    @staticmethod
    def backward(ctx, dxs):
        x, s = ctx.saved_tensors
        c = x.shape[1]
        if torch.is_grad_enabled():                      # double backward: plain tensor algebra
            g = dxs[..., :c].permute(0, 3, 1, 2).float()
            if s is None:
                return g, None
            return g * s[:, :, None, None], (g * x).sum((2, 3)) if ctx.needs_input_grad[1] else None
        if s is None:
            return nhwc_bf16_to_nchw(dxs, c), None
        if ctx.needs_input_grad[1]:
            dx, ds = nhwc_bf16_to_nchw(dxs, c, scale_nc=s, dot_with=x)
            return (dx if ctx.needs_input_grad[0] else None), ds
        return nhwc_bf16_to_nchw(dxs, c, scale_nc=s), None


def _shared_conv_reference(xs, weight, d, mode, dilation, cin):
    """SharedWeightConv out of conv2d_gradfix ops (differentiable to any order) — the double-backward route."""
    from . import conv2d_gradfix

    cout, _, k, _ = weight.shape
    xf = xs[..., :cin].permute(0, 3, 1, 2).float()
    w = weight * (1.0 / math.sqrt(cin * k * k))
    if mode == "up":
        z = conv2d_gradfix.conv_transpose2d(xf, w.transpose(0, 1), stride=2, padding=0, dilation=dilation)
    elif mode == "down":
        z = conv2d_gradfix.conv2d(xf, w, stride=2, padding=0, dilation=dilation)
    else:
        z = conv2d_gradfix.conv2d(xf, w, padding=((k - 1) * dilation) // 2, dilation=dilation)
    return z if d is None else z * d[:, :, None, None]


class SharedWeightConv(Function):
    """y = d[b,o] * conv(xs, wscale * W) on tcgen05: xs [B,H,W,Cin_pad] bf16 NHWC, W [Cout,Cin,k,k] fp32 (shared by the
    batch), d [B,Cout] fp32 or None -> y [B,Cout,OH,OW] fp32 NCHW.
    mode: "same" (stride 1, padding (k-1)*dil/2), "down" (stride 2, padding 0), "up" (transposed, stride 2, padding 0).
    Backward: dz = d * dy (fused into the NHWC conversion of dy together with sum_p dy*y, which gives dL/dd = that / d);
    dxs = the adjoint convolution of dz (NHWC bf16 out); dW = wscale * the batch-summed pixel-K GEMM of dz and xs."""

    @staticmethod
    def forward(ctx, xs, weight, d, mode, dilation, cache=None):
        cout, cin, k, _ = weight.shape
        wscale = 1.0 / math.sqrt(cin * k * k)
        wd = weight.detach()
        wq = _weight_derived(cache, wd, "wq", lambda: pack_weights(wd, wscale=wscale)[0])
        ctx.cache = cache
        if wq.shape[3] != xs.shape[3]:
            raise RuntimeError(f"modulated_conv2d: activation has {xs.shape[3]} channels, weight expects {wq.shape[3]}")
        epi = make_epilogue(row_scale=d.contiguous()) if d is not None else None
        if mode == "up":
            out = conv_transposed(xs, wq, cout, k, k, 2, dil=(dilation, dilation), epi=epi)
        elif mode == "down":
            out = conv_fprop(xs, wq, cout, k, k, 2, 0, dilation, epi=epi)
        else:
            out = conv_fprop(xs, wq, cout, k, k, 1, ((k - 1) * dilation) // 2, dilation, epi=epi)
        ctx.save_for_backward(xs, weight, d, out if d is not None else None)
        ctx.cfg = (mode, dilation, wscale)
        return out

    @staticmethod
    def backward(ctx, dy):
        xs, weight, d, y = ctx.saved_tensors
        mode, dilation, wscale = ctx.cfg
        cout, cin, k, _ = weight.shape
        need_x, need_w, need_d = ctx.needs_input_grad[:3]
        if torch.is_grad_enabled():
            with torch.enable_grad():
                out = _shared_conv_reference(xs, weight, d, mode, dilation, cin)
                wanted = [t for t, n in zip((xs, weight, d), (need_x, need_w, need_d)) if n]
                grads = list(torch.autograd.grad(out, wanted, dy, create_graph=True, allow_unused=True))
            res = [grads.pop(0) if n else None for n in (need_x, need_w, need_d)]
            if res[0] is not None and res[0].dtype != xs.dtype:
                res[0] = res[0].to(xs.dtype)
            return (*res, None, None, None)

        b, h, w_, _ = xs.shape
        pad = ((k - 1) * dilation) // 2
        dxs = dw = dd = None
        if d is not None:
            dzq, dot = nchw_to_nhwc_bf16(dy, scale_nc=d, dot_with=y)       # dz = d * dy ; dot = sum_p dy * y
            if need_d:
                dd = dot / d
        else:
            dzq = nchw_to_nhwc_bf16(dy)
        taps = _taps(k)
        if need_x:
            wd = weight.detach()
            wq_t = _weight_derived(ctx.cache, wd, "wq_t", lambda: pack_weights(wd, wscale=wscale, transpose=True)[0])   # n = Cin, k = Cout
            cpad = xs.shape[3]
            if mode == "same":
                dxs = conv_gather(dzq, wq_t, cin, [i * k + j for i, j in taps], [pad - i * dilation for i, j in taps],
                                  [pad - j * dilation for i, j in taps], 1, (h, w_), out_nhwc=True,
                                  out=_nhwc_out(b, h, w_, cin, cpad, xs.device))
            elif mode == "down":
                reach = dilation * (k - 1) + 1
                dxs = conv_transposed(dzq, wq_t, cin, k, k, 2, dil=(dilation, dilation), out_nhwc=True,
                                      out_pad=(h - ((dy.shape[2] - 1) * 2 + reach), w_ - ((dy.shape[3] - 1) * 2 + reach)),
                                      out=_nhwc_out(b, h, w_, cin, cpad, xs.device))
            else:
                dxs = conv_fprop(dzq, wq_t, cin, k, k, 2, 0, dilation, out_nhwc=True,
                                 out=_nhwc_out(b, h, w_, cin, cpad, xs.device))
        if need_w:
            if mode == "up":
                # y[2p + t*dil] += xs[p] w[t]  =>  G[t,o,i] = sum_p xs[p,i] dz[2p + t*dil, o]: roles swapped, then transposed
                gw = conv_wgrad(xs, dzq, 1, k, k, 2, 0, dilation).transpose(2, 3)
            elif mode == "down":
                gw = conv_wgrad(dzq, xs, 1, k, k, 2, 0, dilation)
            else:
                gw = conv_wgrad(dzq, xs, 1, k, k, 1, pad, dilation)
            dw = (gw[0, :, :cout, :cin].permute(1, 2, 0) * wscale).reshape(weight.shape)
        return dxs, dw, dd, None, None, None


def _nhwc_out(b, h, w, c, c_pad, device):
    """NHWC bf16 gradient buffer; padding channels (c..c_pad) must read as zero."""
    if c_pad == c:
        return torch.empty((b, h, w, c_pad), dtype=torch.bfloat16, device=device)
    return torch.zeros((b, h, w, c_pad), dtype=torch.bfloat16, device=device)


def modulate_input(x, s=None):
    """x [B,C,H,W] fp32 (* s [B,C]) -> NHWC bf16 activation shared by every convolution that consumes x with this style
    (the four dilated branches of a SMART layer convert their common input once)."""
    return ModulateInput.apply(x, s)


def modulated_conv2d(x, weight, s, demodulate=True, mode="same", dilation=1, xs=None, eps=1e-8, cache=None):
    """x [B,Cin,H,W] fp32, weight [1,Cout,Cin,k,k], s [B,Cin] (already through ``modulation``) -> [B,Cout,OH,OW] fp32.
    ``xs``: the result of ``modulate_input(x, s)`` when the caller shares it between several convolutions.
    ``cache``: a dict owned by the calling layer for tensors derived from ``weight`` alone (sum_t W^2, packed bf16 weights),
    revalidated against the parameter's version on every call."""
    _, cout, cin, k, _ = weight.shape
    if xs is None:
        if x.shape[1] != cin or s.shape != (x.shape[0], cin):
            raise RuntimeError(f"modulated_conv2d: shape mismatch x{tuple(x.shape)} w{tuple(weight.shape)} s{tuple(s.shape)}")
        xs = modulate_input(x, s)
    w4 = weight.reshape(cout, cin, k, k)
    d = demod_coefs(w4, s, 1.0 / math.sqrt(cin * k * k), eps, cache) if demodulate else None
    return SharedWeightConv.apply(xs, w4, d, mode, dilation, cache)
