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
import weakref

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

_nhwc_memo = [None]     # (weakref to the source tensor, its _version, c_pad, converted tensor) of the latest plain conversion


def nchw_to_nhwc_bf16(x, scale_nc=None, c_pad=None):
    """[N,C,H,W] fp32 -> [N,H,W,c_pad] bf16 (optionally * scale_nc[n,c]); extra channels are zero.

    The latest un-scaled conversion is memoised on the identity and version of its source: the four dilated branches of a
    SMART layer (models/RestoreNet.py:229-233) convert the same input in forward and again in backward — a quarter of the
    ~500 layout conversions of a training iteration."""
    n, c, h, w = x.shape
    c_pad = c_pad or _round_up(c, 8)
    memo = _nhwc_memo[0]
    if (scale_nc is None and memo is not None and memo[0]() is x and memo[1] == x._version and memo[2] == c_pad
            and not torch.cuda.is_current_stream_capturing()):
        return memo[3]
    xc = x.contiguous()
    y = torch.empty((n, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    if y.numel():
        with torch.cuda.device(x.device):
            rc = _lib.load().vsp_nchw_f32_to_nhwc_bf16(ptr(xc), ptr(scale_nc), ptr(y), n, c, h * w, c_pad, stream_ptr())
        _lib.check(rc, "nchw_f32_to_nhwc_bf16")
    if scale_nc is None and not torch.cuda.is_current_stream_capturing():
        try:
            _nhwc_memo[0] = (weakref.ref(x), x._version, c_pad, y)
        except TypeError:
            _nhwc_memo[0] = None
    return y


def nhwc_bf16_to_nchw(x, c=None):
    """[N,H,W,c_pad] bf16 -> [N,c,H,W] fp32."""
    n, h, w, c_pad = x.shape
    c = c or c_pad
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    if y.numel():
        with torch.cuda.device(x.device):
            rc = _lib.load().vsp_nhwc_bf16_to_nchw_f32(ptr(x.contiguous()), ptr(y), n, c, h * w, c_pad, stream_ptr())
        _lib.check(rc, "nhwc_bf16_to_nchw_f32")
    return y


def nchw_to_bf16(x, scale_nc=None):
    """[N,C,H,W] fp32 -> same layout bf16, optionally * scale_nc[n,c]."""
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty_like(x, dtype=torch.bfloat16)
    if y.numel():
        with torch.cuda.device(x.device):
            rc = _lib.load().vsp_nchw_f32_to_bf16(ptr(x), ptr(scale_nc), ptr(y), n * c, h * w, stream_ptr())
        _lib.check(rc, "nchw_f32_to_bf16")
    return y


def weight_sumsq(weight):
    """wsq[o,i] = sum over taps of W^2 — the style-independent half of the demodulation sum."""
    cout, cin, kh, kw = weight.shape
    weight = weight.contiguous()
    wsq = torch.empty((cout, cin), dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
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
    with torch.cuda.device(weight.device):
        rc = _lib.load().vsp_modulate_weights_bf16(ptr(weight), ptr(style), ptr(demod), ptr(wq), g, cout, cin, taps,
                                                   wscale, eps, int(transpose), int(fold_demod), n_pad, k_pad,
                                                   ptr(wsq), stream_ptr())
    _lib.check(rc, "modulate_weights_bf16")
    return wq, demod


def make_epilogue(row_scale=None, noise=None, noise_weight=0.0, noise_weight_dev=None, bias=None, act=0, alpha=0.2,
                  scale=1.0, residual=None, residual2=None, pre_bias=None, pre_act=0):
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
    return e, (row_scale, noise, noise_weight_dev, bias, residual, residual2, pre_bias)


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


def conv_fprop(x_nhwc, wq, cout, kh, kw, stride=1, pad=0, dil=1, epi=None, out=None, out_nhwc=False, co_off=0):
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
    with torch.cuda.device(x_nhwc.device):
        rc = _prof("conv_fprop", 2.0 * b * oh * ow * cout * cin * kh * kw, lambda: _lib.load().vsp_conv2d_fprop_bf16(
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
    with torch.cuda.device(x_nhwc.device):
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
    with torch.cuda.device(x_nhwc.device):
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
    with torch.cuda.device(x_nhwc.device):
        rc = _lib.load().vsp_scale_nhwc_bf16(ptr(x_nhwc), ptr(s), ptr(y), b, h * w, c, stream_ptr())
    _lib.check(rc, "scale_nhwc_bf16")
    return y


def demod_from_wsq(s, wsq, wscale, eps=1e-8):
    """d[b,o] = rsqrt(wscale^2 * sum_i s[b,i]^2 * wsq[o,i] + eps) (models/RestoreNet.py:513-516 with cached sum_t W^2)."""
    b, cin = s.shape
    cout = wsq.shape[0]
    d = torch.empty((b, cout), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        rc = _lib.load().vsp_modulate_weights_bf16(ptr(wsq), ptr(s.contiguous()), ptr(d), None, b, cout, cin, 1, wscale, eps,
                                                   0, 0, cout, cin, ptr(wsq), stream_ptr())
    _lib.check(rc, "modulate_weights_bf16(demod)")
    return d


def conv_branches(x_nhwc, wq, cout, dils, epi=None, out=None, out_nhwc=True, co_off=0):
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
    with torch.cuda.device(x_nhwc.device):
        rc = _prof("conv_branches", 2.0 * b * h * w * cout * cin * 9, lambda: _lib.load().vsp_conv2d_branches_bf16(
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
    with torch.cuda.device(x_nhwc.device):
        # algorithmic FLOPs of the layer it replaces (the transposed conv); the dense form executes 4x as many
        rc = _prof("conv_up2_fused", 2.0 * b * h * w * cout * cin * 9,
                   lambda: _lib.load().vsp_conv2d_up2_fused_bf16(
                       ptr(x_nhwc), ptr(wq), ptr(out), b, g, h, w, cin, cout, out.shape[3], co_off,
                       ctypes.byref(e) if e is not None else None, stream_ptr()),
                   detail=f"b{b} {cin}->{cout} up2-fused {h}x{w} g{g}",
                   nbytes=2.0 * b * (h * w * cin + 4 * h * w * cout))
    _lib.check(rc, "conv2d_up2_fused_bf16")
    return out


def conv_wgrad(dy_nhwc, x_nhwc, groups, kh, kw, stride, pad, dil):
    """gw[g,t,o,i] = sum_p dy[b,p,o] x[b, p*stride + t*dil - pad, i] (tap-major); groups == batch or 1.
    dy_nhwc [B,OH,OW,Cout_pad] bf16, x_nhwc [B,H,W,Cin_pad] bf16 -> [groups, kh*kw, Cout_pad, Cin_pad] fp32."""
    b, oh, ow, cout = dy_nhwc.shape
    _, h, w, cin = x_nhwc.shape
    gw = torch.empty((groups, kh * kw, cout, cin), dtype=torch.float32, device=dy_nhwc.device)
    with torch.cuda.device(dy_nhwc.device):
        rc = _lib.load().vsp_conv2d_wgrad_bf16(ptr(dy_nhwc), ptr(x_nhwc), ptr(gw), b, groups, h, w, cin, cout,
                                               oh, ow, kh, kw, stride, pad, dil, stream_ptr())
    _lib.check(rc, "conv2d_wgrad_bf16")
    return gw


def weight_style_grad(gw, weight, style, demod, wscale, want_dw=True, want_ds=True):
    b, taps, cout, cin = gw.shape
    dw = torch.empty((cout, cin, taps), dtype=torch.float32, device=gw.device) if want_dw else None
    ds = torch.empty((b, cin), dtype=torch.float32, device=gw.device) if want_ds else None
    with torch.cuda.device(gw.device):
        rc = _lib.load().vsp_modconv_weight_style_grad(ptr(gw), ptr(weight.contiguous()), ptr(style.contiguous()),
                                                       ptr(demod), ptr(dw), ptr(ds), b, cout, cin, taps, wscale,
                                                       stream_ptr())
    _lib.check(rc, "modconv_weight_style_grad")
    return dw, ds


# ----------------------------------------------------------------------------------------------
# plain convolution primitives for conv2d_gradfix (NCHW fp32 boundary)
# ----------------------------------------------------------------------------------------------

def plain_conv_supported(input, weight_shape, stride, padding, dilation, groups, transpose):
    """The tcgen05 path covers square stride-1/2 convs with <= 16 taps; groups == 1, or the
    reference's grouped form (input [1, B*Cin, H, W], groups == B)."""
    if input.dtype != torch.float32 or input.ndim != 4:
        return False
    if transpose:
        return False  # transposed fprop of conv2d_gradfix goes through ATen for now
    cout_total, cin_g, kh, kw = weight_shape
    if kh * kw > 16 or stride[0] != stride[1] or stride[0] not in (1, 2):
        return False
    if padding[0] != padding[1] or dilation[0] != dilation[1]:
        return False
    if groups != 1 and input.shape[0] != 1:
        return False
    if cin_g < 8 or input.shape[2] < 1:
        return False
    return True


def plain_conv_fprop(input, weight, bias, stride, padding, dilation, groups):
    n, c_total, h, w = input.shape
    cout_total, cin, kh, kw = weight.shape
    cout = cout_total // groups
    if groups == 1:
        x = nchw_to_nhwc_bf16(input)
        wq, _ = pack_weights(weight)
        epi = make_epilogue(bias=bias) if bias is not None else None
        return conv_fprop(x, wq, cout, kh, kw, stride[0], padding[0], dilation[0], epi=epi)
    # grouped form of the modulated conv: [1, B*Cin, H, W] x [B*Cout, Cin, k, k]
    x = nchw_to_nhwc_bf16(input.reshape(groups, cin, h, w))
    wq = weight.reshape(groups, cout, cin, kh * kw).permute(0, 3, 1, 2)
    k_pad = _round_up(cin, 8)
    wq_p = torch.zeros((groups, kh * kw, cout, k_pad), dtype=torch.bfloat16, device=input.device)
    wq_p[..., :cin] = wq.to(torch.bfloat16)
    out = conv_fprop(x, wq_p, cout, kh, kw, stride[0], padding[0], dilation[0])
    out = out.reshape(1, groups * cout, out.shape[2], out.shape[3])
    if bias is not None:
        out = out + bias.reshape(1, -1, 1, 1)
    return out


def plain_conv_dgrad(*a, **k):  # pragma: no cover - routed to ATen by plain_conv_supported
    raise RuntimeError("plain_conv_dgrad: not routed to tcgen05")


def plain_conv_wgrad(grad_output, input, weight_shape, stride, padding, dilation, groups):
    """Returns None when the wgrad kernel does not cover the shape (caller uses ATen)."""
    return None


# ----------------------------------------------------------------------------------------------
# the modulated convolution (NCHW fp32 boundary, fused tcgen05 path)
# ----------------------------------------------------------------------------------------------

def _modconv_reference_form(x, weight, s, demodulate, mode, dilation, eps=1e-8):
    """The reference's own formulation (models/RestoreNet.py:510-553: materialised per-sample
    weights + grouped conv through conv2d_gradfix).  Differentiable to any order; used for the
    double-backward path (create_graph=True) of ModulatedConv2dFunction."""
    from . import conv2d_gradfix

    b, cin, h, w_ = x.shape
    _, cout, _, k, _ = weight.shape
    wscale = 1.0 / math.sqrt(cin * k * k)
    wm = wscale * weight * s.reshape(b, 1, cin, 1, 1)
    if demodulate:
        d = torch.rsqrt(wm.pow(2).sum([2, 3, 4]) + eps)
        wm = wm * d.reshape(b, cout, 1, 1, 1)
    xin = x.reshape(1, b * cin, h, w_)
    if mode == "up":
        wt = wm.transpose(1, 2).reshape(b * cin, cout, k, k)
        out = conv2d_gradfix.conv_transpose2d(xin, wt, padding=0, stride=2, groups=b, dilation=dilation)
    elif mode == "down":
        out = conv2d_gradfix.conv2d(xin, wm.reshape(b * cout, cin, k, k), padding=0, stride=2, groups=b,
                                    dilation=dilation)
    else:
        out = conv2d_gradfix.conv2d(xin, wm.reshape(b * cout, cin, k, k), padding=((k - 1) * dilation) // 2,
                                    groups=b, dilation=dilation)
    return out.reshape(b, cout, out.shape[2], out.shape[3])


class ModulatedConv2dFunction(Function):
    """y[b] = demod[b] * conv(x[b], wscale * W * s[b]) on tcgen05 (bf16 operands, fp32 accumulate).

    mode: "same" (stride 1, padding (k-1)*dil/2), "down" (stride 2, padding 0 — the caller blurs
    first), "up" (transposed stride 2, padding 0 — the caller blurs afterwards).
    First-order backward runs on the same kernels (dgrad = gather conv with transposed weights,
    wgrad = pixel-K GEMM, style/weight gradients incl. the demodulation term); when autograd asks
    for a differentiable backward (create_graph=True) it switches to the reference formulation
    through conv2d_gradfix, which is closed under differentiation.
    """

    @staticmethod
    def forward(ctx, x, weight, s, demodulate, mode, dilation):
        b, cin, h, w_ = x.shape
        _, cout, cin_w, k, _ = weight.shape
        if cin_w != cin or s.shape != (b, cin):
            raise RuntimeError(f"modulated_conv2d: shape mismatch x{tuple(x.shape)} w{tuple(weight.shape)} s{tuple(s.shape)}")
        if x.device.type != "cuda":
            raise RuntimeError("modulated_conv2d: input must be a CUDA tensor (vspbfr_b200 has no CPU path)")
        wscale = 1.0 / math.sqrt(cin * k * k)
        w4 = weight.reshape(cout, cin, k, k)
        xq = nchw_to_nhwc_bf16(x)
        # demodulation from sum_t W^2 (one well-parallelised pass over the weights + a [B,Cout] pass over it) instead of the
        # warp-per-(b,o) walk over all of W (52 us per call at 512x512x9, 165 calls per training iteration)
        wq, d = pack_weights(w4, s, wscale=wscale, want_demod=demodulate, wsq=weight_sumsq(w4) if demodulate else None)
        epi = make_epilogue(row_scale=d) if demodulate else None
        if mode == "up":
            if dilation != 1:
                raise RuntimeError("modulated_conv2d: dilated transposed convolution is not supported")
            out = conv_transpose_s2(xq, wq, cout, k, k, epi=epi)
        elif mode == "down":
            out = conv_fprop(xq, wq, cout, k, k, 2, 0, dilation, epi=epi)
        else:
            out = conv_fprop(xq, wq, cout, k, k, 1, ((k - 1) * dilation) // 2, dilation, epi=epi)
        ctx.save_for_backward(x, weight, s, d if demodulate else None)
        ctx.cfg = (demodulate, mode, dilation, wscale)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight, s, d = ctx.saved_tensors
        demodulate, mode, dilation, wscale = ctx.cfg
        if torch.is_grad_enabled():
            # differentiable backward requested (double backward): closed-set formulation
            with torch.enable_grad():
                y = _modconv_reference_form(x, weight, s, demodulate, mode, dilation)
                need = [t for t, n in zip((x, weight, s), ctx.needs_input_grad[:3]) if n]
                grads = list(torch.autograd.grad(y, need, dy, create_graph=True, allow_unused=True))
            out = []
            for n in ctx.needs_input_grad[:3]:
                out.append(grads.pop(0) if n else None)
            return (*out, None, None, None)

        b, cin, h, w_ = x.shape
        _, cout, _, k, _ = weight.shape
        w4 = weight.reshape(cout, cin, k, k)
        pad = ((k - 1) * dilation) // 2
        dzq = nchw_to_nhwc_bf16(dy, scale_nc=d)           # dz = demod * dy, channels-last bf16
        taps = [(i, j) for i in range(k) for j in range(k)]
        dx = dw = ds = None
        if ctx.needs_input_grad[0]:
            wq_t, _ = pack_weights(w4, s, wscale=wscale, transpose=True)
            if mode == "same":
                dx = conv_gather(dzq, wq_t, cin, [i * k + j for i, j in taps], [pad - i * dilation for i, j in taps],
                                 [pad - j * dilation for i, j in taps], 1, (h, w_))
            elif mode == "down":
                dx = conv_transpose_s2(dzq, wq_t, cin, k, k)
                if dx.shape[2] != h or dx.shape[3] != w_:   # even-sized inputs lose their last row/col to stride 2
                    full = torch.zeros_like(x)
                    full[:, :, :dx.shape[2], :dx.shape[3]] = dx[:, :, :h, :w_]
                    dx = full
            else:
                dx = conv_fprop(dzq, wq_t, cin, k, k, 2, 0, 1)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            xq = nchw_to_nhwc_bf16(x)
            if mode == "up":
                # y[2p+t] += x[p] w[t]  =>  G[t,o,i] = sum_p x[p,i] dz[2p+t,o]: roles swapped, then transposed
                gw = conv_wgrad(xq, dzq, b, k, k, 2, 0, 1).transpose(2, 3).contiguous()
            elif mode == "down":
                gw = conv_wgrad(dzq, xq, b, k, k, 2, 0, dilation)
            else:
                gw = conv_wgrad(dzq, xq, b, k, k, 1, pad, dilation)
            gw = gw[:, :, :cout, :cin]
            if gw.shape[2] != cout or not gw.is_contiguous():
                gw = gw.contiguous()
            dw, ds = weight_style_grad(gw, w4, s, d, wscale, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
            if dw is not None:
                dw = dw.reshape(weight.shape)
        return dx, dw, ds, None, None, None


def modulated_conv2d(x, weight, s, demodulate=True, mode="same", dilation=1):
    """x [B,Cin,H,W] fp32, weight [1,Cout,Cin,k,k], s [B,Cin] (already through ``modulation``)."""
    return ModulatedConv2dFunction.apply(x, weight, s, demodulate, mode, dilation)
