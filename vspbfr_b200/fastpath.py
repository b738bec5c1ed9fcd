"""Fused channels-last bf16 inference pipeline over the layer modules of ``layers.py``.

Same arithmetic as the reference forwards (models/RestoreNet.py:968-1046,
e4e/models/stylegan2/model.py:475-552) with the memory traffic removed
(SURVEY.md §8 f-1): activations stay NHWC bf16 between layers; per-sample style modulation is a
weight prologue; demodulation, NoiseInjection, bias + leaky-ReLU, the SMART double activation, the
``out + feat + sty_de_feat`` skip fusion and the ToRGB bias + skip add run in conv / blur epilogues;
the four dilated SMART branches write channel slices of one buffer (no ``torch.cat``).  The RGB skip
chain stays fp32 NCHW (3 channels: negligible traffic, keeps the image path at full precision).

No autograd here — use the module ``forward`` methods for training.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref

import torch
from torch.nn import functional as F

from . import _lib
from ._lib import ptr, stream_ptr
from .layers import LargeConvLayer, SMART_layer, StyledConv, ToRGB
from .op import modconv as mc
from .op.upfirdn2d import upfirdn2d_raw

_cache: dict = {}
# ModulatedConv2d(upsample=True) layers with Cin up to this run as the fused up-convolution (4x the transposed
# conv's FLOPs, but no (2H+1)^2 intermediate and no blur pass): a win wherever the layer is bandwidth-bound
_UP_FUSED_MAX_CIN = int(os.environ.get("VSP_UP_FUSED_MAX_CIN", "256"))
_UP2H = os.environ.get("VSP_UP2H", "1") != "0"          # half-composed up-convolution on the wide levels (conv_up2h_sm100.cu)
# Activations with at most this many pixels per sample run in the input-modulated form (x * s, shared cached
# weights, demodulation in the epilogue: models/RestoreNet.py:481-508): below it the per-sample weight prologue
# costs more than scaling the activation, and shared weights let one 128-row tile stack several samples
_LOWRES_PIXELS = int(os.environ.get("VSP_LOWRES_PIXELS", "1024"))
# Low-resolution up-convolutions with at least this many input pixels (the 512-channel 32x32 layers) run as transposed conv
# (1x the layer's FLOPs, one class-mode launch) + blur instead of the composite dense form (4x the FLOPs): 175 + 72 us vs
# 359 us per 32 faces (with both skip residuals 171 + 109 vs 380), and 0.9 TFLOP less per micro-batch under the power cap.
# The 16x16 layers stay dense: their transposed conv is not eligible for the one-launch form (122 + 39 us vs 133 us), and
# with them split the full-size parity test sits at 0.99e-2 of the range (0.87e-2 with this default, 0.92e-2 all dense).
_LOWRES_UP_SPLIT_PIXELS = int(os.environ.get("VSP_LOWRES_UP_SPLIT_PIXELS", "1024"))
# SMART layers up to this width run their four dilated branches as one launch of the generic kernel
_BRANCH_MAX_W = int(os.environ.get("VSP_BRANCH_MAX_W", "64"))
_SEP_BLUR = os.environ.get("VSP_NO_SEP_BLUR") is None
# (measured: 64->16 x4 @512^2 1250 us merged vs 4 x 295 us separate, 128->32 x4 @256^2 640 vs 4 x 165 us — the fold kernel
# is bound by its per-row epilogue round trip, not by re-reading x, so the merged form is off by default)
_BRANCH_FOLD = os.environ.get("VSP_BRANCH_FOLD") is not None


def _cached(owner, tag, tensors, build):
    """Memoise a derived tensor per module, invalidated when any source tensor is modified in place
    or re-assigned (weights are static during inference, so this runs once)."""
    key = (id(owner), tag)
    ver = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
    hit = _cache.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is owner:     # id() may be recycled: check the owner itself
        return hit[1]
    val = build()
    _cache[key] = (ver, val, weakref.ref(owner))
    return val


def clear_cache():
    _cache.clear()
    _bank_cache.clear()
    _bank_clear()


def _linear(lin, x):
    """EqualLinear without activation (the per-layer ``modulation``), fp32."""
    w = _cached(lin, "w_scaled", [lin.weight], lambda: (lin.weight.detach() * lin.scale).contiguous())
    b = _cached(lin, "b_scaled", [lin.bias], lambda: (lin.bias.detach() * lin.lr_mul).contiguous())
    return F.linear(x, w, b)


class ModulationBank:
    """All style-modulation linears of one network pass as a single grouped launch (vsp_grouped_linear_f32), followed
    by every demodulation vector d[b,o] = rsqrt(scale^2 * sum_i s[b,i]^2 * wsq[o,i] + eps) as a second grouped launch
    over the squared styles (models/RestoreNet.py:510-516 with the style-independent sum_t W^2 cached).

    ``entries`` = [(EqualLinear, style_index, demod)]: problem j reads style row ``styles[:, style_index_j]`` of a
    [B, n, D] style tensor and yields s_j [B, out_dim_j]; ``demod`` is None or (owner, wsq_fn, wscale, eps), giving
    d_j [B, Cout] for ``owner``.  Descriptor tables are built once and cached on the device."""

    def __init__(self, entries):
        self.entries = [(e + (None,))[:3] for e in entries]
        self._key = None

    @staticmethod
    def _table(descs, device):
        return torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(device)

    def _build(self, device, d_style, batch):
        wsqs = [dm[1]() if dm is not None else None for _, _, dm in self.entries]
        key = tuple((lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version,
                     w.data_ptr() if w is not None else 0) for (lin, _, _), w in zip(self.entries, wsqs))
        key += (d_style, batch, str(device))
        if key == self._key:
            return
        descs = (_lib.LinearDesc * len(self.entries))()
        rows, offs, y_off, r = [], [], 0, 0
        for j, (lin, idx, _) in enumerate(self.entries):
            out_dim, in_dim = lin.weight.shape
            assert in_dim == d_style and lin.weight.is_contiguous() and lin.weight.dtype == torch.float32
            dsc = descs[j]
            dsc.w, dsc.bias = lin.weight.data_ptr(), lin.bias.data_ptr()
            dsc.x_off, dsc.y_off, dsc.x_bstride = idx * d_style, y_off, 0
            dsc.in_dim, dsc.out_dim = in_dim, out_dim
            dsc.wscale, dsc.bscale = lin.scale, lin.lr_mul
            if getattr(lin, "activation", None):         # EqualLinear(activation="fused_lrelu"): lrelu(. + b, 0.2) * sqrt(2)
                dsc.act, dsc.alpha, dsc.gain = 3, 0.2, math.sqrt(2.0)
            rows.append(r)
            offs.append((y_off, out_dim))
            r += (out_dim + 7) // 8 * 8          # a block of the kernel covers 8 rows of one problem
            y_off += batch * out_dim
        self._descs = self._table(descs, device)
        self._rows = torch.tensor(rows, dtype=torch.int32, device=device)
        self._offs, self._total, self._rows_total = offs, y_off, r
        # demodulation problems: weight = wsq [Cout, Cin], input = squared styles of the owning problem
        dem = [(j, dm, w) for j, ((_, _, dm), w) in enumerate(zip(self.entries, wsqs)) if dm is not None]
        self._dem = dem
        if dem:
            ddesc = (_lib.LinearDesc * len(dem))()
            drows, doffs, d_off, r = [], [], 0, 0
            for q, (j, (owner, _, wscale, eps), wsq) in enumerate(dem):
                cout, cin = wsq.shape
                assert cin == offs[j][1] and wsq.is_contiguous()
                dsc = ddesc[q]
                dsc.w, dsc.bias = wsq.data_ptr(), None
                dsc.x_off, dsc.x_bstride, dsc.y_off = offs[j][0], cin, d_off
                dsc.in_dim, dsc.out_dim = cin, cout
                dsc.wscale, dsc.bscale = wscale * wscale, 0.0
                drows.append(r)
                doffs.append((d_off, cout, id(owner), eps))
                r += (cout + 7) // 8 * 8
                d_off += batch * cout
            self._ddescs = self._table(ddesc, device)
            self._drows = torch.tensor(drows, dtype=torch.int32, device=device)
            self._doffs, self._dtotal, self._drows_total = doffs, d_off, r
            self._wsq_keep = wsqs
        self._key = key

    def __call__(self, styles):
        """styles [B, n, D] fp32 -> ({id(lin): s [B, out_dim]}, {id(owner): d [B, Cout]})"""
        b, n, d = styles.shape
        styles = styles.contiguous().float()
        self._build(styles.device, d, b)
        lib = _lib.load()
        y = torch.empty(self._total, dtype=torch.float32, device=styles.device)
        with _lib.device_guard(styles.device):
            rc = lib.vsp_grouped_linear_f32(ptr(self._descs), ptr(self._rows), len(self.entries), self._rows_total,
                                            ptr(styles), n * d, ptr(y), b, stream_ptr())
        _lib.check(rc, "grouped_linear_f32")
        s_out = {id(lin): y[o:o + b * od].view(b, od) for (lin, _, _), (o, od) in zip(self.entries, self._offs)}
        d_out = {}
        if self._dem:
            y2 = y * y
            dpre = torch.empty(self._dtotal, dtype=torch.float32, device=styles.device)
            with _lib.device_guard(styles.device):
                rc = lib.vsp_grouped_linear_f32(ptr(self._ddescs), ptr(self._drows), len(self._dem), self._drows_total,
                                                ptr(y2), 0, ptr(dpre), b, stream_ptr())
            _lib.check(rc, "grouped_linear_f32(demod)")
            eps = self._doffs[0][3]
            dall = torch.rsqrt(dpre + eps)
            for o, cout, owner_id, e in self._doffs:
                assert e == eps
                d_out[owner_id] = dall[o:o + b * cout].view(b, cout)
        return s_out, d_out


_mod_ctx: dict = {}
_demod_ctx: dict = {}
_bank_cache: dict = {}


def _bank_apply(bank, styles):
    s_out, d_out = bank(styles)
    _mod_ctx.update(s_out)
    _demod_ctx.update(d_out)


def _bank_clear():
    _mod_ctx.clear()
    _demod_ctx.clear()


def _clears_banks(fn):
    """The per-pass modulation / demodulation tables are module state: drop them when a forward ends OR raises midway, so a
    failed pass can never leak its styles into the next one."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        try:
            return fn(*a, **k)
        finally:
            _bank_clear()
    return wrapped


def _conv_wsq(conv):
    """Cached sum_t W^2 [Cout, Cin] of a ModulatedConv2d (style-independent part of its demodulation)."""
    return _cached(conv, "wsq", [conv.weight], lambda: mc.weight_sumsq(
        conv.weight.detach().view(conv.out_channel, conv.in_channel, conv.kernel_size, conv.kernel_size)))


def _smart_wcat(m):
    branches = list(m.ModulatedConv2ds)
    cq, k = branches[0].out_channel, branches[0].kernel_size
    return _cached(m, "wcat", [br.weight for br in branches],
                   lambda: torch.cat([br.weight.detach().view(cq, br.in_channel, k, k) for br in branches], 0).contiguous())


def _smart_wsq(m):
    return _cached(m, "wsq", [br.weight for br in m.ModulatedConv2ds], lambda: mc.weight_sumsq(_smart_wcat(m)))


def _demod_entry_conv(sc):
    """Bank entry of a StyledConv / StyledConv_down: its modulation linear + (when it demodulates) its demod problem."""
    conv = sc.conv
    return (conv, lambda: _conv_wsq(conv), conv.scale, conv.eps) if conv.demodulate else None


def _demod_entry_smart(m):
    br = m.ModulatedConv2ds[0]
    return (m, lambda: _smart_wsq(m), br.scale, br.eps) if br.demodulate else None


def _banks_for(owner, build):
    hit = _bank_cache.get(id(owner))
    if hit is None or hit[1]() is not owner:
        hit = _bank_cache[id(owner)] = (build(), weakref.ref(owner))
    return hit[0]


_TAIL_LEVELS = int(os.environ.get("VSP_TAIL_LEVELS", "2"))     # decoder levels of the restorer that run in sample groups (grouped tail)
_bank_rows = None      # slice of the batch the current calls work on (tail groups of restoration_forward), or None


def _rows(t):
    return t if (_bank_rows is None or t is None) else t[_bank_rows].contiguous()


def _modulation(lin, style):
    """s = modulation(style): from the pass's ModulationBank when one is active, else a plain linear."""
    hit = _mod_ctx.get(id(lin))
    return _rows(hit) if hit is not None else _linear(lin, style)


def _demod_pre(obj):
    """Demodulation coefficients of ``obj`` from the pass's ModulationBank (rows of the current batch slice), or None."""
    return _rows(_demod_ctx.get(id(obj)))


_sep_cache: dict = {}


def _separable_taps(kernel):
    """(fy, fx) ctypes float arrays with kernel == outer(fy, fx), or None.  Factorised once per filter tensor on the
    host (one device->host copy, cached on data_ptr/version): the model's blur filters never change."""
    key = (kernel.data_ptr(), kernel._version, tuple(kernel.shape), str(kernel.device))
    hit = _sep_cache.get(key)
    if hit is None:
        k = kernel.detach().double().cpu()
        kh, kw = k.shape
        res = False
        if kh <= 4 and kw <= 4 and float(k.abs().max()) > 0:
            j0, i0 = divmod(int(k.abs().argmax()), kw)
            fx, fy = k[j0, :].clone(), k[:, i0] / k[j0, i0]
            if float((torch.outer(fy, fx) - k).abs().max()) <= 1e-7 * float(k.abs().max()):
                res = ((ctypes.c_float * kh)(*[float(v) for v in fy]), (ctypes.c_float * kw)(*[float(v) for v in fx]))
        hit = _sep_cache[key] = (res, kernel)      # keep the tensor alive so the key cannot be recycled
    return hit[0] or None


def upfirdn_nhwc(x, kernel, up=1, down=1, pad=(0, 0), epi=None):
    """NHWC bf16 up-FIR-down with an optional fused epilogue."""
    n, h, w, c = x.shape
    kh, kw = kernel.shape
    lib = _lib.load()
    oh = lib.vsp_upfirdn2d_out_size(h, kh, up, down, pad[0], pad[1])
    ow = lib.vsp_upfirdn2d_out_size(w, kw, up, down, pad[0], pad[1])
    y = torch.empty((n, oh, ow, c), dtype=torch.bfloat16, device=x.device)
    e, keep = epi if epi is not None else (None, None)
    taps = _separable_taps(kernel) if (up == 1 and down == 1 and _SEP_BLUR) else None
    if taps is not None:
        with _lib.device_guard(x.device):
            rc = lib.vsp_blur_sep_nhwc_bf16(ptr(x), taps[0], taps[1], ptr(y), n, h, w, c, kh, kw, pad[0], pad[1], pad[0],
                                            pad[1], ctypes.byref(e) if e is not None else None, stream_ptr())
        _lib.check(rc, "blur_sep_nhwc_bf16")
        return y
    with _lib.device_guard(x.device):
        rc = lib.vsp_upfirdn2d_nhwc_bf16(ptr(x), ptr(kernel), ptr(y), n, h, w, c, kh, kw, up, up, down, down,
                                         pad[0], pad[1], pad[0], pad[1],
                                         ctypes.byref(e) if e is not None else None, stream_ptr())
    _lib.check(rc, "upfirdn2d_nhwc_bf16")
    return y


class _NoisePool:
    """N(0,1) noise images of a whole pass from ONE generator launch: layers take consecutive slices of a pool
    sized by the previous pass (shapes are static during inference)."""

    def __init__(self):
        self.buf, self.off, self.used, self.want = None, 0, 0, 0

    def begin(self, device):
        self.want = max(self.want, self.used)
        self.used = self.off = 0
        self.buf = torch.randn(self.want, device=device, dtype=torch.float32) if self.want else None

    def take(self, n, device):
        self.used += n
        if self.buf is None or self.off + n > self.buf.numel() or self.buf.device != device:
            return torch.randn(n, device=device, dtype=torch.float32)
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out


_noise_pool = _NoisePool()


def _noise_for(noise, b, h, w, device):
    """``NoiseInjection`` draws N(0,1) per call when no noise is given (models/RestoreNet.py:564-569)."""
    if noise is None:
        return _noise_pool.take(b * h * w, device).view(b, 1, h, w)
    return noise.contiguous().float()


def styled_conv(m: StyledConv, x, style, noise=None, residual=None, residual2=None):
    """StyledConv / StyledConv_down: modulated conv -> noise -> bias + lrelu (-> + residuals)."""
    conv = m.conv
    b, h, w, _ = x.shape
    cout, cin, k = conv.out_channel, conv.in_channel, conv.kernel_size
    s = _modulation(conv.modulation, style)
    w4 = conv.weight.detach().view(cout, cin, k, k)
    wsq = _conv_wsq(conv) if conv.demodulate else None
    d_pre = _demod_pre(conv) if conv.demodulate else None      # from the pass's ModulationBank, if any
    act = dict(bias=m.activate.bias.detach(), act=3, alpha=m.activate.negative_slope, scale=m.activate.scale,
               noise_weight_dev=m.noise.weight.detach())
    if h * w <= _LOWRES_PIXELS and cin % 64 == 0 and x.shape[3] == cin:
        # input-modulated form: scale the (small) activation, convolve with shared cached weights
        xs = mc.scale_nhwc(x, s)
        d = (d_pre if d_pre is not None else mc.demod_from_wsq(s, wsq, conv.scale, conv.eps)) if conv.demodulate else None
        if conv.upsample and k == 3 and h * w >= _LOWRES_UP_SPLIT_PIXELS:
            wqs = _cached(conv, "wq_shared", [conv.weight], lambda: mc.pack_weights(w4, wscale=conv.scale)[0])
            y = mc.conv_transpose_s2(xs, wqs, cout, k, k, epi=mc.make_epilogue(row_scale=d) if d is not None else None,
                                     out_nhwc=True)
            nz = _noise_for(noise, b, 2 * h, 2 * w, x.device)
            return upfirdn_nhwc(y, conv.blur.kernel, pad=conv.blur.pad,
                                epi=mc.make_epilogue(noise=nz, residual=residual, residual2=residual2, **act))
        if conv.upsample and k == 3 and cout % 32 == 0:
            wq3 = _cached(conv, "wq_up2_shared", [conv.weight, conv.blur.kernel], lambda: mc.pack_weights(
                mc.compose_up2_weights(w4, conv.blur.kernel), wscale=conv.scale)[0])
            nz = _noise_for(noise, b, 2 * h, 2 * w, x.device)
            return mc.conv_up2_fused(xs, wq3, cout, epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual,
                                                                         residual2=residual2, **act))
        if not conv.upsample:
            wqs = _cached(conv, "wq_shared", [conv.weight], lambda: mc.pack_weights(w4, wscale=conv.scale)[0])
            if conv.downsample:
                xb = upfirdn_nhwc(xs, conv.blur.kernel, pad=conv.blur.pad)
                nz = _noise_for(noise, b, (xb.shape[1] - k) // 2 + 1, (xb.shape[2] - k) // 2 + 1, x.device)
                return mc.conv_fprop(xb, wqs, cout, k, k, 2, 0, 1, out_nhwc=True,
                                     epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual, residual2=residual2, **act))
            nz = _noise_for(noise, b, h, w, x.device)
            return mc.conv_fprop(xs, wqs, cout, k, k, 1, conv.padding, 1, out_nhwc=True,
                                 epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual, residual2=residual2, **act))
    if (conv.upsample and k == 3 and _UP2H and x.shape[3] == cin and mc.up2h_supported(cin, cout, h, w)
            and tuple(conv.blur.pad) == (1, 1) and tuple(conv.blur.kernel.shape) == (4, 4)
            and _separable_taps(conv.blur.kernel) is not None):
        # wide levels: horizontal half of the blur in the weights, vertical half in the epilogue (2x instead of 4x the FLOPs)
        fy, fx = _separable_taps(conv.blur.kernel)
        wc = _cached(conv, "w_up2h", [conv.weight, conv.blur.kernel], lambda: mc.compose_up2h_weights(w4, list(fx)))
        ky = _cached(conv, "ky_up2h", [conv.blur.kernel], lambda: (ctypes.c_float * 4)(*[float(fy[3 - u]) for u in range(4)]))
        wq2, _ = mc.pack_weights(wc, s, wscale=conv.scale)
        d = None
        if conv.demodulate:
            d = d_pre if d_pre is not None else mc.demod_from_wsq(s, wsq, conv.scale, conv.eps)
        nz = _noise_for(noise, b, 2 * h, 2 * w, x.device)
        return mc.conv_up2h(x, wq2, cout, ky, epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual,
                                                                    residual2=residual2, **act))
    if conv.upsample and k == 3 and cin <= _UP_FUSED_MAX_CIN and cout % 32 == 0 and w >= 32:
        # transposed conv + blur as ONE dense conv with the composite weights and a pixel-shuffle epilogue
        w3 = _cached(conv, "w_up2", [conv.weight, conv.blur.kernel], lambda: mc.compose_up2_weights(w4, conv.blur.kernel))
        wq3, _ = mc.pack_weights(w3, s, wscale=conv.scale)
        d = None
        if conv.demodulate:
            d = d_pre if d_pre is not None else mc.demod_from_wsq(s, wsq, conv.scale, conv.eps)
        nz = _noise_for(noise, b, 2 * h, 2 * w, x.device)
        return mc.conv_up2_fused(x, wq3, cout, epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual,
                                                                     residual2=residual2, **act))
    wq, d = mc.pack_weights(w4, s, wscale=conv.scale, eps=conv.eps, want_demod=conv.demodulate and d_pre is None, wsq=wsq)
    if d_pre is not None:
        d = d_pre
    if conv.upsample:
        y = mc.conv_transpose_s2(x, wq, cout, k, k, epi=mc.make_epilogue(row_scale=d) if d is not None else None,
                                 out_nhwc=True)
        nz = _noise_for(noise, b, 2 * h, 2 * w, x.device)
        return upfirdn_nhwc(y, conv.blur.kernel, pad=conv.blur.pad,
                            epi=mc.make_epilogue(noise=nz, residual=residual, residual2=residual2, **act))
    if conv.downsample:
        xb = upfirdn_nhwc(x, conv.blur.kernel, pad=conv.blur.pad)
        nz = _noise_for(noise, b, (xb.shape[1] - k) // 2 + 1, (xb.shape[2] - k) // 2 + 1, x.device)
        return mc.conv_fprop(xb, wq, cout, k, k, 2, 0, 1, out_nhwc=True,
                             epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual, residual2=residual2, **act))
    nz = _noise_for(noise, b, h, w, x.device)
    return mc.conv_fprop(x, wq, cout, k, k, 1, conv.padding, 1, out_nhwc=True,
                         epi=mc.make_epilogue(row_scale=d, noise=nz, residual=residual, residual2=residual2, **act))


def _plain_weights(conv):
    """Packed bf16 weights of an EqualConv2d (equalised-lr scale folded in), cached."""
    return _cached(conv, "wq", [conv.weight], lambda: mc.pack_weights(conv.weight.detach(), wscale=conv.scale)[0])


def smart_layer(m: SMART_layer, x, style, noise=None):
    """SMART_layer: 4 dilated modulated branches -> channel slices -> 3x3 fusion conv with the
    double activation + noise in its epilogue (models/RestoreNet.py:225-244)."""
    b, h, w, cin = x.shape
    branches = list(m.ModulatedConv2ds)
    cq = branches[0].out_channel
    cout = cq * len(branches)
    k = branches[0].kernel_size
    s = _modulation(m.modulation, style)
    wcat = _smart_wcat(m)
    wsq = _smart_wsq(m) if branches[0].demodulate else None
    d_pre = _demod_pre(m) if branches[0].demodulate else None
    dils = [br.dilation for br in branches]
    # one launch for all branches: the generic kernel's branch mode up to 64 pixels wide, the kh-folded row-ring kernel's
    # branch slices (x read from HBM once, the other three branches hit L2) for the wide 16/32-channel branches
    wide_fold = w >= 128 and cq in (16, 32) and cin <= 128 and _BRANCH_FOLD
    one_launch = (k == 3 and (w <= _BRANCH_MAX_W or wide_fold) and cq >= 16 and (cq & (cq - 1)) == 0 and len(branches) <= 4
                  and all(br.padding == br.dilation for br in branches))
    if one_launch and h * w <= _LOWRES_PIXELS and cin % 64 == 0 and x.shape[3] == cin:
        # input-modulated form on shared cached weights, all branches in one launch
        xs = mc.scale_nhwc(x, s)
        d = (d_pre if d_pre is not None else mc.demod_from_wsq(s, wsq, branches[0].scale, branches[0].eps)) if wsq is not None else None
        wqs = _cached(m, "wq_shared", [br.weight for br in branches], lambda: mc.pack_weights(wcat, wscale=branches[0].scale)[0])
        buf = mc.conv_branches(xs, wqs, cout, dils, epi=mc.make_epilogue(row_scale=d) if d is not None else None)
    else:
        wq, d = mc.pack_weights(wcat, s, wscale=branches[0].scale, eps=branches[0].eps,
                                want_demod=branches[0].demodulate and d_pre is None, wsq=wsq)
        if d_pre is not None:
            d = d_pre
        if one_launch:
            buf = mc.conv_branches(x, wq, cout, dils, epi=mc.make_epilogue(row_scale=d) if d is not None else None)
        else:
            buf = torch.empty((b, h, w, cout), dtype=torch.bfloat16, device=x.device)
            for j, br in enumerate(branches):
                dj = d[:, j * cq:(j + 1) * cq].contiguous() if d is not None else None
                mc.conv_fprop(x, wq[:, :, j * cq:(j + 1) * cq, :], cq, k, k, 1, br.padding, br.dilation, out=buf,
                              out_nhwc=True, co_off=j * cq, epi=mc.make_epilogue(row_scale=dj) if dj is not None else None)
    fconv, fact = m.fusion[0], m.fusion[1]
    nz = _noise_for(noise, b, h, w, x.device)
    kw = dict(pre_bias=fact.bias.detach(), pre_act=3, noise=nz, noise_weight_dev=m.noise.weight.detach(),
              alpha=fact.negative_slope, scale=fact.scale)
    if m.activate is not None:
        kw.update(bias=m.activate.bias.detach(), act=3)
    return mc.conv_fprop(buf, _plain_weights(fconv), cout, 3, 3, 1, fconv.padding, 1, out_nhwc=True,
                         epi=mc.make_epilogue(**kw))


def large_conv_layer(m: LargeConvLayer, x):
    """LargeConvLayer (un-modulated): dilated convs -> slices -> 1x1 fusion + two activations."""
    assert not m.downsample, "downsampling LargeConvLayer is not used by the networks"
    b, h, w, cin = x.shape
    convs = list(m.dilated_convs)
    cq = convs[0].weight.shape[0]
    cout = cq * len(convs)
    k = convs[0].weight.shape[2]
    fconv, fact = m.fusion[0], m.fusion[1]
    if k == 1 and fconv.weight.shape[2] == 1:
        # dilation is meaningless for 1x1, so the four branches are ONE 1x1 conv with concatenated weights — and with no
        # bias or activation between it and the 1x1 fusion conv (models/RestoreNet.py:770-787) the two compose exactly into a
        # single 1x1 conv, W = (scale_f * W_fusion) @ (scale_c * W_cat), formed once in fp32: one pass over the 512^2 image
        # instead of two and no bf16 intermediate (down_from_big: 364 + 402 us -> one 8 -> 64 launch per 32 faces)
        def compose():
            wc = torch.cat([c.weight.detach() for c in convs], 0)[:, :, 0, 0].double() * convs[0].scale
            wf = fconv.weight.detach()[:, :, 0, 0].double() * fconv.scale
            return _pack_padded((wf @ wc).float()[:, :, None, None].contiguous(), 1.0, cin)
        wq = _cached(m, "wq1x1_composed", [c.weight for c in convs] + [fconv.weight], compose)
        kw = dict(pre_bias=fact.bias.detach(), pre_act=3, alpha=fact.negative_slope, scale=fact.scale)
        if m.activate is not None:
            kw.update(bias=m.activate.bias.detach() if m.activate.bias is not None else None, act=3)
        return mc.conv_fprop(x, wq, cout, 1, 1, 1, 0, 1, out_nhwc=True, epi=mc.make_epilogue(**kw), algo_cin=cin + cout)
    if k == 1:
        wq = _cached(m, "wq1x1", [c.weight for c in convs], lambda: _pack_padded(
            torch.cat([c.weight.detach() for c in convs], 0), convs[0].scale, cin))
        buf = mc.conv_fprop(x, wq, cout, 1, 1, 1, 0, 1, out_nhwc=True)
    elif (k == 3 and w <= _BRANCH_MAX_W and cq >= 16 and (cq & (cq - 1)) == 0 and len(convs) <= 4
          and all(c.padding == c.dilation for c in convs)):
        wq = _cached(m, "wq_branches", [c.weight for c in convs], lambda: mc.pack_weights(
            torch.cat([c.weight.detach() for c in convs], 0).contiguous(), wscale=convs[0].scale)[0])
        buf = mc.conv_branches(x, wq, cout, [c.dilation for c in convs])
    else:
        buf = torch.empty((b, h, w, cout), dtype=torch.bfloat16, device=x.device)
        for j, c in enumerate(convs):
            mc.conv_fprop(x, _plain_weights(c), cq, k, k, 1, c.padding, c.dilation, out=buf, out_nhwc=True, co_off=j * cq)
    kw = dict(pre_bias=fact.bias.detach(), pre_act=3, alpha=fact.negative_slope, scale=fact.scale)
    if m.activate is not None:
        kw.update(bias=m.activate.bias.detach() if m.activate.bias is not None else None, act=3)
    return mc.conv_fprop(buf, _plain_weights(fconv), cout, 1, 1, 1, 0, 1, out_nhwc=True, epi=mc.make_epilogue(**kw))


# ----------------------------------------------------------------------------------------------
# two-term (hi + lo) bf16 operands for the low-resolution half of the restorer's encoder
# ----------------------------------------------------------------------------------------------
# The encoder's low-resolution layers end in ``final_linear`` -> x_global, which is concatenated into the style of EVERY decoder
# layer (models/RestoreNet.py:937-940, :1022-1037): a rounding error there is not a local pixel error, it perturbs all
# decoder modulations coherently.  The precision-floor experiment (tests/sim_bf16_floor.py) attributes half
# of the full-network max-abs error of an all-bf16 pipeline to these few layers (2.6 % of the FLOPs).  They therefore run
# with two-term operands: x = x_hi + x_lo, w = w_hi + w_lo (each term bf16, fp32 accumulation on the tensor core),
#   conv(x, w) ~= conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)            (error ~2^-17 instead of 2^-9)
# expressed to the SAME tcgen05 kernels as one convolution over 3*Cin channels: activation [hi | lo | hi], weight
# [w_hi | w_hi | w_lo]; activations between these layers stay fp32 NCHW (they are at most 32x32).
# Measured on the B200 (tests/parity_diag_fullsize.py, parity test's seed): no two-term layers 1.51e-2 of the range / 51.5 dB;
# <= 16x16: 1.06e-2 / 55.6 dB; <= 32x32 (default): 0.77e-2 / 57.7 dB, for 4 % of the throughput (1024 -> 983 faces/s).
_SPLIT_PIXELS = 0 if os.environ.get("VSP_NO_SPLIT_LOWRES") is not None else int(os.environ.get("VSP_SPLIT_PIXELS", "1024"))


def _split3_nhwc(x, s=None):
    """x [B,C,H,W] fp32 (* s [B,C]) -> [B,H,W,3C] bf16 = [hi | lo | hi] with hi = bf16(x), lo = bf16(x - hi)."""
    b, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((b, h, w, 3 * c), dtype=torch.bfloat16, device=x.device)
    with _lib.device_guard(x.device):
        rc = _lib.load().vsp_nchw_f32_to_nhwc_split3_bf16(ptr(x), ptr(s.contiguous()) if s is not None else None, ptr(y),
                                                          b, c, h * w, stream_ptr())
    _lib.check(rc, "nchw_f32_to_nhwc_split3_bf16")
    return y


def _split3_weights(owner, tag, sources, build_w, wscale):
    """Packed [1,taps,Cout,3*Cin] two-term weights [w_hi | w_hi | w_lo] of ``wscale * build_w()`` ([Cout,Cin,k,k]), cached."""
    def build():
        w = build_w().detach().float() * wscale
        hi = w.to(torch.bfloat16).float()
        lo = (w - hi).to(torch.bfloat16).float()
        return mc.pack_weights(torch.cat([hi, hi, lo], 1).contiguous())[0]
    return _cached(owner, tag, sources, build)


def smart_layer_split(m: SMART_layer, x, style, noise=None):
    """:func:`smart_layer` with two-term operands; x and the result are fp32 NCHW."""
    b, cin, h, w = x.shape
    branches = list(m.ModulatedConv2ds)
    cq, k = branches[0].out_channel, branches[0].kernel_size
    cout = cq * len(branches)
    s = _modulation(m.modulation, style)
    wsq = _smart_wsq(m) if branches[0].demodulate else None
    d_pre = _demod_pre(m) if branches[0].demodulate else None
    d = (d_pre if d_pre is not None else mc.demod_from_wsq(s, wsq, branches[0].scale, branches[0].eps)) if wsq is not None else None
    wq = _split3_weights(m, "wq_split3", [br.weight for br in branches], lambda: _smart_wcat(m), branches[0].scale)
    buf = mc.conv_branches(_split3_nhwc(x, s), wq, cout, [br.dilation for br in branches], out_nhwc=False, algo_cin=cin,
                           epi=mc.make_epilogue(row_scale=d) if d is not None else None)
    fconv, fact = m.fusion[0], m.fusion[1]
    nz = _noise_for(noise, b, h, w, x.device)
    kw = dict(pre_bias=fact.bias.detach(), pre_act=3, noise=nz, noise_weight_dev=m.noise.weight.detach(),
              alpha=fact.negative_slope, scale=fact.scale)
    if m.activate is not None:
        kw.update(bias=m.activate.bias.detach(), act=3)
    wf = _split3_weights(fconv, "wq_split3", [fconv.weight], lambda: fconv.weight, fconv.scale)
    return mc.conv_fprop(_split3_nhwc(buf), wf, cout, 3, 3, 1, fconv.padding, 1, epi=mc.make_epilogue(**kw), algo_cin=cout)


def styled_conv_down_split(m, x, style, noise=None):
    """StyledConv_down (blur pad (2,2) -> modulated 3x3 stride-2 conv -> noise -> bias + lrelu) with two-term operands;
    x and the result are fp32 NCHW.  The blur runs in fp32 on the modulated input (modulation commutes with the
    per-channel FIR)."""
    conv = m.conv
    b, cin, h, w = x.shape
    cout, k = conv.out_channel, conv.kernel_size
    s = _modulation(conv.modulation, style)
    d = None
    if conv.demodulate:
        d = _demod_pre(conv)
        if d is None:
            d = mc.demod_from_wsq(s, _conv_wsq(conv), conv.scale, conv.eps)
    xb = upfirdn2d_raw((x * s[:, :, None, None]).contiguous(), conv.blur.kernel, (1, 1), (1, 1), tuple(conv.blur.pad) * 2)
    wq = _split3_weights(conv, "wq_split3", [conv.weight], lambda: conv.weight.view(cout, cin, k, k), conv.scale)
    nz = _noise_for(noise, b, (xb.shape[2] - k) // 2 + 1, (xb.shape[3] - k) // 2 + 1, x.device)
    return mc.conv_fprop(_split3_nhwc(xb), wq, cout, k, k, 2, 0, 1, algo_cin=cin, epi=mc.make_epilogue(
        row_scale=d, noise=nz, bias=m.activate.bias.detach(), act=3, alpha=m.activate.negative_slope, scale=m.activate.scale,
        noise_weight_dev=m.noise.weight.detach()))


def large_conv_layer_split(m: LargeConvLayer, x):
    """:func:`large_conv_layer` (3x3 dilated form, the encoder's 4x4 head) with two-term operands; fp32 NCHW in/out."""
    convs = list(m.dilated_convs)
    cq, k = convs[0].weight.shape[0], convs[0].weight.shape[2]
    cout = cq * len(convs)
    assert k == 3 and not m.downsample and all(c.padding == c.dilation for c in convs)
    wq = _split3_weights(m, "wq_split3", [c.weight for c in convs],
                         lambda: torch.cat([c.weight.detach() for c in convs], 0), convs[0].scale)
    buf = mc.conv_branches(_split3_nhwc(x), wq, cout, [c.dilation for c in convs], out_nhwc=False, algo_cin=x.shape[1])
    fconv, fact = m.fusion[0], m.fusion[1]
    kw = dict(pre_bias=fact.bias.detach(), pre_act=3, alpha=fact.negative_slope, scale=fact.scale)
    if m.activate is not None:
        kw.update(bias=m.activate.bias.detach() if m.activate.bias is not None else None, act=3)
    wf = _split3_weights(fconv, "wq_split3", [fconv.weight], lambda: fconv.weight, fconv.scale)
    return mc.conv_fprop(_split3_nhwc(buf), wf, cout, 1, 1, 1, 0, 1, epi=mc.make_epilogue(**kw), algo_cin=cout)


def _split_ok(m):
    """Shapes the two-term path covers: 4 power-of-two 3x3 dilated branches with Cin a multiple of 64."""
    branches = list(getattr(m, "ModulatedConv2ds", [])) or list(getattr(m, "dilated_convs", []))
    if not branches:
        return False
    w = branches[0].weight
    cq, cin, k = w.shape[-4], w.shape[-3], w.shape[-1]
    return (k == 3 and len(branches) <= 4 and cq >= 16 and (cq & (cq - 1)) == 0 and cin % 64 == 0
            and all(br.padding == br.dilation for br in branches))


def _pack_padded(weight, wscale, cin_pad):
    """Pack [Cout,Cin,kh,kw] to bf16 with the input-channel axis padded to the activation's width."""
    cout, cin, kh, kw = weight.shape
    wq, _ = mc.pack_weights(weight, wscale=wscale)
    if wq.shape[3] == cin_pad:
        return wq
    out = torch.zeros((1, kh * kw, cout, cin_pad), dtype=torch.bfloat16, device=weight.device)
    out[..., :wq.shape[3]] = wq
    return out


def to_rgb(m: ToRGB, x, style, skip=None):
    """ToRGB: 1x1 modulated conv (no demod) + bias + FIR-upsampled skip; fp32 NCHW in/out for RGB.
    N = 3 makes this a memory-bound read of the feature map: dedicated SIMT kernel, modulation applied
    to the weights in shared memory (no weight prologue launch)."""
    conv = m.conv
    b, h, w, c = x.shape
    s = _modulation(conv.modulation, style)
    res = None
    if skip is not None:
        f = m.upsample.factor
        res = upfirdn2d_raw(skip, m.upsample.kernel, (f, f), (1, 1), (m.upsample.pad[0], m.upsample.pad[1]) * 2)
    bias = _cached(m, "bias3", [m.bias], lambda: m.bias.detach().reshape(3).contiguous())
    w3 = _cached(m, "w3", [conv.weight], lambda: _pad_cols(conv.weight.detach().reshape(3, conv.in_channel), c))
    if c != conv.in_channel:
        s = F.pad(s, (0, c - conv.in_channel))
    out = torch.empty((b, 3, h, w), dtype=torch.float32, device=x.device)
    with _lib.device_guard(x.device):
        rc = _lib.load().vsp_torgb_nhwc_bf16(ptr(x), ptr(w3), ptr(s.contiguous()), ptr(bias), ptr(res), ptr(out),
                                             b, h * w, c, conv.scale, stream_ptr())
    _lib.check(rc, "torgb_nhwc_bf16")
    return out


_pool_taps_cache: dict = {}


def _pool_upsample_taps(up):
    """3x3 composite (HOST floats) of ``Upsample`` (2x zero-stuffing + FIR, models/RestoreNet.py:43-61) followed by a 2x2
    mean: pooled[y, x] = sum_{dy,dx} K[dy, dx] * skip[y + dy - 1, x + dx - 1].  Derived numerically from the module's own
    filter and pads on an impulse (host, cached)."""
    kern = up.kernel
    key = (kern.data_ptr(), kern._version, tuple(kern.shape), up.factor, tuple(up.pad))
    hit = _pool_taps_cache.get(key)
    if hit is None:
        assert up.factor == 2
        k = kern.detach().double().cpu()
        kh, kw = k.shape
        p0, p1 = up.pad
        n = 7
        z = torch.zeros(2 * n + p0 + p1, 2 * n + p0 + p1, dtype=torch.float64)
        z[p0 + 2 * 3, p0 + 2 * 3] = 1.0                                     # impulse at low-res (3, 3), zero-stuffed + padded
        kf = torch.flip(k, [0, 1])
        oh, ow = z.shape[0] - kh + 1, z.shape[1] - kw + 1
        up_img = torch.zeros(oh, ow, dtype=torch.float64)
        for i in range(oh):
            for j in range(ow):
                up_img[i, j] = (z[i:i + kh, j:j + kw] * kf).sum()
        pooled = up_img[:2 * n, :2 * n].reshape(n, 2, n, 2).mean(dim=(1, 3))  # response at (y, x) to the impulse at (3, 3)
        taps = torch.zeros(3, 3, dtype=torch.float64)
        for dy in range(3):
            for dx in range(3):
                taps[dy, dx] = pooled[3 - (dy - 1), 3 - (dx - 1)]                 # K[dy,dx] multiplies skip[y+dy-1, x+dx-1]
        assert abs(float(pooled.sum() - taps.sum())) < 1e-9, "upsample + pool support exceeds 3x3"
        hit = _pool_taps_cache[key] = ((ctypes.c_float * 9)(*[float(v) for v in taps.flatten()]), kern)
    return hit[0]


def to_rgb_pooled(m: ToRGB, x, style, skip):
    """Last ToRGB of the decoder + face_pool in one kernel: AvgPool2x2(conv + bias + Upsample(skip)) (see
    vsp_torgb_pool2_nhwc_bf16); x [B,H,W,C] -> [B,3,H/2,W/2] fp32, skip [B,3,H/2,W/2]."""
    conv = m.conv
    b, h, w, c = x.shape
    assert h % 2 == 0 and w % 2 == 0 and skip is not None and tuple(skip.shape) == (b, 3, h // 2, w // 2)
    s = _modulation(conv.modulation, style)
    bias = _cached(m, "bias3", [m.bias], lambda: m.bias.detach().reshape(3).contiguous())
    w3 = _cached(m, "w3", [conv.weight], lambda: _pad_cols(conv.weight.detach().reshape(3, conv.in_channel), c))
    if c != conv.in_channel:
        s = F.pad(s, (0, c - conv.in_channel))
    out = torch.empty((b, 3, h // 2, w // 2), dtype=torch.float32, device=x.device)
    # upsample-then-pool of the skip = one 3x3 FIR pass at the output resolution (own upfirdn2d kernel, 3 channels: ~1 % of the
    # feature-map traffic); the ToRGB kernel then only adds it
    k3 = _cached(m.upsample, "pool_k3", [m.upsample.kernel], lambda: torch.tensor(
        list(_pool_upsample_taps(m.upsample)), dtype=torch.float32, device=x.device).view(3, 3).flip(0, 1).contiguous())
    res = upfirdn2d_raw(skip.contiguous(), k3, (1, 1), (1, 1), (1, 1, 1, 1))
    with _lib.device_guard(x.device):
        rc = _lib.load().vsp_torgb_pool2_nhwc_bf16(ptr(x), ptr(w3), ptr(s.contiguous()), ptr(bias), ptr(res), None,
                                                   ptr(out), b, h // 2, w // 2, c, conv.scale, stream_ptr())
    _lib.check(rc, "torgb_pool2_nhwc_bf16")
    return out


def _pad_cols(w, c):
    if w.shape[1] == c:
        return w.contiguous()
    out = torch.zeros((w.shape[0], c), dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


def _style_mlp(mapping, z):
    """Style MLP — PixelNorm + n_mlp x EqualLinear(activation="fused_lrelu") (models/RestoreNet.py:845-856) — with every layer
    as ONE launch of the grouped-linear kernel (equalised-lr scale, bias * lr_mul and the leaky relu in its epilogue).  The
    module path issues per layer weight * scale, bias * lr_mul, a library SIMT GEMM on 16 CTAs and the bias-act kernel: 9 + 20
    + 10 launches, ~350 us of a 32-face micro-batch for an [32, 512] MLP.  Falls back to the module for any other structure."""
    mods = list(mapping) if isinstance(mapping, torch.nn.Sequential) else []
    ok = (len(mods) >= 2 and type(mods[0]).__name__ == "PixelNorm" and z.is_cuda and z.dtype == torch.float32 and z.dim() == 2
          and all(type(m).__name__ == "EqualLinear" and getattr(m, "activation", None) and m.bias is not None
                  and m.weight.dtype == torch.float32 and m.weight.is_contiguous() for m in mods[1:]))
    if not ok:
        return mapping(z)
    x = mods[0](z)
    for lin in mods[1:]:
        bank = _banks_for(lin, lambda: ModulationBank([(lin, 0)]))
        x = bank(x[:, None, :])[0][id(lin)]
    return x


def _mapped_latent(mapping, n_latent, styles, inject_index, truncation, truncation_latent, input_is_latent):
    """Style MLP + truncation + broadcast/mixing to [B, n_latent, D] (models/RestoreNet.py:982-1011);
    written against attributes the reference's own modules also have, so it accepts either."""
    from .restorenet import assemble_latent

    if not input_is_latent:
        styles = [_style_mlp(mapping, s) for s in styles]
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    return assemble_latent(styles, n_latent, inject_index)


def _as_nhwc(t):
    """Accept the decoder features either as NHWC bf16 (fast path) or NCHW fp32 (reference layout)."""
    if t.dtype == torch.bfloat16:
        return t
    return mc.nchw_to_nhwc_bf16(t)


@torch.no_grad()
@_clears_banks
def restoration_forward(net, images, de_feats, pre_styles, noise_styles, inject_index=None, truncation=1,
                        truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True, tail_groups=1,
                        tail_hook=None):
    """``Restoration_net.forward`` (models/RestoreNet.py:968-1046) as a fused pipeline.
    images [B,3,S,S] fp32 -> restored [B,3,S,S] fp32."""
    b = images.shape[0]
    noise_latent = _mapped_latent(net.style, net.n_latent, noise_styles, inject_index, truncation, truncation_latent,
                                  input_is_latent)
    latent = torch.cat([pre_styles[:, :noise_latent.shape[1], :], noise_latent], dim=-1)
    if isinstance(noise, dict):
        # explicit noise for every layer of both halves: {"encoder": [14 images in encoder-layer order], "decoder": [15]}.
        # The reference's one list cannot say this (its encoder reads the REVERSED decoder list, whose down-conv shapes do
        # not match, models/RestoreNet.py:924-927), so parity tests with live noise paths use this form.
        noise_rev, noise = list(noise["encoder"]), list(noise["decoder"])
    else:
        if noise is None:
            noise = ([None] * net.num_layers if randomize_noise
                     else [getattr(net.noises, f"noise_{i}") for i in range(net.num_layers)])
        noise_rev = noise[::-1]
    lat_rev = torch.flip(latent, dims=[1])

    enc = net.encoder_convs
    banks = _banks_for(net, lambda: (
        ModulationBank([(enc[ii].modulation, ii, _demod_entry_smart(enc[ii])) for ii in range(0, len(enc), 2)] +
                       [(enc[ii + 1].conv.modulation, ii, _demod_entry_conv(enc[ii + 1])) for ii in range(0, len(enc), 2)]),
        ModulationBank([(net.conv1.modulation, 0, _demod_entry_smart(net.conv1)), (net.to_rgb1.conv.modulation, 1)] +
                       [e for q, (up, smart, rgb) in enumerate(zip(net.convs[::2], net.convs[1::2], net.to_rgbs))
                        for e in ((up.conv.modulation, 1 + 2 * q, _demod_entry_conv(up)),
                                  (smart.modulation, 2 + 2 * q, _demod_entry_smart(smart)),
                                  (rgb.conv.modulation, 3 + 2 * q))])))
    _bank_apply(banks[0], lat_rev)
    out = large_conv_layer(net.down_from_big, mc.nchw_to_nhwc_bf16(images, c_pad=8))
    features = []
    exact = None                                                            # fp32 NCHW activation once on the two-term path
    for ii in range(0, len(enc), 2):
        if (exact is None and out.shape[1] * out.shape[2] <= _SPLIT_PIXELS and _split_ok(enc[ii])
                and enc[ii + 1].conv.in_channel % 64 == 0 and enc[ii + 1].conv.kernel_size == 3):
            exact = mc.nhwc_bf16_to_nchw(out)
        if exact is not None:
            exact = smart_layer_split(enc[ii], exact, lat_rev[:, ii], noise_rev[ii])
            features.append(mc.nchw_to_nhwc_bf16(exact))
            exact = styled_conv_down_split(enc[ii + 1], exact, lat_rev[:, ii], noise_rev[ii + 1])
            continue
        out = smart_layer(enc[ii], out, lat_rev[:, ii], noise_rev[ii])
        features.append(out)
        out = styled_conv(enc[ii + 1], out, lat_rev[:, ii], noise_rev[ii + 1])
    if exact is not None and _split_ok(net.final_layer):
        head = large_conv_layer_split(net.final_layer, exact)              # [B,C,4,4] fp32
    else:
        src = out if exact is None else mc.nchw_to_nhwc_bf16(exact)
        head = large_conv_layer(net.final_layer, src).permute(0, 3, 1, 2).float()
    flat = head.reshape(b, -1)                                             # reference flattens NCHW
    x_global = net.final_linear[0](flat)                                   # Dropout2d is the identity in eval
    early = net.final_transfer(x_global).view(b, -1, 4, 4)
    features.append(mc.nchw_to_nhwc_bf16((head + early).contiguous()))
    features = features[::-1]

    n_sty = 2 * len(net.to_rgbs) + 2
    stys = torch.cat([latent[:, :n_sty], x_global[:, None, :].expand(-1, n_sty, -1)], dim=2)   # [B, n, 2048]
    _bank_apply(banks[1], stys)

    def sty(i):
        return stys[:, i]

    out = smart_layer(net.conv1, features[0], sty(0), noise[0])
    skip = to_rgb(net.to_rgb1, out, sty(1))
    i = 1
    levels = list(zip(net.convs[::2], net.convs[1::2], noise[1::2], noise[2::2], net.to_rgbs))
    for li, (up, smart, n_up, n_smart, rgb) in enumerate(levels):
        level = (i + 1) // 2
        if li == len(levels) - min(_TAIL_LEVELS, len(levels)) and tail_groups > 1 and b % tail_groups == 0:
            # Last level in sample groups (images are independent): a caller that streams results to the host copies group g
            # while groups g+1.. still compute (``tail_hook(g, lo, hi, restored)`` runs between groups; under CUDA-graph
            # capture it ends one graph and begins the next).  Same kernels on batch slices: results are bit-identical.
            global _bank_rows
            restored = torch.empty((b, 3) + tuple(images.shape[2:]), dtype=torch.float32, device=images.device)
            gsz = b // tail_groups
            if tail_hook is not None:
                tail_hook(-1, 0, 0, restored)                 # everything before the tail has been enqueued
            try:
                for g in range(tail_groups):
                    lo, hi = g * gsz, (g + 1) * gsz
                    _bank_rows = slice(lo, hi)
                    cut = lambda t: None if t is None else (t if t.shape[0] == 1 else t[lo:hi])
                    o, sk, j = out[lo:hi], skip[lo:hi], i
                    for (up_, smart_, nu, ns, rgb_) in levels[li:]:
                        lv = (j + 1) // 2
                        o = styled_conv(up_, o, sty(j)[lo:hi], cut(nu), residual=features[lv][lo:hi],
                                        residual2=_as_nhwc(de_feats[lv])[lo:hi])
                        o = smart_layer(smart_, o, sty(j + 1)[lo:hi], cut(ns))
                        sk = to_rgb(rgb_, o, sty(j + 2)[lo:hi], sk)
                        j += 2
                    restored[lo:hi] = sk
                    if tail_hook is not None:
                        tail_hook(g, lo, hi, restored)
            finally:
                _bank_rows = None
            _bank_clear()
            return restored
        out = styled_conv(up, out, sty(i), n_up, residual=features[level], residual2=_as_nhwc(de_feats[level]))
        out = smart_layer(smart, out, sty(i + 1), n_smart)
        skip = to_rgb(rgb, out, sty(i + 2), skip)
        i += 2
    _bank_clear()
    return skip


@torch.no_grad()
@_clears_banks
def generator_forward(gen, styles, inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                      noise=None, randomize_noise=True, return_features=True, features_nchw=False, pool_image=False):
    """Style decoder ``Generator.forward`` (e4e/models/stylegan2/model.py:475-552) fused.
    Returns (image fp32 NCHW, features) — features NHWC bf16, or NCHW fp32 when ``features_nchw``."""
    latent = _mapped_latent(gen.style, gen.n_latent, styles, inject_index, truncation, truncation_latent,
                            input_is_latent)
    b = latent.shape[0]
    if noise is None:
        noise = ([None] * gen.num_layers if randomize_noise
                 else [getattr(gen.noises, f"noise_{i}") for i in range(gen.num_layers)])
    bank = _banks_for(gen, lambda: (ModulationBank(
        [(gen.conv1.conv.modulation, 0, _demod_entry_conv(gen.conv1)), (gen.to_rgb1.conv.modulation, 1)] +
        [e for q, (up, conv, rgb) in enumerate(zip(gen.convs[::2], gen.convs[1::2], gen.to_rgbs))
         for e in ((up.conv.modulation, 1 + 2 * q, _demod_entry_conv(up)),
                   (conv.conv.modulation, 2 + 2 * q, _demod_entry_conv(conv)),
                   (rgb.conv.modulation, 3 + 2 * q))]),))[0]
    _bank_apply(bank, latent)
    const = _cached(gen.input, "nhwc", [gen.input.input], lambda: mc.nchw_to_nhwc_bf16(gen.input.input.detach()))
    out = styled_conv(gen.conv1, const.expand(b, -1, -1, -1).contiguous(), latent[:, 0], noise[0])
    skip = to_rgb(gen.to_rgb1, out, latent[:, 1])
    feats = [out] if return_features else []
    i = 1
    for up, conv, n_up, n_conv, rgb in zip(gen.convs[::2], gen.convs[1::2], noise[1::2], noise[2::2], gen.to_rgbs):
        out = styled_conv(up, out, latent[:, i], n_up)
        if return_features:
            feats.append(out)
        out = styled_conv(conv, out, latent[:, i + 1], n_conv)
        if pool_image and rgb is gen.to_rgbs[-1]:
            skip = to_rgb_pooled(rgb, out, latent[:, i + 2], skip)     # image at half resolution (face_pool fused in)
        else:
            skip = to_rgb(rgb, out, latent[:, i + 2], skip)
        i += 2
    _bank_clear()
    if return_features and features_nchw:
        feats = [mc.nhwc_bf16_to_nchw(f) for f in feats]
    return skip, (feats if return_features else None)


@torch.no_grad()
def decode_stage(decoder, codes, size, out_n_latent=16, noise=None):
    """First half of :func:`restore_faces`: style decoder features (+ its image, pooled to ``size``) from the w+ codes.
    Needs only the codes — the degraded images are not touched until :func:`restore_stage`."""
    _noise_pool.begin(codes.device)
    fuse_pool = decoder.size == 2 * size and len(decoder.to_rgbs) > 0 and decoder.to_rgbs[-1].upsample.factor == 2
    image, feats = generator_forward(decoder, [codes], input_is_latent=True, randomize_noise=True, pool_image=fuse_pool,
                                     noise=noise)
    if image.shape[-1] != size:
        # face_pool (e4e/models/psp.py:245-246): AdaptiveAvgPool2d to (size, size) is an exact k x k mean when divisible
        k = image.shape[-1] // size
        image = (F.avg_pool2d(image, k) if image.shape[-1] == k * size and image.shape[-2] == k * size
                 else F.adaptive_avg_pool2d(image, (size, size)))
    return image, feats[:out_n_latent]


@torch.no_grad()
def restore_stage(net, low_imgs, feats, codes, noise_styles, noise=None, tail_groups=1, tail_hook=None):
    """Second half of :func:`restore_faces`: the restoration network on the degraded images and the decoder features."""
    return restoration_forward(net, low_imgs, feats, codes, noise_styles, noise=noise, tail_groups=tail_groups,
                               tail_hook=tail_hook)


@torch.no_grad()
def restore_faces(net, decoder, low_imgs, codes, noise_styles=None, out_n_latent=16, dec_noise=None, net_noise=None):
    """The hot path of one restoration batch (restoration_test.py:130-131): style decoder features from
    the (diffused) w+ codes, then the restoration network.  Returns (restored, decoder image at 512).
    ``dec_noise`` / ``net_noise``: explicit NoiseInjection images (default: drawn per call, as the reference does) — a
    per-layer list for the decoder; a list or {"encoder": [...], "decoder": [...]} for the restorer."""
    if noise_styles is None:
        noise_styles = [torch.randn(low_imgs.shape[0], net.style_dim, device=low_imgs.device)]
    image, feats = decode_stage(decoder, codes, low_imgs.shape[-1], out_n_latent, noise=dec_noise)
    restored = restore_stage(net, low_imgs, feats, codes, noise_styles, noise=net_noise)
    return restored, image


def noise_shapes(net, decoder, batch):
    """Shapes of every NoiseInjection image of one hot-path pass: (decoder list, {"encoder": [...], "decoder": [...]})."""
    dec = [(batch, 1) + tuple(getattr(decoder.noises, f"noise_{i}").shape[2:]) for i in range(decoder.num_layers)]
    rdec = [(batch, 1) + tuple(getattr(net.noises, f"noise_{i}").shape[2:]) for i in range(net.num_layers)]
    renc, r = [], net.size
    for _ in range(0, len(net.encoder_convs), 2):
        renc += [(batch, 1, r, r), (batch, 1, r // 2, r // 2)]           # SMART layer at r, then its stride-2 conv
        r //= 2
    return dec, {"encoder": renc, "decoder": rdec}


class GraphedRestorer:
    """One micro-batch of :func:`restore_faces` captured as a CUDA graph ("capture the whole inference
    step in a CUDA graph").  Every entry point of the C ABI is capture-safe (explicit stream, no allocation, no sync;
    tensor maps are encoded on the host from pointers that the graph's private memory pool keeps fixed), so after two
    eager warm-up passes (which fill the weight / descriptor caches) the ~270 launches of a micro-batch are recorded once
    and replayed with two ``cudaGraphLaunch`` calls (decoder half, restorer half): no Python, ctypes or allocator work per
    batch.  Noise is still drawn
    per replay (the CUDA generator's Philox offset is graph-aware).

    ``g = GraphedRestorer(net, decoder, micro); restored, image = g(low, codes, z)`` — inputs must have exactly
    ``micro`` rows; outputs are copies (the static output buffers are overwritten by the next replay) unless
    ``clone=False``."""

    def __init__(self, net, decoder, micro, size=None, n_latent=None, device=None, warmup=2, explicit_noise=False,
                 tail_groups=1):
        device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        size = size or net.size
        n_latent = n_latent or decoder.n_latent
        self.micro, self.device = micro, device
        self.low = torch.zeros(micro, 3, size, size, device=device)
        self.codes = torch.zeros(micro, n_latent, decoder.style_dim, device=device)
        self.z = torch.zeros(micro, net.style_dim, device=device)
        # explicit_noise: the NoiseInjection images are static input buffers the caller fills (``dec_noise`` list,
        # ``net_noise`` dict, see :func:`noise_shapes`) instead of per-replay draws — what parity tests compare against
        self.dec_noise = self.net_noise = None
        if explicit_noise:
            dsh, nsh = noise_shapes(net, decoder, micro)
            self.dec_noise = [torch.zeros(sh, device=device) for sh in dsh]
            self.net_noise = {k: [torch.zeros(sh, device=device) for sh in v] for k, v in nsh.items()}
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                restore_faces(net, decoder, self.low, self.codes, [self.z], dec_noise=self.dec_noise, net_noise=self.net_noise)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        # two graphs sharing one memory pool: the decoder half needs only the codes, so a caller that streams its inputs
        # from the host can start it while the (200x larger) image copy is still in flight (``before_low``)
        self.graph = torch.cuda.CUDAGraph()
        self.graph_restore = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.image, feats = decode_stage(decoder, self.codes, size, noise=self.dec_noise)
        # tail_groups > 1: the restorer's last level runs in sample groups, each its own graph, so that a caller streaming the
        # results to the host can copy group g while groups g+1.. compute (``on_group`` of __call__): for a shard that is a
        # single micro-batch (one rank of an 8-GPU job) the device->host copy of the result is otherwise fully exposed
        self.tail_groups = tail_groups if (tail_groups > 1 and micro % tail_groups == 0) else 1
        self.graph_tail = []
        if self.tail_groups == 1:
            with torch.cuda.graph(self.graph_restore, pool=self.graph.pool()):
                self.restored = restore_stage(net, self.low, feats, self.codes, [self.z], noise=self.net_noise)
        else:
            state = {"cm": torch.cuda.graph(self.graph_restore, pool=self.graph.pool())}

            def hook(g, lo, hi, restored):
                state["cm"].__exit__(None, None, None)         # ends the main graph (g == -1) or tail graph g
                if g + 1 < self.tail_groups:
                    nxt = torch.cuda.CUDAGraph()
                    self.graph_tail.append(nxt)
                    state["cm"] = torch.cuda.graph(nxt, pool=self.graph.pool())
                    state["cm"].__enter__()
                else:
                    state["cm"] = None

            state["cm"].__enter__()
            try:
                self.restored = restore_stage(net, self.low, feats, self.codes, [self.z], noise=self.net_noise,
                                              tail_groups=self.tail_groups, tail_hook=hook)
            finally:
                if state["cm"] is not None:                    # an exception between hooks: leave capture mode
                    state["cm"].__exit__(None, None, None)
            _bank_clear()
        self.launches = _lib.launch_count() - n0          # sm_100a kernels of this library inside one replay of both

    def set_noise(self, dec_noise, net_noise):
        """Fill the static NoiseInjection buffers of an ``explicit_noise`` restorer (shapes: :func:`noise_shapes`)."""
        if self.dec_noise is None:
            raise RuntimeError("GraphedRestorer was captured without explicit_noise=True")
        for dst, src in zip(self.dec_noise, dec_noise):
            dst.copy_(src, non_blocking=True)
        for key in ("encoder", "decoder"):
            for dst, src in zip(self.net_noise[key], net_noise[key]):
                dst.copy_(src, non_blocking=True)

    def __call__(self, low, codes, z, clone=True, before_low=None, on_group=None):
        """``before_low``: optional callable run after the decoder half has been enqueued and before ``low`` is read (e.g.
        ``lambda: stream.wait_event(low_arrived)``).  ``on_group(g, lo, hi, restored)`` (``tail_groups`` > 1): called after
        the kernels of samples [lo, hi) have been enqueued — ``restored[lo:hi]`` is final once they finish."""
        if low.shape[0] != self.micro:
            raise ValueError(f"GraphedRestorer captured for micro-batch {self.micro}, got {low.shape[0]}")
        self.codes.copy_(codes, non_blocking=True)
        self.z.copy_(z, non_blocking=True)
        self.graph.replay()
        if before_low is not None:
            before_low()
        self.low.copy_(low, non_blocking=True)
        self.graph_restore.replay()
        gsz = self.micro // self.tail_groups
        for g, tail in enumerate(self.graph_tail):
            tail.replay()
            if on_group is not None:
                on_group(g, g * gsz, (g + 1) * gsz, self.restored)
        if clone:
            return self.restored.clone(), self.image.clone()
        return self.restored, self.image
