"""GPU parity of the tcgen05 implicit-GEMM convolution (through the C ABI).

Two references: (1) the same convolution in fp32 on bf16-ROUNDED operands (isolates the kernel's
indexing/accumulation: only summation order differs -> tight tolerance), (2) the un-rounded fp32
result (the north_star bf16 tolerance: max-abs <= 1e-2 of the dynamic range, PSNR > 45 dB)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from vspbfr_b200.op import modconv as mc

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def bf16r(t):
    """Round like an ACTIVATION operand (bf16)."""
    return t.to(torch.bfloat16).to(torch.float32)


def wr(t):
    """Round like a packed WEIGHT operand: bf16 as well.  (fp16 weights against bf16 activations were tried — kind::f16
    has separate a_format / b_format fields — and trap with "illegal instruction" on B200: the formats must match.)"""
    return t.to(torch.bfloat16).to(torch.float32)


def assert_close_tight(got, want, tol=2e-3):
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(1.0, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


def psnr(got, want):
    peak = float(want.max() - want.min())
    mse = float(((got - want) ** 2).mean())
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


CASES = [
    # b, cin, cout, h, w, k, stride, pad, dil
    (2, 64, 64, 16, 16, 3, 1, 1, 1),
    (1, 128, 256, 32, 32, 3, 1, 1, 1),
    (2, 64, 16, 32, 32, 3, 1, 2, 2),
    (2, 64, 16, 32, 32, 3, 1, 4, 4),
    (1, 64, 16, 32, 32, 3, 1, 8, 8),
    (2, 32, 32, 64, 64, 3, 1, 1, 1),
    (2, 8, 24, 16, 16, 3, 1, 1, 1),
    (2, 64, 3, 16, 16, 1, 1, 0, 1),
    (3, 512, 512, 4, 4, 3, 1, 1, 1),
    (2, 512, 128, 8, 8, 3, 1, 1, 1),
    (1, 64, 64, 33, 33, 3, 1, 1, 1),
    (1, 64, 64, 17, 40, 3, 1, 1, 1),
    (2, 64, 128, 17, 17, 3, 2, 0, 1),
    (1, 128, 64, 65, 65, 3, 2, 0, 1),
    (1, 64, 64, 16, 16, 1, 2, 0, 1),
    (1, 64, 64, 256, 256, 3, 1, 1, 1),
    (1, 192, 320, 24, 24, 3, 1, 1, 1),
]


@pytest.mark.parametrize("b,cin,cout,h,w,k,stride,pad,dil", CASES)
def test_fprop_shared_weights(b, cin, cout, h, w, k, stride, pad, dil):
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + h)
    x = torch.randn(b, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(DEV)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(wt)
    out = mc.conv_fprop(xq, wq, cout, k, k, stride, pad, dil)
    want = F.conv2d(bf16r(x), wr(wt), None, stride, pad, dil)
    assert_close_tight(out, want)


def test_layout_roundtrip():
    x = torch.randn(3, 20, 9, 11, device=DEV)
    s = torch.rand(3, 20, device=DEV) + 0.5
    q = mc.nchw_to_nhwc_bf16(x, s)
    assert q.shape == (3, 9, 11, 24) and float(q[..., 20:].abs().max()) == 0.0
    want = bf16r(x * s[:, :, None, None])
    assert torch.equal(q[..., :20].permute(0, 3, 1, 2).float(), want)
    assert torch.equal(mc.nhwc_bf16_to_nchw(q, 20), want)
    assert torch.equal(mc.nchw_to_bf16(x, s).float(), want)


@pytest.mark.parametrize("transpose", [False, True])
def test_weight_prologue(transpose):
    torch.manual_seed(5)
    b, cout, cin, k = 3, 24, 40, 3
    w = torch.randn(cout, cin, k, k, device=DEV)
    s = torch.randn(b, cin, device=DEV)
    scale = 1 / math.sqrt(cin * k * k)
    wq, demod = mc.pack_weights(w, s, wscale=scale, transpose=transpose, want_demod=True, fold_demod=transpose)
    m = scale * w[None] * s[:, None, :, None, None]
    d = torch.rsqrt(m.pow(2).sum((2, 3, 4)) + 1e-8)
    torch.testing.assert_close(demod, d, rtol=1e-5, atol=1e-6)
    if transpose:
        want = (m * d[:, :, None, None, None]).permute(0, 3, 4, 2, 1).reshape(b, k * k, cin, cout)
    else:
        want = m.permute(0, 3, 4, 1, 2).reshape(b, k * k, cout, cin)
    torch.testing.assert_close(wq.float()[..., :want.shape[-1]], wr(want), rtol=1e-2, atol=1e-6)


def _modconv_ref(x, w, s, demod, stride=1, pad=1, dil=1, round_ops=True):
    b, cin = s.shape
    cout, _, k, _ = w.shape
    scale = 1 / math.sqrt(cin * k * k)
    m = scale * w[None] * s[:, None, :, None, None]
    d = torch.rsqrt(m.pow(2).sum((2, 3, 4)) + 1e-8) if demod else torch.ones(b, cout, device=x.device)
    outs = []
    for i in range(b):
        if round_ops:
            y = F.conv2d(bf16r(x[i:i + 1]), wr(m[i]), None, stride, pad, dil) * d[i][None, :, None, None]
        else:
            y = F.conv2d(x[i:i + 1], m[i] * d[i][:, None, None, None], None, stride, pad, dil)
        outs.append(y)
    return torch.cat(outs, 0)


@pytest.mark.parametrize("b,cin,cout,h,k,dil,demod", [(2, 64, 64, 16, 3, 1, True), (3, 128, 32, 16, 3, 2, True),
                                                       (2, 64, 3, 32, 1, 1, False), (4, 512, 512, 8, 3, 1, True)])
def test_fprop_per_sample_weights_and_demod_epilogue(b, cin, cout, h, k, dil, demod):
    torch.manual_seed(cin + cout)
    x = torch.randn(b, cin, h, h, device=DEV)
    w = torch.randn(cout, cin, k, k, device=DEV)
    s = torch.randn(b, cin, device=DEV) * 0.5 + 1.0
    scale = 1 / math.sqrt(cin * k * k)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, d = mc.pack_weights(w, s, wscale=scale, want_demod=demod)
    pad = (k - 1) * dil // 2
    out = mc.conv_fprop(xq, wq, cout, k, k, 1, pad, dil, epi=mc.make_epilogue(row_scale=d) if demod else None)
    assert_close_tight(out, _modconv_ref(x, w, s, demod, 1, pad, dil))
    full = _modconv_ref(x, w, s, demod, 1, pad, dil, round_ops=False)
    peak = float(full.max() - full.min())
    assert float((out - full).abs().max()) <= 1e-2 * peak
    assert psnr(out, full) > 45.0


def test_epilogue_noise_bias_act_residual_nhwc_and_nchw():
    torch.manual_seed(9)
    b, cin, cout, h = 2, 64, 48, 16
    x = torch.randn(b, cin, h, h, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / 24
    noise = torch.randn(b, 1, h, h, device=DEV)
    bias = torch.randn(cout, device=DEV)
    res = torch.randn(b, cout, h, h, device=DEV)
    res2 = torch.randn(b, cout, h, h, device=DEV)
    rs = torch.rand(b, cout, device=DEV) + 0.5
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(w)
    base = F.conv2d(bf16r(x), wr(w), None, 1, 1) * rs[:, :, None, None] + 0.37 * noise + bias[None, :, None, None]
    want = F.leaky_relu(base, 0.2) * math.sqrt(2)
    # NCHW fp32 output, fp32 residuals
    epi = mc.make_epilogue(row_scale=rs, noise=noise, noise_weight=0.37, bias=bias, act=3, alpha=0.2,
                           scale=math.sqrt(2), residual=res, residual2=res2)
    out = mc.conv_fprop(xq, wq, cout, 3, 3, 1, 1, 1, epi=epi)
    assert_close_tight(out, want + res + res2)
    # device-side noise weight, shared noise image
    nw = torch.tensor([0.37], device=DEV)
    epi = mc.make_epilogue(row_scale=rs, noise=noise[:1], noise_weight_dev=nw, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2))
    out = mc.conv_fprop(xq, wq, cout, 3, 3, 1, 1, 1, epi=epi)
    base1 = F.conv2d(bf16r(x), wr(w), None, 1, 1) * rs[:, :, None, None] + 0.37 * noise[:1] + bias[None, :, None, None]
    assert_close_tight(out, F.leaky_relu(base1, 0.2) * math.sqrt(2))
    # NHWC bf16 output with bf16 residual, written at a channel offset of a wider buffer
    resq = mc.nchw_to_nhwc_bf16(res)  # same layout as the output (48 channels)
    buf = torch.zeros(b, h, h, 64, dtype=torch.bfloat16, device=DEV)
    epi = mc.make_epilogue(row_scale=rs[:, :16].contiguous(), bias=bias[:16].contiguous(), act=3, alpha=0.2, scale=math.sqrt(2))
    wq16, _ = mc.pack_weights(w[:16].contiguous())
    mc.conv_fprop(xq, wq16, 16, 3, 3, 1, 1, 1, epi=epi, out=buf, out_nhwc=True, co_off=32)
    want16 = F.leaky_relu(F.conv2d(bf16r(x), wr(w[:16]), None, 1, 1) * rs[:, :16, None, None] + bias[None, :16, None, None], 0.2) * math.sqrt(2)
    got = buf[..., 32:48].permute(0, 3, 1, 2).float()
    assert_close_tight(got, want16, tol=1e-2)
    assert float(buf[..., :32].abs().max()) == 0 and float(buf[..., 48:].abs().max()) == 0
    epi = mc.make_epilogue(residual=resq)
    wq48, _ = mc.pack_weights(w)
    o2 = mc.conv_fprop(xq, wq48, cout, 3, 3, 1, 1, 1, epi=epi, out_nhwc=True)
    want2 = F.conv2d(bf16r(x), wr(w), None, 1, 1) + bf16r(res)
    assert_close_tight(o2[..., :cout].permute(0, 3, 1, 2).float(), want2, tol=1e-2)


@pytest.mark.parametrize("b,cin,cout,h,w_", [(2, 64, 64, 8, 8), (1, 128, 32, 16, 16), (2, 64, 64, 4, 4), (1, 64, 128, 7, 9)])
def test_transposed_stride2(b, cin, cout, h, w_):
    torch.manual_seed(h * 7 + cin)
    x = torch.randn(b, cin, h, w_, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(w)
    out = mc.conv_transpose_s2(xq, wq, cout, 3, 3)
    want = F.conv_transpose2d(bf16r(x), wr(w).transpose(0, 1), None, stride=2, padding=0)
    assert_close_tight(out, want)


@pytest.mark.parametrize("b,cin,cout,h,w_,shared", [(2, 64, 128, 32, 32, True), (2, 128, 256, 33, 40, False),
                                                     (1, 64, 512, 32, 64, True), (2, 192, 128, 130, 129, False),
                                                     (3, 64, 256, 64, 32, True)])
def test_transposed_stride2_one_launch_class_mode(b, cin, cout, h, w_, shared):
    """Stride-2 transposed conv with an NHWC bf16 output and Cout % 128 == 0: the four output parity classes as ONE
    pixel-shuffle launch (each channel tile runs only its class's 4 / 2 / 2 / 1 taps) plus thin launches for the last row
    and column, shared and per-sample weights, demodulation row scale in the epilogue; against conv_transpose2d
    (models/RestoreNet.py:522-529) on the same bf16-rounded operands.  Every output element is checked, so a class
    written to the wrong parity, a missing border row / column or a tile that ran another class's taps fails."""
    torch.manual_seed(h * 11 + cout)
    x = torch.randn(b, cin, h, w_, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    s = None if shared else torch.randn(b, cin, device=DEV) * 0.3 + 1
    rs = torch.rand(b, cout, device=DEV) + 0.5
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(w, s)
    out = mc.conv_transpose_s2(xq, wq, cout, 3, 3, epi=mc.make_epilogue(row_scale=rs), out_nhwc=True)
    assert out.shape == (b, 2 * h + 1, 2 * w_ + 1, cout)
    if shared:
        want = F.conv_transpose2d(bf16r(x), wr(w).transpose(0, 1), None, stride=2, padding=0)
    else:
        want = torch.cat([F.conv_transpose2d(bf16r(x[i:i + 1]), wr(w * s[i][None, :, None, None]).transpose(0, 1), None,
                                             stride=2, padding=0) for i in range(b)])
    want = want * rs[:, :, None, None]
    assert_close_tight(out.permute(0, 3, 1, 2).float(), want, tol=1e-2)


def test_dgrad_via_gather_matches_autograd():
    torch.manual_seed(21)
    b, cin, cout, h, k, dil = 2, 64, 96, 16, 3, 2
    pad = dil
    x = torch.randn(b, cin, h, h, device=DEV, requires_grad=True)
    w = torch.randn(cout, cin, k, k, device=DEV) / 24
    dy = torch.randn(b, cout, h, h, device=DEV)
    y = F.conv2d(x, wr(w), None, 1, pad, dil)
    (want,) = torch.autograd.grad(y, x, bf16r(dy))
    wq_t, _ = mc.pack_weights(w, transpose=True)
    dyq = mc.nchw_to_nhwc_bf16(dy)
    taps = [(i, j) for i in range(k) for j in range(k)]
    got = mc.conv_gather(dyq, wq_t, cin, [i * k + j for i, j in taps], [pad - i * dil for i, j in taps],
                         [pad - j * dil for i, j in taps], 1, (h, h))
    assert_close_tight(got, want)


def test_config2_full_size_forward():
    """BASELINE config 2: B=8, 512->512, 3x3, 64x64, demodulated."""
    torch.manual_seed(0)
    b, c, h = 8, 512, 64
    x = torch.randn(b, c, h, h, device=DEV)
    w = torch.randn(c, c, 3, 3, device=DEV)
    s = torch.randn(b, c, device=DEV) * 0.3 + 1.0
    scale = 1 / math.sqrt(c * 9)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, d = mc.pack_weights(w, s, wscale=scale, want_demod=True)
    out = mc.conv_fprop(xq, wq, c, 3, 3, 1, 1, 1, epi=mc.make_epilogue(row_scale=d))
    assert_close_tight(out, _modconv_ref(x, w, s, True))
    full = _modconv_ref(x, w, s, True, round_ops=False)
    peak = float(full.max() - full.min())
    assert float((out - full).abs().max()) <= 1e-2 * peak
    assert psnr(out, full) > 45.0


# ---------------------------------------------------------------------------------------------
# weight gradient (MN-major UMMA, K = pixels) and the full modulated-conv autograd Function
# ---------------------------------------------------------------------------------------------
from conftest import load_golden  # noqa: E402
from vspbfr_b200 import op  # noqa: E402

MOD = load_golden("modconv")


@pytest.mark.parametrize("b,cin,cout,h,w_,k,stride,pad,dil,groups", [
    (2, 64, 64, 16, 16, 3, 1, 1, 1, 2),
    (2, 64, 128, 16, 16, 3, 1, 1, 1, 1),
    (1, 128, 256, 8, 8, 3, 1, 1, 1, 1),
    (2, 256, 64, 32, 32, 3, 1, 2, 2, 2),
    (1, 64, 16, 64, 64, 3, 1, 4, 4, 1),
    (2, 32, 24, 12, 20, 3, 1, 1, 1, 2),
    (2, 64, 8, 16, 16, 1, 1, 0, 1, 2),
    (2, 64, 64, 17, 17, 3, 2, 0, 1, 2),
    (1, 64, 64, 128, 128, 3, 1, 1, 1, 1),
    (3, 512, 512, 4, 4, 3, 1, 1, 1, 3),
])
def test_wgrad(b, cin, cout, h, w_, k, stride, pad, dil, groups):
    torch.manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w_, device=DEV)
    oh, ow = mc.conv_out_size(h, k, stride, pad, dil), mc.conv_out_size(w_, k, stride, pad, dil)
    dy = torch.randn(b, cout, oh, ow, device=DEV)
    gw = mc.conv_wgrad(mc.nchw_to_nhwc_bf16(dy), mc.nchw_to_nhwc_bf16(x), groups, k, k, stride, pad, dil)
    assert gw.shape == (groups, k * k, cout, cin)
    wants = []
    for bi in range(b):
        xr = bf16r(x[bi:bi + 1]).requires_grad_(True)
        wdummy = torch.zeros(cout, cin, k, k, device=DEV, requires_grad=True)
        y = F.conv2d(xr, wdummy, None, stride, pad, dil)
        (g,) = torch.autograd.grad(y, wdummy, bf16r(dy[bi:bi + 1]))
        wants.append(g)
    want = torch.stack(wants)                      # [b, cout, cin, k, k]
    if groups == 1:
        want = want.sum(0, keepdim=True)
    want = want.reshape(groups, cout, cin, k * k).permute(0, 3, 1, 2)
    assert_close_tight(gw, want, tol=3e-3)


def _run_modconv(name, x, style_or_s, sd, rate=1):
    w = torch.from_numpy(sd["weight"]).to(DEV).requires_grad_(True)
    x = torch.from_numpy(x).to(DEV).requires_grad_(True)
    st = torch.from_numpy(style_or_s).to(DEV).requires_grad_(True)
    if "modulation.weight" in sd:
        mw = torch.from_numpy(sd["modulation.weight"]).to(DEV)
        mb = torch.from_numpy(sd["modulation.bias"]).to(DEV)
        s = F.linear(st, mw * (1.0 / mw.shape[1] ** 0.5), mb)
    else:
        s = st
    demod = "nodemod" not in name and "torgb" not in name
    mode = "up" if name.endswith("_up") else ("down" if name.endswith("_down") else "same")
    blur = torch.tensor(np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64.0, dtype=torch.float32, device=DEV)
    xin = x
    if mode == "down":
        xin = op.upfirdn2d(x, blur, pad=(2, 2))
    y = mc.modulated_conv2d(xin, w, s, demod, mode, rate)
    if mode == "up":
        y = op.upfirdn2d(y, blur * 4, pad=(1, 1))
    return x, st, w, y


@pytest.mark.parametrize("name", [str(n) for n in MOD["names"]])
def test_modulated_conv_golden_forward_backward(name):
    """bf16 tensor-core path vs the reference's fp32 CPU output: max-abs <= 1e-2 of the dynamic
    range and PSNR > 45 dB (north_star), for outputs and first-order gradients."""
    sd = {k.split(".sd.")[1]: MOD[k] for k in MOD.files if k.startswith(name + ".sd.")}
    rate = int(name.split("_r")[1]) if name.startswith("dilated") else 1
    x, st, w, y = _run_modconv(name, MOD[f"{name}.x"], MOD[f"{name}.style"], sd, rate)

    def check(got, want_np, what):
        want = torch.from_numpy(want_np).to(DEV)
        assert got.shape == want.shape, (what, got.shape, want.shape)
        peak = float(want.max() - want.min())
        err = float((got - want).abs().max())
        assert err <= 1e-2 * peak, f"{what}: max-abs {err} > 1e-2 * {peak}"
        assert psnr(got, want) > 45.0, f"{what}: psnr {psnr(got, want)}"

    check(y, MOD[f"{name}.y"], "y")
    gx, gs, gw = torch.autograd.grad(y, [x, st, w], torch.from_numpy(MOD[f"{name}.go"]).to(DEV))
    check(gx, MOD[f"{name}.gx"], "gx")
    check(gs, MOD[f"{name}.gstyle"], "gstyle")
    check(gw, MOD[f"{name}.gw"], "gw")


def test_modulated_conv_double_backward_runs_and_matches_fp32():
    """R1-style double backward through the modulated conv (create_graph=True)."""
    torch.manual_seed(3)
    b, cin, cout, h = 2, 16, 16, 8
    x = torch.randn(b, cin, h, h, device=DEV, requires_grad=True)
    w = torch.randn(1, cout, cin, 3, 3, device=DEV, requires_grad=True)
    s = (torch.randn(b, cin, device=DEV) * 0.3 + 1).requires_grad_(True)

    def penalty(fn):
        y = fn(x, w, s)
        (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        return torch.autograd.grad(gx.pow(2).sum(), [w, s])

    got = penalty(lambda x_, w_, s_: mc.modulated_conv2d(x_, w_, s_, True, "same", 1))

    def ref(x_, w_, s_):
        scale = 1 / math.sqrt(cin * 9)
        m = scale * w_ * s_.reshape(b, 1, cin, 1, 1)
        m = m * torch.rsqrt(m.pow(2).sum([2, 3, 4]) + 1e-8).reshape(b, cout, 1, 1, 1)
        return F.conv2d(x_.reshape(1, b * cin, h, h), m.reshape(b * cout, cin, 3, 3), padding=1, groups=b).reshape(b, cout, h, h)

    want = penalty(ref)
    for g, wv in zip(got, want):
        peak = float(wv.abs().max())
        assert float((g - wv).abs().max()) <= 3e-2 * peak


def test_two_stage_epilogue_smart_fusion():
    """conv -> +b1 -> lrelu*sqrt2 -> +noise -> +b2 -> lrelu*sqrt2 (SMART_layer.forward, models/RestoreNet.py:234-238)."""
    torch.manual_seed(13)
    b, c, h = 2, 64, 16
    x = torch.randn(b, c, h, h, device=DEV)
    w = torch.randn(c, c, 3, 3, device=DEV) / 24
    b1, b2 = torch.randn(c, device=DEV), torch.randn(c, device=DEV)
    noise = torch.randn(b, 1, h, h, device=DEV)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(w)
    epi = mc.make_epilogue(pre_bias=b1, pre_act=3, noise=noise, noise_weight=0.5, bias=b2, act=3, alpha=0.2, scale=math.sqrt(2))
    out = mc.conv_fprop(xq, wq, c, 3, 3, 1, 1, 1, epi=epi)
    y = F.leaky_relu(F.conv2d(bf16r(x), wr(w), None, 1, 1) + b1[None, :, None, None], 0.2) * math.sqrt(2)
    y = F.leaky_relu(y + 0.5 * noise + b2[None, :, None, None], 0.2) * math.sqrt(2)
    assert_close_tight(out, y)


HALO_CASES = [
    # b, cin, cout, h, w, dil   (stride-1 3x3 'same' convs on wide images -> row-halo kernel)
    (1, 64, 64, 8, 128, 1),
    (2, 64, 16, 6, 128, 2),
    (1, 64, 16, 20, 256, 4),
    (1, 64, 16, 20, 128, 8),
    (2, 32, 32, 5, 256, 1),
    (1, 128, 128, 6, 128, 1),
    (1, 128, 32, 10, 256, 2),
    (1, 256, 256, 4, 128, 1),
    (1, 256, 64, 12, 128, 8),
    (1, 64, 64, 5, 200, 1),
    (1, 64, 48, 3, 130, 4),
    (3, 64, 64, 3, 128, 1),
]


@pytest.mark.parametrize("b,cin,cout,h,w,dil", HALO_CASES)
@pytest.mark.parametrize("per_sample", [False, True])
def test_rowhalo_conv(b, cin, cout, h, w, dil, per_sample):
    g = torch.Generator(device="cpu").manual_seed(cin + cout + w + dil)
    x = torch.randn(b, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
    xq = mc.nchw_to_nhwc_bf16(x)
    if per_sample:
        s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV)
        wq, _ = mc.pack_weights(wt, s)
        want = torch.cat([F.conv2d(bf16r(x[i:i + 1]), wr(wt * s[i][None, :, None, None]), None, 1, dil, dil)
                          for i in range(b)])
    else:
        wq, _ = mc.pack_weights(wt)
        want = F.conv2d(bf16r(x), wr(wt), None, 1, dil, dil)
    out = mc.conv_fprop(xq, wq, cout, 3, 3, 1, dil, dil)
    assert_close_tight(out, want)


RING_CASES = [
    # b, cin, cout, h, w, k, dil   (wide shallow stride-1 layers -> row-ring kernel: rolling row slots, R rows per hand-off)
    (2, 64, 64, 37, 128, 3, 1),      # resident weights, R=4 with a remainder step
    (1, 32, 32, 64, 256, 3, 1),      # K-skip (Cin < 64), two strips
    (2, 64, 16, 24, 128, 3, 2),      # dilated chains
    (1, 64, 16, 32, 256, 3, 4),
    (1, 64, 16, 48, 128, 3, 8),
    (1, 128, 32, 20, 128, 3, 1),     # kc = 2, resident
    (1, 128, 32, 32, 128, 3, 4),
    (1, 128, 128, 21, 128, 3, 1),    # kc = 2, streamed weights, R=2
    (2, 64, 128, 19, 128, 3, 1),     # N = 128
    (1, 64, 48, 9, 200, 3, 1),       # ragged width and channels
    (2, 8, 64, 40, 128, 1, 1),       # 1x1, Cin = 8 (first LargeConvLayer)
    (1, 64, 64, 33, 256, 1, 1),      # 1x1
    (1, 128, 24, 17, 130, 1, 1),     # 1x1, kc = 2
    (4, 64, 64, 300, 128, 3, 1),     # many units per CTA: ring wrap, weight reload per sample
]


@pytest.mark.parametrize("b,cin,cout,h,w,k,dil", RING_CASES)
@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("nhwc", [False, True])
def test_ring_conv(b, cin, cout, h, w, k, dil, per_sample, nhwc):
    g = torch.Generator(device="cpu").manual_seed(cin + cout + w + dil + h)
    x = torch.randn(b, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(DEV)
    pad = dil * (k // 2)
    xq = mc.nchw_to_nhwc_bf16(x)
    if per_sample:
        s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV)
        wq, _ = mc.pack_weights(wt, s)
        want = torch.cat([F.conv2d(bf16r(x[i:i + 1]), wr(wt * s[i][None, :, None, None]), None, 1, pad, dil)
                          for i in range(b)])
    else:
        wq, _ = mc.pack_weights(wt)
        want = F.conv2d(bf16r(x), wr(wt), None, 1, pad, dil)
    if wq.shape[3] != xq.shape[3]:   # Cin padded to 8 on the activation side
        wq = F.pad(wq, (0, xq.shape[3] - wq.shape[3]))
    if nhwc:   # staged epilogue: swizzled shared-memory tile + TMA store
        out = mc.conv_fprop(xq, wq, cout, k, k, 1, pad, dil, out_nhwc=True)
        assert_close_tight(out[..., :cout].permute(0, 3, 1, 2).float(), want, tol=1e-2)
        assert float(out[..., cout:].abs().max() if out.shape[3] > cout else 0.0) == 0.0
    else:
        out = mc.conv_fprop(xq, wq, cout, k, k, 1, pad, dil)
        assert_close_tight(out, want)


def test_ring_conv_epilogue_nhwc_slices():
    """Row-ring kernel with the full fused epilogue, NHWC bf16 output at a channel offset (SMART branch layout)
    and the two-stage activation + noise + residuals (SMART fusion / StyledConv tails)."""
    torch.manual_seed(21)
    b, cin, cout, h, w = 2, 64, 64, 26, 256
    x = torch.randn(b, cin, h, w, device=DEV)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / 24
    noise = torch.randn(b, 1, h, w, device=DEV)
    b1, b2 = torch.randn(cout, device=DEV), torch.randn(cout, device=DEV)
    rs = torch.rand(b, cout, device=DEV) + 0.5
    res = torch.randn(b, cout, h, w, device=DEV)
    xq = mc.nchw_to_nhwc_bf16(x)
    wq, _ = mc.pack_weights(wt)
    nw = torch.tensor([0.41], device=DEV)
    resq = mc.nchw_to_nhwc_bf16(res)
    epi = mc.make_epilogue(row_scale=rs, pre_bias=b1, pre_act=3, noise=noise, noise_weight_dev=nw, bias=b2, act=3, alpha=0.2,
                           scale=math.sqrt(2), residual=resq)
    out = mc.conv_fprop(xq, wq, cout, 3, 3, 1, 1, 1, epi=epi, out_nhwc=True)
    y = F.leaky_relu(F.conv2d(bf16r(x), wr(wt), None, 1, 1) * rs[:, :, None, None] + b1[None, :, None, None], 0.2) * math.sqrt(2)
    y = F.leaky_relu(y + 0.41 * noise + b2[None, :, None, None], 0.2) * math.sqrt(2) + bf16r(res)
    assert_close_tight(out.permute(0, 3, 1, 2).float(), y, tol=1e-2)
    # four dilated 16-channel branches into slices of one 64-channel buffer
    buf = torch.zeros(b, h - 2, w, 64, dtype=torch.bfloat16, device=DEV)
    xs = x[:, :, :h - 2].contiguous()
    xsq = mc.nchw_to_nhwc_bf16(xs)
    for j, dil in enumerate((1, 2, 4, 8)):
        wj = wt[16 * j:16 * j + 16].contiguous()
        wqj, _ = mc.pack_weights(wj)
        epi = mc.make_epilogue(row_scale=rs[:, 16 * j:16 * j + 16].contiguous())
        mc.conv_fprop(xsq, wqj, 16, 3, 3, 1, dil, dil, epi=epi, out=buf, out_nhwc=True, co_off=16 * j)
        want = F.conv2d(bf16r(xs), wr(wj), None, 1, dil, dil) * rs[:, 16 * j:16 * j + 16, None, None]
        assert_close_tight(buf[..., 16 * j:16 * j + 16].permute(0, 3, 1, 2).float(), want, tol=1e-2)


@pytest.mark.parametrize("b,cin,cout,h,w_", [(2, 64, 32, 32, 32), (1, 128, 64, 16, 64), (2, 64, 64, 8, 160), (1, 256, 128, 32, 32),
                                             (3, 64, 32, 5, 96)])
@pytest.mark.parametrize("per_sample", [False, True])
def test_up2_fused_matches_transposed_conv_then_blur(b, cin, cout, h, w_, per_sample):
    """Fused up-conv (composite 6x6 stride-2 kernel as a dense 3x3 conv with a pixel-shuffle epilogue) against the
    reference formulation: conv_transpose2d(stride 2) -> upfirdn2d blur pad (1,1), with noise + bias + lrelu and two
    residuals in the epilogue (the StyledConv(up) tail of the restoration decoder)."""
    from oracle import upfirdn2d_ref
    g = torch.Generator(device="cpu").manual_seed(cin + cout + h + w_)
    x = torch.randn(b, cin, h, w_, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k4 = (torch.outer(k1, k1) / 64 * 4).to(DEV)
    s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV) if per_sample else None
    noise = torch.randn(b, 1, 2 * h, 2 * w_, generator=g).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    rs = (torch.rand(b, cout, generator=g) + 0.5).to(DEV)
    res = torch.randn(b, cout, 2 * h, 2 * w_, generator=g).to(DEV)
    res2 = torch.randn(b, cout, 2 * h, 2 * w_, generator=g).to(DEV)
    w3 = mc.compose_up2_weights(wt, k4)
    wq, _ = mc.pack_weights(w3, s)
    xq = mc.nchw_to_nhwc_bf16(x)
    epi = mc.make_epilogue(row_scale=rs, noise=noise, noise_weight=0.3, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2),
                           residual=mc.nchw_to_nhwc_bf16(res), residual2=mc.nchw_to_nhwc_bf16(res2))
    out = mc.conv_up2_fused(xq, wq, cout, epi=epi)
    assert out.shape == (b, 2 * h, 2 * w_, cout)
    # reference: per-sample transposed conv (fp32 on bf16-rounded activations), then the FIR blur via the CPU oracle
    ys = []
    for i in range(b):
        wi = wt * s[i][None, :, None, None] if per_sample else wt
        ys.append(F.conv_transpose2d(bf16r(x[i:i + 1]), wi.transpose(0, 1), stride=2))
    y = torch.cat(ys)
    yb = torch.from_numpy(upfirdn2d_ref(y.cpu().numpy(), k4.cpu().numpy(), 1, 1, (1, 1))).to(DEV)
    want = F.leaky_relu(yb * rs[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    want = want + bf16r(res) + bf16r(res2)
    got = out.permute(0, 3, 1, 2).float()
    # tight check of the kernel itself: fp32 conv with the bf16-rounded composite weights + pixel shuffle
    zs = []
    for i in range(b):
        wi = wr(w3 * s[i][None, :, None, None]) if per_sample else wr(w3)
        zs.append(F.pixel_shuffle(F.conv2d(bf16r(x[i:i + 1]), wi, None, 1, 1)
                                  .view(1, 4, cout, h, w_).transpose(1, 2).reshape(1, cout * 4, h, w_), 2))
    z = torch.cat(zs)
    want_t = F.leaky_relu(z * rs[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    assert_close_tight(got, want_t + bf16r(res) + bf16r(res2), tol=1e-2)
    # composite weights are rounded to bf16 once (instead of the 3x3 weights): compare at the bf16 tolerance
    assert_close_tight(got, want, tol=2e-2)
    assert psnr(got, want) > 45.0


@pytest.mark.parametrize("b,cin,cout,h,w_", [(2, 64, 32, 9, 128), (1, 128, 64, 7, 256), (2, 64, 64, 33, 160), (1, 128, 32, 5, 384),
                                             (3, 64, 32, 4, 128), (1, 64, 32, 70, 512)])
@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_up2h_half_composed_up_convolution(b, cin, cout, h, w_, per_sample, with_res):
    """vsp_conv2d_up2h_bf16 (horizontal blur in the weights, vertical blur in the epilogue from a TMEM ring) against the
    reference formulation conv_transpose2d(stride 2) -> upfirdn2d blur pad (1,1) with the StyledConv(up) tail, and tightly
    against an fp32 emulation on the same bf16-rounded operands (asymmetric filter: flips and tap order must be right)."""
    from oracle import upfirdn2d_ref
    g = torch.Generator(device="cpu").manual_seed(cin + cout + h + w_)
    x = torch.randn(b, cin, h, w_, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
    fy, fx = torch.tensor([1.0, 3.0, 4.0, 2.0]), torch.tensor([2.0, 5.0, 3.0, 1.0])
    fy, fx = fy / fy.sum() * 2, fx / fx.sum() * 2
    k4 = torch.outer(fy, fx).to(DEV)
    s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV) if per_sample else None
    noise = torch.randn(b, 1, 2 * h, 2 * w_, generator=g).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    rs = (torch.rand(b, cout, generator=g) + 0.5).to(DEV)
    res = torch.randn(b, cout, 2 * h, 2 * w_, generator=g).to(DEV)
    res2 = torch.randn(b, cout, 2 * h, 2 * w_, generator=g).to(DEV)
    assert mc.up2h_supported(cin, cout, h, w_)
    wc = mc.compose_up2h_weights(wt, fx.tolist())
    wq, _ = mc.pack_weights(wc, s)
    xq = mc.nchw_to_nhwc_bf16(x)
    kw = dict(residual=mc.nchw_to_nhwc_bf16(res), residual2=mc.nchw_to_nhwc_bf16(res2)) if with_res else {}
    epi = mc.make_epilogue(row_scale=rs, noise=noise, noise_weight=0.3, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2), **kw)
    ky = [float(fy[3 - u]) for u in range(4)]
    out = mc.conv_up2h(xq, wq, cout, ky, epi=epi)
    assert out.shape == (b, 2 * h, 2 * w_, cout)
    got = out.permute(0, 3, 1, 2).float()
    tail = (bf16r(res) + bf16r(res2)) if with_res else 0.0
    # tight: fp32 emulation of the kernel's own decomposition on the bf16-rounded composed weights
    from test_up2h_cpu import emulate_up2h
    zs = []
    for i in range(b):
        wi = wr(wc * s[i][None, :, None, None]) if per_sample else wr(wc)
        zs.append(emulate_up2h(bf16r(x[i:i + 1]).double().cpu(), wi.double().cpu(), ky, cout).float().to(DEV))
    z = torch.cat(zs)
    want_t = F.leaky_relu(z * rs[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2) + tail
    assert_close_tight(got, want_t, tol=1e-2)
    # reference formulation (3x3 weights un-rounded, blur in fp32)
    ys = []
    for i in range(b):
        wi = wt * s[i][None, :, None, None] if per_sample else wt
        ys.append(F.conv_transpose2d(bf16r(x[i:i + 1]), wi.transpose(0, 1), stride=2))
    yb = torch.from_numpy(upfirdn2d_ref(torch.cat(ys).cpu().numpy(), k4.cpu().numpy(), 1, 1, (1, 1))).to(DEV)
    want = F.leaky_relu(yb * rs[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2) + tail
    assert_close_tight(got, want, tol=2e-2)
    assert psnr(got, want) > 45.0


@pytest.mark.parametrize("b,cin,cq,h,w_", [(8, 512, 128, 4, 4), (3, 512, 128, 8, 8), (2, 128, 32, 16, 16), (2, 64, 16, 32, 32),
                                           (2, 256, 64, 64, 64), (5, 64, 64, 4, 8),
                                           # wide 16/32-channel branches -> branch slices of the kh-folded row-ring kernel
                                           (2, 64, 16, 64, 128), (1, 128, 32, 32, 256), (3, 64, 16, 128, 130), (1, 32, 16, 16, 128)])
@pytest.mark.parametrize("per_sample", [False, True])
def test_conv_branches_one_launch(b, cin, cq, h, w_, per_sample):
    """Four dilated branches (SMART_layer) as one launch == four separate dilated convs written to channel slices;
    shared weights stack several small samples into one 128-row tile, with the per-sample demod in the epilogue."""
    g = torch.Generator(device="cpu").manual_seed(cin + cq + h)
    dils = (1, 2, 4, 8)
    x = torch.randn(b, cin, h, w_, generator=g).to(DEV)
    wt = (torch.randn(4 * cq, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
    rs = (torch.rand(b, 4 * cq, generator=g) + 0.5).to(DEV)
    xq = mc.nchw_to_nhwc_bf16(x)
    if per_sample:
        s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV)
        wq, _ = mc.pack_weights(wt, s)
    else:
        wq, _ = mc.pack_weights(wt)
    out = mc.conv_branches(xq, wq, 4 * cq, dils, epi=mc.make_epilogue(row_scale=rs))
    for j, dil in enumerate(dils):
        wj = wt[j * cq:(j + 1) * cq]
        if per_sample:
            want = torch.cat([F.conv2d(bf16r(x[i:i + 1]), wr(wj * s[i][None, :, None, None]), None, 1, dil, dil)
                              for i in range(b)])
        else:
            want = F.conv2d(bf16r(x), wr(wj), None, 1, dil, dil)
        want = want * rs[:, j * cq:(j + 1) * cq, None, None]
        assert_close_tight(out[..., j * cq:(j + 1) * cq].permute(0, 3, 1, 2).float(), want, tol=1e-2)


@pytest.mark.parametrize("b,cin,cout,h,w_", [(8, 512, 512, 4, 4), (3, 128, 64, 8, 8), (2, 64, 32, 16, 16), (5, 64, 32, 4, 8)])
def test_up2_fused_lowres_shared_weights(b, cin, cout, h, w_):
    """Fused up-conv below 32 pixels wide: stacked tiles + pixel-shuffle through the direct epilogue, input-modulated
    (x * s with shared weights) == per-sample weights."""
    g = torch.Generator(device="cpu").manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w_, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k4 = (torch.outer(k1, k1) / 64 * 4).to(DEV)
    s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(DEV)
    noise = torch.randn(b, 1, 2 * h, 2 * w_, generator=g).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    wsq = mc.weight_sumsq(wt)
    d = mc.demod_from_wsq(s, wsq, 1.0)
    d_ref = torch.rsqrt(((wt[None] * s[:, None, :, None, None]) ** 2).sum(dim=(2, 3, 4)) + 1e-8)
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.cpu().numpy(), rtol=1e-4)
    w3 = mc.compose_up2_weights(wt, k4)
    xq = mc.nchw_to_nhwc_bf16(x)
    xs = mc.scale_nhwc(xq, s)
    np.testing.assert_allclose(xs.float().cpu().numpy(), (xq.float() * s[:, None, None, :]).to(torch.bfloat16).float().cpu().numpy())
    wq, _ = mc.pack_weights(w3)
    epi = mc.make_epilogue(row_scale=d, noise=noise, noise_weight=0.3, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2))
    out = mc.conv_up2_fused(xs, wq, cout, epi=epi)
    z = F.pixel_shuffle(F.conv2d(xs.permute(0, 3, 1, 2).float(), wr(w3), None, 1, 1)
                        .view(b, 4, cout, h, w_).transpose(1, 2).reshape(b, cout * 4, h, w_), 2)
    want = F.leaky_relu(z * d[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    assert_close_tight(out.permute(0, 3, 1, 2).float(), want, tol=1e-2)
    # and against the reference formulation (per-sample transposed conv -> blur -> demod)
    from oracle import upfirdn2d_ref
    y = torch.cat([F.conv_transpose2d(x[i:i + 1], (wt * s[i][None, :, None, None]).transpose(0, 1), stride=2) for i in range(b)])
    yb = torch.from_numpy(upfirdn2d_ref(y.cpu().numpy(), k4.cpu().numpy(), 1, 1, (1, 1))).to(DEV)
    ref = F.leaky_relu(yb * d_ref[:, :, None, None] + 0.3 * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    assert psnr(out.permute(0, 3, 1, 2).float(), ref) > 45.0


@pytest.mark.parametrize("b,cout,cin,h,w", [(2, 16, 3, 64, 64), (3, 20, 5, 12, 12), (1, 64, 3, 96, 40), (4, 7, 8, 10, 6)])
def test_conv1x1_small_cin_weight_gradient(b, cout, cin, h, w):
    """Weight gradient of the RGB-side 1x1 convolutions (Cin <= 8: LargeConvLayer 3->16, Discriminator stem 3->64) through
    conv2d_gradfix == autograd of F.conv2d in fp32 (streaming-reduction kernel, fp32 throughout)."""
    from vspbfr_b200.op import conv2d_gradfix
    torch.manual_seed(b * 100 + cout)
    x = torch.randn(b, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 1, 1, device="cuda", requires_grad=True)
    go = torch.randn(b, cout, h, w, device="cuda")
    (got,) = torch.autograd.grad(conv2d_gradfix.conv2d(x, wt), wt, go)
    w2 = wt.detach().clone().requires_grad_(True)
    (want,) = torch.autograd.grad(torch.nn.functional.conv2d(x, w2), w2, go)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
