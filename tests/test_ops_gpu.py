"""GPU parity of upfirdn2d / fused bias-act (through the C ABI) against the oracle and the
reference-generated golden vectors.  Tolerance (north_star): 1e-5 relative in fp32."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from vspbfr_b200 import op
from vspbfr_b200.op.upfirdn2d import upfirdn2d_bias_act

pytestmark = pytest.mark.gpu
UFD = load_golden("upfirdn2d")
ACT = load_golden("fused_act")
DEV = "cuda"
RTOL = 1e-5


def close(got, want, rtol=RTOL, atol=None):
    want = np.asarray(want)
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == want.shape, (got.shape, want.shape)
    if atol is None:
        atol = rtol * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def _ufd_args(name):
    p = UFD[f"{name}.params"]
    return (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), tuple(int(v) for v in p[4:8])


@pytest.mark.parametrize("name", [str(n) for n in UFD["names"]])
def test_upfirdn2d_golden_forward_backward(name):
    up, down, pad = _ufd_args(name)
    x = torch.from_numpy(UFD[f"{name}.x"]).to(DEV).requires_grad_(True)
    k = torch.from_numpy(UFD[f"{name}.k"]).to(DEV)
    y = op.upfirdn2d(x, k, up=up, down=down, pad=pad)
    close(y, UFD[f"{name}.y"])
    if y.numel():
        (gx,) = torch.autograd.grad(y, x, torch.from_numpy(UFD[f"{name}.go"]).to(DEV))
        close(gx, UFD[f"{name}.gx"])


BLUR = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32) / 64


@pytest.mark.parametrize("shape,gain,up,down,pad", [
    ((2, 16, 64, 64), 4, 2, 1, (2, 1)),       # Upsample (ToRGB skip)
    ((2, 8, 65, 65), 4, 1, 1, (1, 1)),        # Blur after transposed conv (odd width -> LDG staging)
    ((2, 8, 64, 64), 1, 1, 1, (2, 2)),        # Blur before stride-2 conv (TMA staging)
    ((1, 4, 129, 129), 4, 1, 1, (1, 1)),
    ((1, 4, 128, 128), 1, 1, 2, (1, 1)),      # backward of Upsample
    ((3, 5, 4, 4), 4, 2, 1, (2, 1)),          # tiny planes, several per block
    ((2, 7, 8, 8), 1, 1, 1, (2, 2)),
    ((1, 3, 16, 16), 4, 2, 1, (2, 1)),
    ((1, 2, 32, 32), 1, 1, 2, (1, 1)),
    ((1, 2, 200, 300), 4, 2, 1, (2, 1)),      # multiple tiles in x and y
    ((1, 2, 37, 53), 1, 1, 1, (1, 2)),
    ((1, 40, 16, 16), 1, 1, 1, (2, 2)),       # more planes than one z iteration
])
def test_upfirdn2d_vs_oracle_seeded(shape, gain, up, down, pad):
    rng = np.random.default_rng(hash((shape, up, down, pad)) % (2 ** 32))
    x = rng.standard_normal(shape).astype(np.float32)
    k = BLUR * gain
    want = oracle.upfirdn2d_ref(x.astype(np.float64), k.astype(np.float64), up, down, pad)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    y = op.upfirdn2d(xt, torch.from_numpy(k).to(DEV), up=up, down=down, pad=pad)
    close(y, want.astype(np.float32))
    go = rng.standard_normal(want.shape).astype(np.float32)
    (gx,) = torch.autograd.grad(y, xt, torch.from_numpy(go).to(DEV), create_graph=True)
    gpad = oracle.upfirdn2d_grad_pads(shape[2:], want.shape[2:], k.shape, up, down, pad)
    want_gx = oracle.upfirdn2d_ref(go.astype(np.float64), k[::-1, ::-1].astype(np.float64), down, up, gpad)
    close(gx, want_gx.astype(np.float32))
    # second order: d/d(go) <gx, v> = upfirdn2d(v) (the op is linear)
    v = rng.standard_normal(shape).astype(np.float32)
    gy = torch.from_numpy(go).to(DEV).requires_grad_(True)
    y2 = op.upfirdn2d(xt, torch.from_numpy(k).to(DEV), up=up, down=down, pad=pad)
    (gx2,) = torch.autograd.grad(y2, xt, gy, create_graph=True)
    (ggo,) = torch.autograd.grad(gx2, gy, torch.from_numpy(v).to(DEV))
    want_ggo = oracle.upfirdn2d_ref(v.astype(np.float64), k.astype(np.float64), up, down, pad)
    close(ggo, want_ggo.astype(np.float32))


def test_upfirdn2d_config1_full_size_properties():
    """BASELINE config 1 at full size: linearity + constant-image known answer + oracle on a slice."""
    torch.manual_seed(0)
    x = torch.randn(4, 512, 64, 64, device=DEV)
    k = torch.from_numpy(BLUR * 4).to(DEV)
    y = op.upfirdn2d(x, k, up=2, down=1, pad=(2, 1))
    assert y.shape == (4, 512, 128, 128)
    x2 = torch.randn_like(x)
    y2 = op.upfirdn2d(x2, k, up=2, down=1, pad=(2, 1))
    y12 = op.upfirdn2d(x + 2 * x2, k, up=2, down=1, pad=(2, 1))
    assert torch.allclose(y12, y + 2 * y2, rtol=1e-4, atol=1e-4)
    const = op.upfirdn2d(torch.full((1, 2, 64, 64), 1.5, device=DEV), k, up=2, down=1, pad=(2, 1))
    assert torch.allclose(const[..., 4:-4, 4:-4], torch.tensor(1.5, device=DEV), rtol=1e-6)
    sl = x[1:2, 100:104].cpu().numpy()
    want = oracle.upfirdn2d_ref(sl.astype(np.float64), (BLUR * 4).astype(np.float64), 2, 1, (2, 1))
    close(y[1:2, 100:104], want.astype(np.float32))


def test_upfirdn2d_largest_model_plane():
    """[1,32,1025,1025] -> [1,32,1024,1024] (style decoder tail), checked on two planes."""
    torch.manual_seed(1)
    x = torch.randn(1, 32, 1025, 1025, device=DEV)
    k = torch.from_numpy(BLUR * 4).to(DEV)
    y = op.upfirdn2d(x, k, pad=(1, 1))
    assert y.shape == (1, 32, 1024, 1024)
    for c in (0, 31):
        want = oracle.upfirdn2d_ref(x[:, c:c + 1].cpu().numpy().astype(np.float64), (BLUR * 4).astype(np.float64), 1, 1, (1, 1))
        close(y[:, c:c + 1], want.astype(np.float32))


@pytest.mark.parametrize("name", [str(n) for n in ACT["names"]])
def test_fused_leaky_relu_golden(name):
    x = torch.from_numpy(ACT[f"{name}.x"]).to(DEV).requires_grad_(True)
    has_b = f"{name}.b" in ACT.files
    b = torch.from_numpy(ACT[f"{name}.b"]).to(DEV).requires_grad_(True) if has_b else None
    y = op.fused_leaky_relu(x, b)
    close(y, ACT[f"{name}.y"])
    go = torch.from_numpy(ACT[f"{name}.go"]).to(DEV).requires_grad_(True)
    ins = (x, b) if has_b else (x,)
    grads = torch.autograd.grad(y, ins, go, create_graph=True)
    close(grads[0], ACT[f"{name}.gx"])
    s = (grads[0] * torch.from_numpy(ACT[f"{name}.vx"]).to(DEV)).sum()
    if has_b:
        close(grads[1], ACT[f"{name}.gb"], rtol=1e-5, atol=1e-4)
        s = s + (grads[1] * torch.from_numpy(ACT[f"{name}.vb"]).to(DEV)).sum()
    (ggo,) = torch.autograd.grad(s, go)
    close(ggo, ACT[f"{name}.ggo"])


@pytest.mark.parametrize("shape", [(4, 512, 32, 32), (2, 64, 128, 128), (8, 512), (3, 1024), (2, 8192),
                                   (2, 5, 7, 3), (1, 3, 1, 1), (2, 512, 4, 4), (1, 16, 16, 16)])
@pytest.mark.parametrize("slope,scale", [(0.2, 2 ** 0.5), (0.1, 1.0)])
def test_fused_leaky_relu_vs_oracle(shape, slope, scale):
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 32))
    x = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape[1]).astype(np.float32)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    bt = torch.from_numpy(b).to(DEV).requires_grad_(True)
    y = op.fused_leaky_relu(xt, bt, slope, scale)
    want = oracle.fused_leaky_relu_ref(x, b, slope, scale)
    close(y, want)
    go = rng.standard_normal(shape).astype(np.float32)
    gx, gb = torch.autograd.grad(y, (xt, bt), torch.from_numpy(go).to(DEV))
    wdx, wdb = oracle.fused_leaky_relu_grads_ref(go.astype(np.float64), want, True, slope, scale)
    close(gx, wdx.astype(np.float32))
    close(gb, wdb.astype(np.float32), rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(go).sum(axis=tuple(i for i in range(len(shape)) if i != 1)).max())))
    # no-bias path and the module
    y0 = op.fused_leaky_relu(xt, None, slope, scale)
    close(y0, oracle.fused_leaky_relu_ref(x, None, slope, scale))


def test_fused_leaky_relu_empty_and_module():
    m = op.FusedLeakyReLU(6).to(DEV)
    assert m(torch.zeros(0, 6, 4, 4, device=DEV)).shape == (0, 6, 4, 4)
    x = torch.randn(2, 6, 4, 4, device=DEV)
    with torch.no_grad():
        m.bias.copy_(torch.arange(6.0))
    close(m(x), oracle.fused_leaky_relu_ref(x.cpu().numpy(), np.arange(6, dtype=np.float32)))


def test_fused_upfirdn_bias_act_matches_two_pass():
    torch.manual_seed(3)
    x = torch.randn(2, 16, 32, 32, device=DEV, requires_grad=True)
    b = torch.randn(16, device=DEV, requires_grad=True)
    k = torch.from_numpy(BLUR * 4).to(DEV)
    y_f = upfirdn2d_bias_act(x, k, b, up=2, down=1, pad=(2, 1))
    y_2 = op.fused_leaky_relu(op.upfirdn2d(x, k, up=2, down=1, pad=(2, 1)), b)
    assert torch.allclose(y_f, y_2, rtol=1e-6, atol=1e-6)
    go = torch.randn_like(y_f)
    gf = torch.autograd.grad(y_f, (x, b), go)
    g2 = torch.autograd.grad(y_2, (x, b), go)
    assert torch.allclose(gf[0], g2[0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(gf[1], g2[1], rtol=1e-4, atol=1e-3)


def test_errors_are_runtime_errors():
    x = torch.zeros(1, 1, 4, 4, device=DEV, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        op.upfirdn2d(x, torch.ones(2, 2, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.zeros(2, 4, device=DEV), torch.zeros(5, device=DEV))


@pytest.mark.parametrize("n,c,h,w,pad,epi", [(2, 64, 33, 33, (1, 1), True), (1, 32, 65, 40, (1, 1), False),
                                             (2, 64, 32, 32, (2, 2), False), (1, 128, 17, 130, (2, 2), True),
                                             (3, 8, 9, 9, (1, 1), True), (1, 16, 100, 7, (2, 2), False),
                                             # TMA strip kernel: C % 64 == 0, width >= 35 (partial last strip, several
                                             # vertical segments, ring wrap-around, odd extents)
                                             (2, 64, 70, 69, (2, 2), False), (1, 64, 200, 40, (2, 2), True),
                                             (2, 192, 37, 36, (1, 1), False), (1, 64, 131, 97, (1, 1), True),
                                             (1, 128, 4, 35, (2, 2), True)])
@pytest.mark.parametrize("separable", [True, False])
def test_blur_nhwc_bf16_streaming_kernel(n, c, h, w, pad, epi, separable):
    """Channels-last bf16 blur (Blur of models/RestoreNet.py:85-101 inside the fused pipeline): the streaming separable
    kernel and its direct branch for non-separable filters, with the noise + bias + lrelu + residual epilogue,
    against the CPU oracle on the same bf16-rounded input."""
    import math
    import oracle
    from vspbfr_b200 import fastpath as fp
    from vspbfr_b200.op import modconv as mc
    rng = np.random.default_rng(n * 1000 + c + h + w)
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    k = (np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64).astype(np.float32)
    if not separable:
        k = k.copy()
        k[1, 2] += 0.03125
    xq = mc.nchw_to_nhwc_bf16(torch.from_numpy(x).cuda())
    xr = xq.float().permute(0, 3, 1, 2).contiguous().cpu().numpy()
    want = oracle.upfirdn2d_ref(xr, k, 1, 1, pad)
    e = None
    if epi:
        oh, ow = want.shape[2:]
        noise = torch.from_numpy(rng.standard_normal((n, 1, oh, ow)).astype(np.float32)).cuda()
        bias = torch.from_numpy(rng.standard_normal(c).astype(np.float32)).cuda()
        res = torch.from_numpy(rng.standard_normal((n, c, oh, ow)).astype(np.float32)).cuda()
        resq = mc.nchw_to_nhwc_bf16(res)
        e = mc.make_epilogue(noise=noise, noise_weight=0.25, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2), residual=resq)
        v = want + 0.25 * noise.cpu().numpy() + bias.cpu().numpy()[None, :, None, None]
        want = np.where(v > 0, v, 0.2 * v) * math.sqrt(2) + resq.float().permute(0, 3, 1, 2).cpu().numpy()
    y = fp.upfirdn_nhwc(xq, torch.from_numpy(k).cuda(), pad=pad, epi=e)
    got = y.float().permute(0, 3, 1, 2).cpu().numpy()
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-2 * max(1.0, float(np.abs(want).max())))
