"""Pin the CPU oracle (oracle/) against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

UFD = load_golden("upfirdn2d")
ACT = load_golden("fused_act")
MOD = load_golden("modconv")


def _ufd_args(name):
    p = UFD[f"{name}.params"]
    return (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), tuple(int(v) for v in p[4:8])


@pytest.mark.parametrize("name", [str(n) for n in UFD["names"]])
def test_upfirdn2d_oracle_matches_reference(name):
    up, down, pad = _ufd_args(name)
    x, k, y = UFD[f"{name}.x"], UFD[f"{name}.k"], UFD[f"{name}.y"]
    got = oracle.upfirdn2d_ref(x, k, up, down, pad)
    assert got.shape == y.shape
    np.testing.assert_allclose(got, y, rtol=1e-5, atol=1e-6)
    got64 = oracle.upfirdn2d_ref(x.astype(np.float64), k.astype(np.float64), up, down, pad)
    np.testing.assert_allclose(got64, y, rtol=1e-5, atol=2e-6)
    port = oracle.upfirdn2d_native_port(torch.from_numpy(x), torch.from_numpy(k), up, down, pad).numpy()
    assert port.shape == y.shape
    np.testing.assert_allclose(port, y, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", [str(n) for n in UFD["names"]])
def test_upfirdn2d_oracle_backward_formula(name):
    """grad_input == upfirdn2d(grad_out, flip(k), up=down, down=up, pad=g_pad) (op/upfirdn2d.py:229-240)."""
    up, down, pad = _ufd_args(name)
    x, k, go, gx = UFD[f"{name}.x"], UFD[f"{name}.k"], UFD[f"{name}.go"], UFD[f"{name}.gx"]
    if x.size == 0:
        return
    gpad = oracle.upfirdn2d_grad_pads(x.shape[2:], go.shape[2:], k.shape, up, down, pad)
    got = oracle.upfirdn2d_ref(go, k[::-1, ::-1].copy(), down, up, gpad)
    assert got.shape == gx.shape
    np.testing.assert_allclose(got, gx, rtol=1e-5, atol=2e-6)


def test_upfirdn2d_constant_image_known_answer():
    """Normalised filter on a constant image returns the constant away from the borders."""
    k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32)
    k /= k.sum()
    x = np.full((1, 1, 12, 12), 2.5, np.float32)
    y = oracle.upfirdn2d_ref(x, k, 1, 1, (2, 1))
    np.testing.assert_allclose(y[..., 3:-3, 3:-3], 2.5, rtol=1e-6)
    y2 = oracle.upfirdn2d_ref(x, k * 4, 2, 1, (2, 1))
    np.testing.assert_allclose(y2[..., 4:-4, 4:-4], 2.5, rtol=1e-6)


@pytest.mark.parametrize("name", [str(n) for n in ACT["names"]])
def test_bias_act_oracle_matches_reference(name):
    x, y, go, gx, ggo, vx = (ACT[f"{name}.{s}"] for s in ("x", "y", "go", "gx", "ggo", "vx"))
    b = ACT[f"{name}.b"] if f"{name}.b" in ACT.files else None
    np.testing.assert_allclose(oracle.fused_leaky_relu_ref(x, b), y, rtol=1e-6, atol=1e-7)
    dx, db = oracle.fused_leaky_relu_grads_ref(go, y, b is not None)
    np.testing.assert_allclose(dx, gx, rtol=1e-6, atol=1e-7)
    if b is not None:
        np.testing.assert_allclose(db, ACT[f"{name}.gb"], rtol=1e-5, atol=1e-5)
    # second order: fused_bias_act(gg_in, gg_bias, ref=out, 3, 1) (op/fused_act.py:153-165)
    vb = ACT[f"{name}.vb"] if b is not None else None
    np.testing.assert_allclose(oracle.bias_act_ref(vx, vb, y, 3, 1), ggo, rtol=1e-5, atol=1e-6)


def test_bias_act_switch_branches():
    x = np.array([[-1.0, 2.0]], np.float32)
    ref = np.array([[1.0, -1.0]], np.float32)
    assert np.allclose(oracle.bias_act_ref(x, None, None, 1, 0, 0.2, 2.0), [[-2.0, 4.0]])
    assert np.allclose(oracle.bias_act_ref(x, None, ref, 3, 1, 0.5, 1.0), [[-1.0, 1.0]])
    assert np.allclose(oracle.bias_act_ref(x, None, ref, 3, 2, 0.5, 1.0), 0.0)
    assert np.allclose(oracle.bias_act_ref(x, None, ref, 1, 2, 0.5, 1.0), 0.0)


def _mod_kwargs(name):
    return dict(demodulate="nodemod" not in name and "torgb" not in name,
                upsample=name.endswith("_up"), downsample=name.endswith("_down"))


@pytest.mark.parametrize("name", [str(n) for n in MOD["names"]])
def test_modconv_oracle_matches_reference(name):
    x = torch.from_numpy(MOD[f"{name}.x"]).requires_grad_(True)
    style = torch.from_numpy(MOD[f"{name}.style"]).requires_grad_(True)
    w = torch.from_numpy(MOD[f"{name}.sd.weight"]).requires_grad_(True)
    if name.startswith("dilated"):
        rate = int(name.split("_r")[1])
        y = oracle.modulated_conv2d_ref(x, w, style, dilation=rate)
        gx, gs, gw = torch.autograd.grad(y, [x, style, w], torch.from_numpy(MOD[f"{name}.go"]))
    else:
        mw = torch.from_numpy(MOD[f"{name}.sd.modulation.weight"])
        mb = torch.from_numpy(MOD[f"{name}.sd.modulation.bias"])
        s = torch.nn.functional.linear(style, mw * (1.0 / mw.shape[1] ** 0.5), mb)  # EqualLinear, lr_mul=1
        y = oracle.modulated_conv2d_ref(x, w, s, **_mod_kwargs(name))
        gx, gs, gw = torch.autograd.grad(y, [x, style, w], torch.from_numpy(MOD[f"{name}.go"]))
    ref = MOD[f"{name}.y"]
    assert tuple(y.shape) == ref.shape
    np.testing.assert_allclose(y.detach().numpy(), ref, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gx.numpy(), MOD[f"{name}.gx"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gs.numpy(), MOD[f"{name}.gstyle"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(gw.numpy(), MOD[f"{name}.gw"], rtol=1e-3, atol=1e-4)


def test_network_oracle_matches_reference():
    """Functional CPU oracle of decoder + Restoration_net vs the real reference's outputs (size 16)."""
    from vspbfr_b200.restorenet import Restoration_net
    from vspbfr_b200.stylegan2 import Generator

    net_g = load_golden("networks")
    size = int(net_g["size"])
    torch.manual_seed(2024)
    net = Restoration_net(size, 512, 2, channel_multiplier=2)
    dec = Generator(size, 512, 2, channel_multiplier=2)
    low, codes, z = (torch.from_numpy(net_g[k]) for k in ("low", "codes", "z"))
    img, feats = oracle.generator_ref(dec.state_dict(), codes, size)
    np.testing.assert_allclose(img.numpy(), net_g["decoder_image"], rtol=1e-3, atol=2e-4)
    for i, f in enumerate(feats):
        np.testing.assert_allclose(f.numpy(), net_g[f"decoder_feat{i}"], rtol=1e-3, atol=2e-4)
    restored = oracle.restoration_ref(net.state_dict(), low, feats, codes, z, size, 2)
    np.testing.assert_allclose(restored.numpy(), net_g["restored"], rtol=1e-3, atol=5e-4)


# ---- plain-C restatement (oracle/c/vsp_oracle.c, built by oracle/build_c.py) against the same reference goldens --------
@pytest.mark.parametrize("name", [str(n) for n in UFD["names"]])
def test_c_oracle_upfirdn2d_matches_reference(name):
    from oracle import build_c as bc
    up, down, pad = _ufd_args(name)
    x, k, y, go, gx = (UFD[f"{name}.{s}"] for s in ("x", "k", "y", "go", "gx"))
    got = bc.upfirdn2d_c(x, k, up, down, pad)
    assert got.shape == y.shape
    np.testing.assert_allclose(got, y, rtol=1e-5, atol=1e-6)
    if x.size:      # and the backward form (op/upfirdn2d.py:229-240) through the same C routine
        gpad = oracle.upfirdn2d_grad_pads(x.shape[2:], go.shape[2:], k.shape, up, down, pad)
        np.testing.assert_allclose(bc.upfirdn2d_c(go, k[::-1, ::-1].copy(), down, up, gpad), gx, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", [str(n) for n in ACT["names"]])
def test_c_oracle_bias_act_matches_reference(name):
    from oracle import build_c as bc
    x, y, go, gx, ggo, vx = (ACT[f"{name}.{s}"] for s in ("x", "y", "go", "gx", "ggo", "vx"))
    b = ACT[f"{name}.b"] if f"{name}.b" in ACT.files else None
    np.testing.assert_allclose(bc.bias_act_c(x, b), y, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(bc.bias_act_c(go, None, y, 3, 1), gx, rtol=1e-6, atol=1e-7)
    vb = ACT[f"{name}.vb"] if b is not None else None
    np.testing.assert_allclose(bc.bias_act_c(vx, vb, y, 3, 1), ggo, rtol=1e-5, atol=1e-6)


def test_c_oracle_agrees_with_numpy_oracle_on_random_modes():
    from oracle import build_c as bc
    rng = np.random.default_rng(5)
    for _ in range(12):
        x = rng.standard_normal((2, 2, int(rng.integers(1, 12)), int(rng.integers(1, 12)))).astype(np.float32)
        k = rng.standard_normal((int(rng.integers(1, 5)), int(rng.integers(1, 5)))).astype(np.float32)
        up = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        down = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        pad = tuple(int(v) for v in rng.integers(-2, 5, size=4))
        want = oracle.upfirdn2d_ref(x, k, up, down, pad)
        got = bc.upfirdn2d_c(x, k, up, down, pad)
        assert got.shape == want.shape, (x.shape, k.shape, up, down, pad)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_conv_oracle_matches_reference_closed_set():
    """conv2d_ref (= F.conv2d, what conv2d_gradfix reduces to on the CPU) and its autograd against the reference's own
    outputs and first-/second-order gradients (tests/golden/gradfix.npz)."""
    import ast

    import torch.nn.functional as F

    g = load_golden("gradfix")
    for name in [str(n) for n in g["names"]]:
        t = lambda k: torch.from_numpy(g[f"{name}.{k}"])
        x, w, go = t("x").requires_grad_(True), t("w").requires_grad_(True), t("go").requires_grad_(True)
        b = t("b") if f"{name}.b" in g.files else None
        kw = ast.literal_eval(str(g[f"{name}.kw"]))
        y = F.conv_transpose2d(x, w, b, **kw) if bool(g[f"{name}.transposed"]) else oracle.conv2d_ref(x, w, b, **kw)
        np.testing.assert_allclose(y.detach().numpy(), g[f"{name}.y"], rtol=1e-5, atol=1e-5)
        gx, gw = torch.autograd.grad(y, [x, w], go, create_graph=True)
        np.testing.assert_allclose(gx.detach().numpy(), g[f"{name}.gx"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(gw.detach().numpy(), g[f"{name}.gw"], rtol=1e-4, atol=1e-4)
        (ggw,) = torch.autograd.grad(gx.pow(2).sum(), [w], retain_graph=True)
        np.testing.assert_allclose(ggw.numpy(), g[f"{name}.ggw_from_x"], rtol=1e-4, atol=1e-4)


def test_discriminator_oracle_matches_reference():
    """discriminator_ref / r1_step_ref against the reference's Discriminator forward and R1 step."""
    from vspbfr_b200.restorenet import Discriminator

    g = load_golden("gradfix")
    torch.manual_seed(4242)
    d = Discriminator(int(g["disc.size"]))
    with torch.no_grad():
        for n_, p in d.named_parameters():
            if n_.endswith("bias"):
                p.normal_(0, 0.2)
    sd = d.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["disc.sd_keys"]]
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g["disc.sd_sums"], rtol=1e-6, atol=1e-6)
    with torch.no_grad():
        pred8 = oracle.discriminator_ref(sd, torch.from_numpy(g["disc.img8"]))
    np.testing.assert_allclose(pred8.numpy(), g["disc.pred8"], rtol=1e-5, atol=1e-5)
    pred, grad_real, r1, grads = oracle.r1_step_ref(sd, torch.from_numpy(g["disc.real"]))
    np.testing.assert_allclose(pred.numpy(), g["disc.pred"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grad_real.numpy(), g["disc.grad_real"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(float(r1), float(g["disc.r1"]), rtol=1e-5)
    for k, absmax in zip([str(k) for k in g["disc.param_keys"]], g["disc.grad_absmax"]):
        np.testing.assert_allclose(float(grads[k].abs().max()), absmax, rtol=1e-3, atol=1e-9)
