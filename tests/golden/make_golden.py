#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (its Python/CPU code path).

The reference has no tests or golden vectors, so parity is pinned by
outputs of the reference itself, produced in the build container:

    cp -r /root/reference /tmp/refprobe && chmod -R u+w /tmp/refprobe
    TORCH_CUDA_ARCH_LIST=10.0a VSP_REF=/tmp/refprobe python tests/golden/make_golden.py

(the reference JIT-builds its extensions into its own directory at import time, hence
the writable copy — SURVEY.md Appendix C).  CPU tensors dispatch to
``upfirdn2d_native`` (op/upfirdn2d.py:356-357), the ``F.leaky_relu`` branch
(op/fused_act.py:217-228) and ``F.conv2d`` (op/conv2d_gradfix.py:34-42).

Nothing in tests/ or bench.py needs the reference at run time: only these fixtures.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("VSP_REF", "/tmp/refprobe")
OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")

sys.path.insert(0, REF)
m = types.ModuleType("matplotlib")
m.use = lambda *a, **k: None
sys.modules["matplotlib"] = m
import op as ref_op  # noqa: E402  (JIT-builds the reference extensions)

sys.modules["op.fused_act_cpu"] = sys.modules["op.fused_act"]
sys.modules["op.upfirdn2d_cpu"] = sys.modules["op.upfirdn2d"]
from models import RestoreNet as R  # noqa: E402
from e4e.models.stylegan2 import model as S  # noqa: E402

SYM6 = (0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
        0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
        -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148)


def np_(t):
    return t.detach().cpu().numpy()


def blur_k(gain=1.0):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None, :] * k[:, None]
    return k / k.sum() * gain


def gen_upfirdn2d():
    g = torch.Generator().manual_seed(1234)
    sym = torch.tensor(SYM6)
    cases = [
        # name, shape, kernel, up, down, pad
        ("upsample_2x", (2, 3, 8, 8), blur_k(4.0), 2, 1, (2, 1)),
        ("upsample_2x_odd", (1, 2, 5, 7), blur_k(4.0), 2, 1, (2, 1)),
        ("blur_up_pad11", (2, 4, 9, 9), blur_k(4.0), 1, 1, (1, 1)),
        ("blur_up_pad11_33", (1, 3, 33, 33), blur_k(4.0), 1, 1, (1, 1)),
        ("blur_down_pad22", (2, 4, 8, 8), blur_k(), 1, 1, (2, 2)),
        ("blur_skip_pad11", (1, 4, 8, 8), blur_k(), 1, 1, (1, 1)),
        ("downsample_2x", (2, 3, 16, 16), blur_k(), 1, 2, (1, 1)),
        ("downsample_2x_odd", (1, 2, 11, 13), blur_k(), 1, 2, (1, 1)),
        ("wide_tiles", (1, 2, 40, 150), blur_k(), 1, 1, (2, 2)),
        ("wide_up", (1, 1, 20, 70), blur_k(4.0), 2, 1, (2, 1)),
        ("up_pad_odd", (1, 2, 6, 6), blur_k(4.0), 2, 1, (1, 2)),
        ("k3", (1, 2, 9, 9), torch.rand(3, 3, generator=g), 1, 1, (1, 1)),
        ("k2_up2", (1, 2, 6, 6), torch.rand(2, 2, generator=g), 2, 1, (1, 0)),
        ("sym6_up_x", (1, 3, 10, 12), sym.unsqueeze(0), (2, 1), 1, (6, 5, 0, 0)),
        ("sym6_up_y", (1, 3, 10, 12), sym.unsqueeze(1), (1, 2), 1, (0, 0, 6, 5)),
        ("sym6_down_x", (1, 3, 24, 28), torch.flip(sym, (0,)).unsqueeze(0), 1, (2, 1), (-1, -1, 0, 0)),
        ("sym6_down_y", (1, 3, 24, 28), torch.flip(sym, (0,)).unsqueeze(1), 1, (1, 2), (0, 0, -1, -1)),
        ("neg_pad_crop", (1, 2, 12, 12), blur_k(), 1, 1, (-2, 1, 3, -1)),
        ("up3_down2_k5", (1, 2, 7, 9), torch.rand(5, 5, generator=g), 3, 2, (3, 2)),
        ("empty_batch", (0, 3, 8, 8), blur_k(), 1, 1, (1, 1)),
    ]
    out = {}
    names = []
    for name, shape, k, up, down, pad in cases:
        x = torch.randn(*shape, generator=g, requires_grad=True)
        y = ref_op.upfirdn2d(x, k, up=up, down=down, pad=pad)
        go = torch.randn(*y.shape, generator=g)
        (gx,) = torch.autograd.grad(y, x, go) if y.numel() else (torch.zeros_like(x),)
        up_t = up if isinstance(up, tuple) else (up, up)
        down_t = down if isinstance(down, tuple) else (down, down)
        out[f"{name}.x"] = np_(x)
        out[f"{name}.k"] = np_(k)
        out[f"{name}.params"] = np.array([*up_t, *down_t, *(pad if len(pad) == 4 else (pad[0], pad[1], pad[0], pad[1]))])
        out[f"{name}.y"] = np_(y)
        out[f"{name}.go"] = np_(go)
        out[f"{name}.gx"] = np_(gx)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "upfirdn2d.npz"), **out)
    print("upfirdn2d:", len(names), "cases")


def gen_fused_act():
    g = torch.Generator().manual_seed(4321)
    out = {}
    names = []
    for name, shape, has_bias in [("nchw_bias", (2, 6, 5, 7), True), ("nchw_nobias", (2, 4, 8, 8), False),
                                  ("nc_bias", (3, 16), True), ("nchw_big", (1, 8, 32, 32), True)]:
        x = torch.randn(*shape, generator=g, requires_grad=True)
        b = torch.randn(shape[1], generator=g, requires_grad=True) if has_bias else None
        y = ref_op.fused_leaky_relu(x, b)  # CPU branch: slope hard-coded 0.2, scale sqrt(2)
        go = torch.randn(*shape, generator=g, requires_grad=True)
        ins = (x, b) if has_bias else (x,)
        grads = torch.autograd.grad(y, ins, go, create_graph=True)
        # second order: v-weighted first-order grads differentiated w.r.t. grad_out
        v_x = torch.randn(*shape, generator=g)
        s = (grads[0] * v_x).sum()
        out[f"{name}.vx"] = np_(v_x)
        if has_bias:
            v_b = torch.randn(shape[1], generator=g)
            s = s + (grads[1] * v_b).sum()
            out[f"{name}.vb"] = np_(v_b)
            out[f"{name}.b"] = np_(b)
            out[f"{name}.gb"] = np_(grads[1])
        (ggo,) = torch.autograd.grad(s, go)
        out[f"{name}.x"] = np_(x)
        out[f"{name}.y"] = np_(y)
        out[f"{name}.go"] = np_(go)
        out[f"{name}.gx"] = np_(grads[0])
        out[f"{name}.ggo"] = np_(ggo)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "fused_act.npz"), **out)
    print("fused_act:", len(names), "cases")


def gen_modconv():
    out = {}
    names = []
    torch.manual_seed(7)
    specs = [
        # name, cls, cin, cout, k, style_dim, kwargs, H
        ("mod3x3", R.ModulatedConv2d, 16, 32, 3, 8, {}, 8),
        ("mod3x3_nodemod", R.ModulatedConv2d, 16, 16, 3, 8, {"demodulate": False}, 8),
        ("mod1x1_torgb", R.ModulatedConv2d, 16, 3, 1, 8, {"demodulate": False}, 8),
        ("mod_up", R.ModulatedConv2d, 16, 16, 3, 8, {"upsample": True}, 6),
        ("mod_down", R.ModulatedConv2d, 16, 32, 3, 8, {"downsample": True}, 8),
        ("e4e_mod3x3", S.ModulatedConv2d, 16, 16, 3, 8, {}, 8),
        ("e4e_mod_up", S.ModulatedConv2d, 16, 32, 3, 8, {"upsample": True}, 4),
    ]
    for name, cls, cin, cout, k, sd, kw, h in specs:
        mod = cls(cin, cout, k, sd, **kw)
        x = torch.randn(2, cin, h, h, requires_grad=True)
        style = torch.randn(2, sd, requires_grad=True)
        y = mod(x, style)
        go = torch.randn_like(y)
        params = [mod.weight, mod.modulation.weight, mod.modulation.bias]
        grads = torch.autograd.grad(y, [x, style] + params, go)
        out[f"{name}.x"], out[f"{name}.style"], out[f"{name}.y"], out[f"{name}.go"] = np_(x), np_(style), np_(y), np_(go)
        out[f"{name}.gx"], out[f"{name}.gstyle"] = np_(grads[0]), np_(grads[1])
        out[f"{name}.gw"], out[f"{name}.gmw"], out[f"{name}.gmb"] = np_(grads[2]), np_(grads[3]), np_(grads[4])
        for kname, v in mod.state_dict().items():
            out[f"{name}.sd.{kname}"] = np_(v)
        names.append(name)
    # dilated branch takes an already-modulated style
    for rate in (1, 2, 4):
        name = f"dilated_r{rate}"
        mod = R.Dilated_ModulatedConv2d(16, 8, 3, 8, dilation=rate)
        x = torch.randn(2, 16, 12, 12, requires_grad=True)
        ms = torch.randn(2, 16, requires_grad=True)
        y = mod(x, ms)
        go = torch.randn_like(y)
        grads = torch.autograd.grad(y, [x, ms, mod.weight], go)
        out[f"{name}.x"], out[f"{name}.style"], out[f"{name}.y"], out[f"{name}.go"] = np_(x), np_(ms), np_(y), np_(go)
        out[f"{name}.gx"], out[f"{name}.gstyle"], out[f"{name}.gw"] = np_(grads[0]), np_(grads[1]), np_(grads[2])
        out[f"{name}.sd.weight"] = np_(mod.weight)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "modconv.npz"), **out)
    print("modconv:", len(names), "cases")


def gen_layers():
    """Composite layers with explicit noise / non-zero noise weight and biases."""
    out = {}
    names = []
    torch.manual_seed(11)

    def randomize(mod):
        with torch.no_grad():
            for n_, p in mod.named_parameters():
                if n_.endswith("noise.weight"):
                    p.fill_(0.37)
                elif n_.endswith("bias") and p.ndim == 1 and "modulation" not in n_:
                    p.normal_(0, 0.3)
                elif n_ == "bias":
                    p.normal_(0, 0.3)

    def record(name, mod, inputs, y):
        for i, t in enumerate(inputs):
            out[f"{name}.in{i}"] = np_(t)
        out[f"{name}.y"] = np_(y)
        for kname, v in mod.state_dict().items():
            out[f"{name}.sd.{kname}"] = np_(v)
        names.append(name)

    sc = R.StyledConv(16, 16, 3, 8); randomize(sc)
    x, st, nz = torch.randn(2, 16, 8, 8), torch.randn(2, 8), torch.randn(2, 1, 8, 8)
    record("styledconv", sc, (x, st, nz), sc(x, st, noise=nz))

    scu = R.StyledConv(16, 32, 3, 8, upsample=True); randomize(scu)
    nz2 = torch.randn(2, 1, 16, 16)
    record("styledconv_up", scu, (x, st, nz2), scu(x, st, noise=nz2))

    scd = R.StyledConv_down(16, 32, 3, 8); randomize(scd)
    nz3 = torch.randn(2, 1, 4, 4)
    record("styledconv_down", scd, (x, st, nz3), scd(x, st, noise=nz3))

    rgb = R.ToRGB(16, 8); randomize(rgb)
    skip = torch.randn(2, 3, 4, 4)
    record("torgb_skip", rgb, (x, st, skip), rgb(x, st, skip))

    rgb1 = R.ToRGB(16, 8, upsample=False); randomize(rgb1)
    record("torgb_noskip", rgb1, (x, st), rgb1(x, st))

    sm = R.SMART_layer(16, 32, 3, 8); randomize(sm)
    record("smart", sm, (x, st, nz), sm(x, st, noise=nz))

    cl = R.ConvLayer(16, 32, 3); randomize(cl)
    record("convlayer", cl, (x,), cl(x))
    cld = R.ConvLayer(16, 32, 3, downsample=True); randomize(cld)
    record("convlayer_down", cld, (x,), cld(x))

    lcl = R.LargeConvLayer(3, 16, kernel_size=1); randomize(lcl)
    img = torch.randn(2, 3, 8, 8)
    record("largeconv_1x1", lcl, (img,), lcl(img))
    lcl3 = R.LargeConvLayer(16, 32, kernel_size=3); randomize(lcl3)
    record("largeconv_3x3", lcl3, (x,), lcl3(x))

    rb = R.ResBlock(16, 32); randomize(rb)
    record("resblock", rb, (x,), rb(x))

    el = R.EqualLinear(24, 16, lr_mul=0.01, activation="fused_lrelu")
    xin = torch.randn(3, 24)
    record("equallinear_act", el, (xin,), el(xin))
    el2 = R.EqualLinear(24, 16, bias_init=1)
    record("equallinear", el2, (xin,), el2(xin))

    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "layers.npz"), **out)
    print("layers:", len(names), "cases")


def gen_networks():
    """Whole networks at a small size. Weights are NOT stored (tens of MB): the test
    re-creates them with the same seed and constructor order and checks per-tensor
    checksums recorded here before comparing outputs."""
    out = {}
    size = 16
    torch.manual_seed(2024)
    net = R.Restoration_net(size, 512, 2, channel_multiplier=2).eval()
    dec = S.Generator(size, 512, 2, channel_multiplier=2).eval()
    torch.manual_seed(99)
    with torch.no_grad():
        for n_, p in list(net.named_parameters()) + list(dec.named_parameters()):
            if n_.endswith("noise.weight"):
                p.fill_(0.0)  # RNG-free outputs (reference default init, models/RestoreNet.py:562)
    b = 2
    low = torch.rand(b, 3, size, size) * 2 - 1
    codes = torch.randn(b, dec.n_latent, 512)
    z = torch.randn(b, 512)
    with torch.no_grad():
        img_dec, feats = dec([codes], input_is_latent=True, randomize_noise=True, return_features=True)
        restored = net(low, feats, codes, [z])
    out["size"] = np.array(size)
    out["low"], out["codes"], out["z"] = np_(low), np_(codes), np_(z)
    out["decoder_image"] = np_(img_dec)
    for i, f in enumerate(feats):
        out[f"decoder_feat{i}"] = np_(f)
    out["restored"] = np_(restored)
    for prefix, mod in (("net", net), ("dec", dec)):
        keys = list(mod.state_dict().keys())
        out[f"{prefix}.keys"] = np.array(keys)
        out[f"{prefix}.shapes"] = np.array([str(tuple(v.shape)) for v in mod.state_dict().values()])
        out[f"{prefix}.sums"] = np.array([float(v.double().sum()) for v in mod.state_dict().values()])
    np.savez_compressed(os.path.join(OUT, "networks.npz"), **out)
    print("networks: restored", tuple(restored.shape), "absmax", float(restored.abs().max()))

    # full-size state_dict key/shape manifest (drop-in contract)
    man = {}
    full = R.Restoration_net(512, 512, 8, channel_multiplier=2)
    man["restoration_net.keys"] = np.array(list(full.state_dict().keys()))
    man["restoration_net.shapes"] = np.array([str(tuple(v.shape)) for v in full.state_dict().values()])
    del full
    gen = S.Generator(1024, 512, 8, channel_multiplier=2)
    man["generator.keys"] = np.array(list(gen.state_dict().keys()))
    man["generator.shapes"] = np.array([str(tuple(v.shape)) for v in gen.state_dict().values()])
    del gen
    disc = R.Discriminator(512)
    man["discriminator.keys"] = np.array(list(disc.state_dict().keys()))
    man["discriminator.shapes"] = np.array([str(tuple(v.shape)) for v in disc.state_dict().values()])
    np.savez_compressed(os.path.join(OUT, "state_dict_manifest.npz"), **man)
    print("manifest written")


def gen_gradfix():
    """conv2d_gradfix closed set (op/conv2d_gradfix.py:134-223) and its consumers: plain / strided / dilated / transposed /
    grouped convolutions with first- and second-order gradients, the Discriminator forward (minibatch-stddev,
    models/RestoreNet.py:1244-1265) and the R1 penalty step (restoration_train.py:66-73, :200-216)."""
    from op import conv2d_gradfix as gf

    out, names = {}, []
    g = torch.Generator().manual_seed(77)

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    cases = [
        # name, transposed, x shape, w shape, bias, kwargs
        ("c3_s1_p1", False, (2, 16, 12, 12), (24, 16, 3, 3), True, dict(padding=1)),
        ("c3_s2_p0_odd", False, (2, 16, 13, 13), (24, 16, 3, 3), False, dict(stride=2)),
        ("c3_s2_p0_even", False, (2, 16, 12, 12), (24, 16, 3, 3), False, dict(stride=2)),
        ("c1_s2_skip", False, (2, 16, 11, 11), (8, 16, 1, 1), False, dict(stride=2)),
        ("c1_cin3_stem", False, (2, 3, 16, 16), (16, 3, 1, 1), True, dict()),
        ("c3_dil2", False, (2, 16, 16, 16), (8, 16, 3, 3), False, dict(padding=2, dilation=2)),
        ("c3_dil4", False, (1, 8, 16, 16), (8, 8, 3, 3), False, dict(padding=4, dilation=4)),
        ("c3_cin21_head", False, (4, 21, 4, 4), (16, 21, 3, 3), False, dict(padding=1)),
        ("c3_pad_asym", False, (1, 8, 9, 10), (8, 8, 3, 3), False, dict(padding=(0, 1))),
        ("t3_s2_p0", True, (2, 16, 6, 6), (16, 8, 3, 3), False, dict(stride=2)),
        ("t3_s2_p1_op1", True, (2, 8, 5, 5), (8, 8, 3, 3), True, dict(stride=2, padding=1, output_padding=1)),
        ("t3_s1_p1", True, (2, 8, 7, 7), (8, 16, 3, 3), False, dict(padding=1)),
        ("t1_s2", True, (1, 8, 5, 5), (8, 8, 1, 1), False, dict(stride=2)),
        ("g3_modconv_form", False, (1, 3 * 8, 10, 10), (3 * 12, 8, 3, 3), False, dict(padding=1, groups=3)),
        ("g3_transposed", True, (1, 3 * 8, 5, 5), (3 * 8, 12, 3, 3), False, dict(stride=2, groups=3)),
        ("g2_batch2", False, (2, 2 * 8, 6, 6), (2 * 8, 8, 3, 3), False, dict(padding=1, groups=2)),
    ]
    for name, tr, xs, ws, has_b, kw in cases:
        x = rnd(*xs).requires_grad_(True)
        w = (rnd(*ws) / (ws[1] * ws[2] * ws[3]) ** 0.5).requires_grad_(True)
        b = rnd(ws[1] * kw.get("groups", 1) if tr else ws[0]).requires_grad_(True) if has_b else None
        fn = gf.conv_transpose2d if tr else gf.conv2d
        y = fn(x, w, b, **kw)
        go = rnd(*y.shape).requires_grad_(True)
        ins = [x, w] + ([b] if has_b else [])
        grads = torch.autograd.grad(y, ins, go, create_graph=True)
        gx, gw = grads[0], grads[1]
        # second order: penalties on the input gradient (R1 form) and on the weight gradient
        pen_x = gx.pow(2).sum()
        ggw_from_x, ggo_from_x = torch.autograd.grad(pen_x, [w, go], retain_graph=True)
        pen_w = gw.pow(2).sum()
        ggx_from_w, ggo_from_w = torch.autograd.grad(pen_w, [x, go], retain_graph=True)
        out.update({f"{name}.x": np_(x), f"{name}.w": np_(w), f"{name}.go": np_(go), f"{name}.y": np_(y),
                    f"{name}.gx": np_(gx), f"{name}.gw": np_(gw), f"{name}.ggw_from_x": np_(ggw_from_x),
                    f"{name}.ggo_from_x": np_(ggo_from_x), f"{name}.ggx_from_w": np_(ggx_from_w),
                    f"{name}.ggo_from_w": np_(ggo_from_w), f"{name}.transposed": np.array(tr),
                    f"{name}.kw": np.array(repr(kw))})
        if has_b:
            out[f"{name}.b"], out[f"{name}.gb"] = np_(b), np_(grads[2])
        names.append(name)
    out["names"] = np.array(names)

    # Discriminator: forward at batch 8 (two minibatch-stddev groups) and the R1 step at batch 4
    size = 16
    torch.manual_seed(4242)
    disc = R.Discriminator(size)
    with torch.no_grad():
        for n_, p in disc.named_parameters():
            if n_.endswith("bias"):
                p.normal_(0, 0.2)
    img8 = torch.rand(8, 3, size, size, generator=g) * 2 - 1
    with torch.no_grad():
        out["disc.pred8"] = np_(disc(img8))
    out["disc.img8"] = np_(img8)
    real = (torch.rand(4, 3, size, size, generator=g) * 2 - 1).requires_grad_(True)
    pred = disc(real)
    with gf.no_weight_gradients():
        (grad_real,) = torch.autograd.grad(outputs=pred.sum(), inputs=real, create_graph=True)
    r1 = grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()
    disc.zero_grad()
    (10.0 / 2 * r1 * 16 + 0 * pred[0]).backward()
    out["disc.size"] = np.array(size)
    out["disc.real"], out["disc.pred"], out["disc.grad_real"], out["disc.r1"] = np_(real), np_(pred), np_(grad_real), np_(r1)
    keys = [k for k, _ in disc.named_parameters()]
    out["disc.param_keys"] = np.array(keys)
    out["disc.sd_keys"] = np.array(list(disc.state_dict().keys()))
    out["disc.sd_sums"] = np.array([float(v.double().sum()) for v in disc.state_dict().values()])
    out["disc.grad_absmax"] = np.array([float(p.grad.abs().max()) for _, p in disc.named_parameters()])
    out["disc.grad_sum"] = np.array([float(p.grad.double().sum()) for _, p in disc.named_parameters()])
    for k_, p in disc.named_parameters():          # full gradients of the small / structurally distinct parameters
        if p.numel() <= 40000 or k_ in ("encoder_convs.1.skip.0.weight",):
            out[f"disc.grad.{k_}"] = np_(p.grad)
    np.savez_compressed(os.path.join(OUT, "gradfix.npz"), **out)
    print("gradfix:", len(names), "conv cases; discriminator pred", np_(pred).ravel(), "r1", float(r1))


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["upfirdn2d", "fused_act", "modconv", "layers", "networks", "gradfix"]
    for w in which:
        globals()[f"gen_{w}"]()
