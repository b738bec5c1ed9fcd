"""Precision-floor experiment (test infrastructure, CPU only): the fp32 CPU oracle of the hot path re-run with bf16 rounding
injected at the points where the fused sm_100a pipeline rounds (conv operands -> bf16, fp32 accumulation, fp32 epilogue,
bf16 activation stores), selectable per site / per network stage.  It answers "what max-abs error does ANY bf16-operand
implementation have on this random-init network at BASELINE size, and which sites dominate it" without a GPU.

    python tests/sim_bf16_floor.py [policy ...]      policies: see POLICIES below
Results are summarised in DESIGN.md §2.
"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import network_ref as nr            # noqa: E402
from oracle.upfirdn2d_ref import upfirdn2d_native_port  # noqa: E402

SQRT2 = math.sqrt(2.0)


class Policy:
    """Which sites round to bf16.  ``w``: conv weights, ``x``: conv input operands, ``store``: activation stores between
    layers; ``stages``: set of stage names the policy applies to (None = all); ``split_w``/``split_x``: keep a second bf16
    term (hi + lo) for that operand, i.e. a 2-pass / 3-pass tensor-core product."""

    def __init__(self, w=True, x=True, store=True, stages=None, split_w=False, split_x=False, split_stages=None,
                 fp32_store_stages=None, w_fp16=False):
        self.w_fp16 = w_fp16            # weight operand in fp16 (11-bit significand) instead of bf16 (8-bit)
        self.w, self.x, self.store, self.stages = w, x, store, stages
        self.split_w, self.split_x, self.split_stages = split_w, split_x, split_stages
        self.fp32_store_stages = fp32_store_stages or set()
        self.stage = ""

    def _on(self):
        return self.stages is None or any(self.stage.startswith(s) for s in self.stages)

    def _split(self):
        return self.split_stages is None or any(self.stage.startswith(s) for s in self.split_stages)

    @staticmethod
    def _r(t, split):
        hi = t.bfloat16().float()
        if split:
            hi = hi + (t - hi).bfloat16().float()
        return hi

    def qw(self, t):
        if self.w_fp16 and self.w and self._on():
            return t.half().float()
        return self._r(t, self.split_w and self._split()) if (self.w and self._on()) else t

    def qx(self, t):
        return self._r(t, self.split_x and self._split()) if (self.x and self._on()) else t

    def qs(self, t):
        if any(self.stage.startswith(s) for s in self.fp32_store_stages):
            return t
        return self._r(t, self.split_x and self._split()) if (self.store and self._on()) else t


P = Policy()


def modconv(sd, p, x, s, demod=True, up=False, down=False, dilation=1):
    """d * conv(q(x), q(wscale * W * s)) with fp32 accumulation — the fused pipeline's form (demod in the epilogue)."""
    weight = sd[p + "weight"]
    b, cin, h, w_ = x.shape
    _, cout, _, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    wmod = scale * weight * s.reshape(b, 1, cin, 1, 1)
    d = torch.rsqrt(wmod.pow(2).sum(dim=(2, 3, 4)) + 1e-8) if demod else None
    wq = P.qw(wmod)
    xq = P.qx(x)
    outs = []
    for i in range(b):
        xi, wi = xq[i:i + 1], wq[i]
        if up:
            yi = F.conv_transpose2d(xi, wi.transpose(0, 1), stride=2, padding=0)
            yi = upfirdn2d_native_port(yi, nr._blur_kernel(4.0), 1, 1, (1, 1))
        elif down:
            xb = P.qs(upfirdn2d_native_port(xi, nr._blur_kernel(1.0), 1, 1, (2, 2)))
            yi = F.conv2d(xb, wi, stride=2)
        else:
            yi = F.conv2d(xi, wi, padding=((k - 1) * dilation) // 2, dilation=dilation)
        outs.append(yi)
    y = torch.cat(outs, 0)
    if d is not None:
        y = y * d.reshape(b, cout, 1, 1)
    return y


def _nz(sd, p, y, noise):
    """NoiseInjection with an explicit image (None: the weight is taken as 0)."""
    return y if noise is None else y + sd[p + "noise.weight"] * noise


def styled_conv(sd, p, x, style, up=False, down=False, residuals=(), noise=None):
    s = nr._equal_linear(sd, p + "conv.modulation.", style)
    y = modconv(sd, p + "conv.", x, s, up=up, down=down)
    y = nr._lrelu(_nz(sd, p, y, noise), sd[p + "activate.bias"])
    for r in residuals:
        y = y + r
    return P.qs(y)


def to_rgb(sd, p, x, style, skip=None):
    s = nr._equal_linear(sd, p + "conv.modulation.", style)
    w = sd[p + "conv.weight"]
    b, cin = s.shape
    wm = (w[0, :, :, 0, 0] / math.sqrt(cin))[None] * s[:, None, :]        # fp32 weights in the ToRGB kernel
    y = torch.einsum("boc,bchw->bohw", wm, x) + sd[p + "bias"]
    if skip is not None:
        y = y + upfirdn2d_native_port(skip, nr._blur_kernel(4.0), 2, 1, (2, 1))
    return y


def smart(sd, p, x, style, rates=(1, 2, 4, 8), noise=None):
    s = nr._equal_linear(sd, p + "modulation.", style)
    outs = [P.qs(modconv(sd, f"{p}ModulatedConv2ds.{j}.", x, s, dilation=r)) for j, r in enumerate(rates)]
    w = sd[p + "fusion.0.weight"]
    y = F.conv2d(P.qx(torch.cat(outs, 1)), P.qw(w / math.sqrt(w.shape[1] * 9)), padding=1)
    y = nr._lrelu(y, sd[p + "fusion.1.bias"])
    return P.qs(nr._lrelu(_nz(sd, p, y, noise), sd[p + "activate.bias"]))


def large_conv(sd, p, x, k, rates=(1, 2, 4, 8)):
    outs = []
    xq = P.qx(x)
    for j, r in enumerate(rates):
        w = sd[f"{p}dilated_convs.{j}.weight"]
        outs.append(P.qs(F.conv2d(xq, P.qw(w / math.sqrt(w.shape[1] * k * k)), padding=((k - 1) * r) // 2, dilation=r)))
    w = sd[p + "fusion.0.weight"]
    y = nr._lrelu(F.conv2d(P.qx(torch.cat(outs, 1)), P.qw(w / math.sqrt(w.shape[1]))), sd[p + "fusion.1.bias"])
    return P.qs(nr._lrelu(y, sd[p + "activate.bias"]))


@torch.no_grad()
def generator(sd, codes, size, noise=None):
    log_size = int(math.log2(size))
    b = codes.shape[0]
    noise = noise or [None] * (2 * (log_size - 2) + 1)
    P.stage = "dec4"
    out = P.qs(sd["input.input"].repeat(b, 1, 1, 1))
    out = styled_conv(sd, "conv1.", out, codes[:, 0], noise=noise[0])
    skip = to_rgb(sd, "to_rgb1.", out, codes[:, 1])
    feats = [out]
    i = 1
    for lvl in range(log_size - 2):
        P.stage = f"dec{2 ** (lvl + 3)}"
        out = styled_conv(sd, f"convs.{2 * lvl}.", out, codes[:, i], up=True, noise=noise[2 * lvl + 1])
        feats.append(out)
        out = styled_conv(sd, f"convs.{2 * lvl + 1}.", out, codes[:, i + 1], noise=noise[2 * lvl + 2])
        skip = to_rgb(sd, f"to_rgbs.{lvl}.", out, codes[:, i + 2], skip)
        i += 2
    return skip, feats


@torch.no_grad()
def restoration(sd, images, de_feats, pre_styles, z, size, n_mlp, noise=None):
    """``noise``: None or {"encoder": [...], "decoder": [...]} (explicit per-layer images, as oracle.restoration_ref)."""
    log_size = int(math.log2(size))
    n_latent = log_size * 2 - 2
    n_enc, n_dec = [None] * (2 * (log_size - 2)), [None] * (2 * (log_size - 2) + 1)
    if noise is not None:
        n_enc, n_dec = list(noise["encoder"]), list(noise["decoder"])
    b = images.shape[0]
    w_noise = nr._style_mlp(sd, z, n_mlp).unsqueeze(1).repeat(1, n_latent, 1)
    latent = torch.cat([pre_styles[:, :n_latent], w_noise], dim=-1)
    lat_rev = torch.flip(latent, dims=[1])
    P.stage = f"enc{size}"
    out = large_conv(sd, "down_from_big.", images, 1)
    features = []
    for lvl in range(log_size - 2):
        ii = 2 * lvl
        P.stage = f"enc{size >> lvl}"
        out = smart(sd, f"encoder_convs.{ii}.", out, lat_rev[:, ii], noise=n_enc[ii])
        features.append(out)
        out = styled_conv(sd, f"encoder_convs.{ii + 1}.", out, lat_rev[:, ii], down=True, noise=n_enc[ii + 1])
    P.stage = "enc4"
    out = large_conv(sd, "final_layer.", out, 3)
    x_global = nr._equal_linear(sd, "final_linear.0.", out.reshape(b, -1), act=True)
    early = nr._equal_linear(sd, "final_transfer.", x_global, act=True).reshape(b, -1, 4, 4)
    features.append(P.qs(out + early))
    features = features[::-1]

    def sty(i):
        return torch.cat([latent[:, i], x_global], dim=1)

    P.stage = "res4"
    out = smart(sd, "conv1.", features[0], sty(0), noise=n_dec[0])
    skip = to_rgb(sd, "to_rgb1.", out, sty(1))
    i = 1
    for lvl in range(log_size - 2):
        P.stage = f"res{2 ** (lvl + 3)}"
        level = (i + 1) // 2
        out = styled_conv(sd, f"convs.{2 * lvl}.", out, sty(i), up=True, residuals=(features[level], de_feats[level]),
                          noise=n_dec[2 * lvl + 1])
        out = smart(sd, f"convs.{2 * lvl + 1}.", out, sty(i + 1), noise=n_dec[2 * lvl + 2])
        skip = to_rgb(sd, f"to_rgbs.{lvl}.", out, sty(i + 2), skip)
        i += 2
    return skip


POLICIES = {
    "fp32": lambda: Policy(w=False, x=False, store=False),
    "bf16_all": lambda: Policy(),
    "w_only": lambda: Policy(x=False, store=False),
    "act_only": lambda: Policy(w=False),
    "dec_only": lambda: Policy(stages={"dec"}),
    "enc_only": lambda: Policy(stages={"enc"}),
    "res_only": lambda: Policy(stages={"res"}),
    "lowres_only": lambda: Policy(stages={f"{n}{r}" for n in ("dec", "enc", "res") for r in (4, 8, 16, 32)}),
    "hires_only": lambda: Policy(stages={f"{n}{r}" for n in ("dec", "enc", "res") for r in (256, 512, 1024)}),
    "mid_only": lambda: Policy(stages={f"{n}{r}" for n in ("dec", "enc", "res") for r in (64, 128)}),
    "split_w": lambda: Policy(split_w=True),
    "split_x": lambda: Policy(split_x=True),
    "split_both_lowres": lambda: Policy(split_w=True, split_x=True,
                                        split_stages={f"{n}{r}" for n in ("dec", "enc", "res") for r in (4, 8, 16, 32)}),
    "split_both_le64": lambda: Policy(split_w=True, split_x=True,
                                      split_stages={f"{n}{r}" for n in ("dec", "enc", "res") for r in (4, 8, 16, 32, 64)}),
}


def restore_faces_rounded(policy, net_sd, dec_sd, low, codes, z, size, dec_size, n_mlp, dec_noise=None, net_noise=None):
    """The hot path (as oracle.restore_faces_ref) under a rounding policy name: ``restored`` only.  GPU tests use
    ``"bf16_all"`` on their own inputs as the precision floor of an all-bf16-operand implementation."""
    global P
    saved, P = P, POLICIES[policy]()
    try:
        _, feats = generator(dec_sd, codes, dec_size, noise=dec_noise)
        return restoration(net_sd, low, feats, codes, z, size, n_mlp, noise=net_noise)
    finally:
        P = saved


def _lv(names, rs):
    return {f"{n}{r}" for n in names for r in rs}


POLICIES.update({
    "enc_lowres_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8, 16, 32])),
    "enc_lowres_exact_encw": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8, 16, 32]),
                                            ),
    "all_lowres_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32])),
    "all_le64_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32, 64])),
    "enc_le128_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8, 16, 32, 64, 128])),
    "enc_exact": lambda: Policy(split_w=True, split_x=True, split_stages={"enc"}),
    "fp16_w": lambda: Policy(w_fp16=True),
    "fp16_w_enc_split_x": lambda: Policy(w_fp16=True, split_x=True, split_stages={"enc"}),
    "fp16_w_le32_split_x": lambda: Policy(w_fp16=True, split_x=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32])),
    "enc_le16_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8, 16])),
    "enc_le8_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8])),
    "enc_le16_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc"], [4, 8, 16])),
    "enc_le32_split_x": lambda: Policy(split_x=True, split_stages=_lv(["enc"], [4, 8, 16, 32])),
    "enc_le64_exact": lambda: Policy(split_w=True, split_x=True, split_stages=_lv(["enc"], [4, 8, 16, 32, 64])),
    "enc_split_w": lambda: Policy(split_w=True, split_stages={"enc"}),
    "enc_le64_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc"], [4, 8, 16, 32, 64])),
    "enc_le32_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc"], [4, 8, 16, 32])),
    "le32_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32])),
    "le64_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32, 64])),
    "le128_split_w": lambda: Policy(split_w=True, split_stages=_lv(["enc", "dec", "res"], [4, 8, 16, 32, 64, 128])),
})


def main():
    global P
    names = sys.argv[1:] or ["bf16_all"]
    n_img = int(os.environ.get("SIM_IMAGES", "1"))
    size, dec_size = int(os.environ.get("SIM_SIZE", "512")), int(os.environ.get("SIM_DEC_SIZE", "1024"))
    from vspbfr_b200.restorenet import Restoration_net
    from vspbfr_b200.stylegan2 import Generator
    seed = int(os.environ.get("SIM_SEED", "11"))
    torch.manual_seed(seed)
    net = Restoration_net(size, 512, 8, channel_multiplier=2).eval()
    dec = Generator(dec_size, 512, 8, channel_multiplier=2).eval()
    nsd, dsd = net.state_dict(), dec.state_dict()
    g = torch.Generator().manual_seed(seed + 1)
    low = torch.rand(2, 3, size, size, generator=g)[:n_img] * 2 - 1
    codes = torch.randn(2, 18, 512, generator=g)[:n_img]
    z = torch.randn(2, 512, generator=g)[:n_img]

    def run():
        _, feats = generator(dsd, codes, dec_size)
        return restoration(nsd, low, feats, codes, z, size, 8)

    P = POLICIES["fp32"]()
    t0 = time.time()
    want = run()
    peak = float(want.max() - want.min())
    print(f"fp32 reference: range {peak:.1f} ({time.time() - t0:.1f} s)", flush=True)
    for name in names:
        P = POLICIES[name]()
        got = run()
        err = (got - want).abs()
        mse = float(((got - want) ** 2).mean())
        print(f"{name:20s} max-abs {float(err.max()) / peak:.3e} of range   psnr {10 * math.log10(peak * peak / mse):.1f} dB   "
              f"rms/range {math.sqrt(mse) / peak:.2e}", flush=True)


if __name__ == "__main__":
    main()
