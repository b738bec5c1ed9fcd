"""Host-side algebra of the half-composed up-convolution (vsp_conv2d_up2h_bf16): the horizontally composed weights
(`compose_up2h_weights`) followed by the kernel's row decomposition and vertical 4-tap pass, emulated in fp64 on the CPU,
must equal conv_transpose2d(stride 2) -> Blur(4x4, pad 1) of models/RestoreNet.py:522-535 (oracle upfirdn2d)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.upfirdn2d_ref import upfirdn2d_ref
from vspbfr_b200.op.modconv import compose_up2h_weights


def emulate_up2h(x, wc, ky, cout):
    """x [B,Cin,H,W] fp64, wc [2*Cout,Cin,3,3] (row q*Cout+o, taps (kh, dx+1)) -> [B,Cout,2H,2W] exactly as the kernel
    assembles it: T_kh per input row, E[a] = T0[a] + T2[a-1], O[a] = T1[a], then the vertical taps."""
    b, cin, h, w = x.shape
    xp = F.pad(x, (1, 1, 0, 0))                                   # column halo (TMA zero fill)
    # T[kh][b, q*Cout+o, r, j] = sum_dx sum_i x[b,i,r,j+dx] wc[q*Cout+o, i, kh, dx+1]
    T = [F.conv2d(xp, wc[:, :, kh:kh + 1, :]) for kh in range(3)]
    zero = torch.zeros_like(T[0][:, :, :1])
    E = torch.cat([T[0], zero], 2) + torch.cat([zero, T[2]], 2)   # rows a = 0..H   (E[H] = T2[H-1])
    O = torch.cat([T[1], zero], 2)                                # rows a = 0..H   (O[H] = 0)
    Om1 = torch.cat([zero, O[:, :, :-1]], 2)                      # O[a-1]
    out = torch.zeros((b, cout, 2 * h, 2 * w), dtype=x.dtype)
    for q in range(2):
        e, o, om = (t[:, q * cout:(q + 1) * cout] for t in (E, O, Om1))
        r0 = ky[0] * om[:, :, :h] + ky[1] * e[:, :, :h] + ky[2] * o[:, :, :h] + ky[3] * e[:, :, 1:h + 1]
        r1 = ky[0] * e[:, :, :h] + ky[1] * o[:, :, :h] + ky[2] * e[:, :, 1:h + 1] + ky[3] * o[:, :, 1:h + 1]
        out[:, :, 0::2, q::2] = r0
        out[:, :, 1::2, q::2] = r1
    return out


@pytest.mark.parametrize("fy,fx", [([1, 3, 3, 1], [1, 3, 3, 1]), ([1, 2, 5, 3], [2, 1, 4, 7])])
def test_half_composed_up_convolution_equals_transposed_conv_then_blur(fy, fx):
    torch.manual_seed(3)
    b, cin, cout, h, w = 2, 5, 3, 6, 7
    x = torch.randn(b, cin, h, w, dtype=torch.float64)
    wt = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    fy, fx = torch.tensor(fy, dtype=torch.float64), torch.tensor(fx, dtype=torch.float64)
    k = torch.outer(fy, fx)
    k = k / k.sum() * 4
    fy_n, fx_n = fy / fy.sum() * 2, fx / fx.sum() * 2             # outer(fy_n, fx_n) == k
    # reference: conv_transpose2d with weight [in, out, kh, kw], no flip, then the blur with pad (1, 1)
    z = F.conv_transpose2d(x, wt.transpose(0, 1).contiguous(), stride=2)
    ref = torch.from_numpy(upfirdn2d_ref(z.numpy(), k.numpy(), up=1, down=1, pad=(1, 1)))
    wc = compose_up2h_weights(wt, fx_n.tolist()).double()
    # the fp32 round trip of the composite costs ~1e-7; recompute in fp64 for the exact identity
    wc64 = torch.zeros((2, cout, cin, 3, 3), dtype=torch.float64)
    for q in range(2):
        for v in range(4):
            for kw in range(3):
                t = q + v - 1 - kw
                if t % 2 == 0 and -1 <= t // 2 <= 1:
                    wc64[q, :, :, :, t // 2 + 1] += fx_n[3 - v] * wt[:, :, :, kw]
    wc64 = wc64.reshape(2 * cout, cin, 3, 3)
    assert float((wc - wc64).abs().max()) < 1e-6 * float(wc64.abs().max())
    ky = [float(fy_n[3 - u]) for u in range(4)]
    got = emulate_up2h(x, wc64, ky, cout)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-11, atol=1e-11)
