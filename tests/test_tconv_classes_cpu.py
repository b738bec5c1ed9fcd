"""Host-side algebra of the one-launch stride-2 transposed convolution (class mode of vsp_conv_transpose2d_s2_bf16,
csrc/conv_sm100.cu): the tap table the dispatcher builds — per output parity class (pa, pb) the taps (kh, kw) with
kh = pa, kw = pb (mod 2), each reading the activation at shift (-(kh - pa) / 2, -(kw - pb) / 2) — emulated in fp64 over the
H x W grid plus the last output row / column it leaves to the thin launches, must equal conv_transpose2d(stride 2, pad 0)
of models/RestoreNet.py:522-529 on every output element."""
import pytest
import torch
import torch.nn.functional as F


def class_taps():
    """ClassTaps as conv_sm100.cu builds it: class c = pa * 2 + pb -> [(shift index, weight tap)], shift index bit 1 = row
    shift -1, bit 0 = column shift -1."""
    table = {}
    for pa in range(2):
        for pb in range(2):
            table[pa * 2 + pb] = [(((i - pa) // 2) * 2 + (j - pb) // 2, i * 3 + j)
                                  for i in range(pa, 3, 2) for j in range(pb, 3, 2)]
    return table


SHIFT = {0: (0, 0), 1: (0, -1), 2: (-1, 0), 3: (-1, -1)}       # dy4 / dx4 of the dispatcher


def emulate(x, w):
    """x [B,Cin,H,W], w [Cout,Cin,3,3] -> [B,Cout,2H+1,2W+1] assembled class by class."""
    b, cin, h, wd = x.shape
    cout = w.shape[0]
    out = torch.zeros(b, cout, 2 * h + 1, 2 * wd + 1, dtype=x.dtype)
    xp = F.pad(x, (1, 1, 1, 1))                                    # out-of-range reads are zero (TMA fill)
    for c, taps in class_taps().items():
        pa, pb = c >> 1, c & 1
        rows, cols = (2 * h + 1 - pa + 1) // 2, (2 * wd + 1 - pb + 1) // 2     # H+1 / W+1 for the even classes
        acc = torch.zeros(b, cout, rows, cols, dtype=x.dtype)
        for shift, tap in taps:
            dy, dx = SHIFT[shift]
            kh, kw = divmod(tap, 3)
            # pixel (a, b) of the class grid reads x[a + dy, b + dx]
            win = xp[:, :, 1 + dy:1 + dy + rows, 1 + dx:1 + dx + cols]
            acc += torch.einsum("bihw,oi->bohw", win, w[:, :, kh, kw])
        out[:, :, pa::2, pb::2] = acc
    return out


def test_tap_table_counts_and_weight_coverage():
    t = class_taps()
    assert [len(t[c]) for c in range(4)] == [4, 2, 2, 1]           # 9 taps in all: 1x the layer's FLOPs
    assert sorted(tap for taps in t.values() for _, tap in taps) == list(range(9))


@pytest.mark.parametrize("h,w", [(4, 5), (1, 1), (7, 3)])
def test_class_decomposition_equals_conv_transpose2d(h, w):
    torch.manual_seed(h * 10 + w)
    x = torch.randn(2, 3, h, w, dtype=torch.float64)
    wt = torch.randn(4, 3, 3, 3, dtype=torch.float64)
    want = F.conv_transpose2d(x, wt.transpose(0, 1), None, stride=2, padding=0)
    got = emulate(x, wt)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-12


def test_tile_rotation_gives_every_cta_every_class():
    """tile_n_index (conv_common.cuh): n_i = (tile % tiles_n + tile / tiles_n) % tiles_n — a persistent CTA's tiles
    (stride = grid size) must cycle through the classes so that their 4 : 2 : 2 : 1 work balances."""
    work = [4, 2, 2, 1]
    for grid, tiles_n, m_tiles in ((148, 4, 32 * 128), (74, 4, 32 * 32), (74, 8, 32 * 16)):
        per_cta = []
        for j in range(grid):
            tot, t = 0, j
            while t < m_tiles * tiles_n:
                n_i = (t % tiles_n + t // tiles_n) % tiles_n
                tot += work[n_i * 4 // tiles_n]
                t += grid
            per_cta.append(tot)
        assert max(per_cta) <= 1.08 * min(per_cta), (grid, tiles_n, min(per_cta), max(per_cta))
