#!/usr/bin/env python
"""N = 256 convolutions: single-CTA kernel vs the CTA-pair (cta_group::2) kernel — run once per setting of VSP_CONV_PAIR
(the switch is read once per process).  Prints JSON rows: shape, us, TFLOP/s, checksum (the two kernels must agree)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc

def main():
    dev = "cuda"
    rows = []
    for (b, cin, cout, h, per_sample) in ((8, 512, 512, 64, True), (32, 512, 512, 64, True), (32, 256, 256, 128, False),
                                          (32, 512, 512, 32, False), (3, 256, 256, 40, True), (32, 128, 128, 256, False), (32, 128, 128, 256, True),
                                          (32, 256, 128, 64, True), (5, 128, 128, 48, True)):
        g = torch.Generator().manual_seed(b + cin + h)
        sets = []
        for _ in range(4):
            x = torch.randn(b, h, h, cin, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(dev)
            s = (torch.randn(b, cin, generator=g) * 0.3 + 1).to(dev) if per_sample else None
            wq, _ = mc.pack_weights(wt, s)
            sets.append((x, wq))
        rs = (torch.rand(b, cout, generator=g) + 0.5).to(dev)
        bias = torch.randn(cout, generator=g).to(dev)
        epi = mc.make_epilogue(row_scale=rs, bias=bias, act=3, alpha=0.2, scale=math.sqrt(2))
        outs = [torch.empty(b, h, h, cout, dtype=torch.bfloat16, device=dev) for _ in sets]
        def run():
            for (x, wq), o in zip(sets, outs):
                mc.conv_fprop(x, wq, cout, 3, 3, 1, 1, 1, epi=epi, out=o, out_nhwc=True)
        run(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); run(); run(); run(); e0.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(e0) / 12 * 1e-3)
        fl = 2.0 * b * h * h * cout * cin * 9
        rows.append({"shape": f"b{b} {cin}->{cout} {h}x{h} {'g=b' if per_sample else 'g=1'}", "pair": os.environ.get("VSP_CONV_PAIR", "1"),
                     "us": best * 1e6, "TFLOPs": fl / best / 1e12, "checksum": float(outs[0].float().double().sum()),
                     "absmax": float(outs[0].float().abs().max())})
        print(json.dumps(rows[-1]), flush=True)

if __name__ == "__main__":
    main()
