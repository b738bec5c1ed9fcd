"""Time one conv shape (CUDA events, 10 iterations): prof_conv_time.py b cin cout h k dil nhwc [noise]"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op import modconv as mc
b, cin, cout, h, k, dil, nhwc = [int(v) for v in sys.argv[1:8]]
x = torch.randn(b, cin, h, h, device="cuda")
w = torch.randn(cout, cin, k, k, device="cuda")
s = torch.randn(b, cin, device="cuda") * 0.3 + 1
xq = mc.nchw_to_nhwc_bf16(x)
wq, d = mc.pack_weights(w, s, wscale=1 / math.sqrt(cin * k * k), want_demod=True)
epi = mc.make_epilogue(row_scale=d)
out = None
for _ in range(3):
    out = mc.conv_fprop(xq, wq, cout, k, k, 1, (k - 1) * dil // 2, dil, epi=epi, out_nhwc=bool(nhwc))
torch.cuda.synchronize()
s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(10):
    mc.conv_fprop(xq, wq, cout, k, k, 1, (k - 1) * dil // 2, dil, epi=epi, out=out, out_nhwc=bool(nhwc))
e0.record(); torch.cuda.synchronize()
print(f"DBG={os.environ.get('VSP_FOLD_DBG','0')} b{b} {cin}->{cout} k{k} d{dil} {h}: {s0.elapsed_time(e0)*100:.1f} us")
