"""HBM yardsticks for write-heavy kernels: copy (1:1), fill (0:1) and a 1:4 read:write expansion, timed with 20 launches
per event pair (the event clock has ~2 us granularity) and an L2 flush between batches."""
import torch
dev = "cuda"
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3 / reps)
    return best

n = 64 * 1024 * 1024            # 256 MB of fp32: larger than L2 (126 MB) so repeated launches stay HBM-bound
a = torch.randn(n, device=dev); b = torch.empty_like(a)
t = timeit(lambda: b.copy_(a)); print(f"copy 256MB->256MB : {2*n*4/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
t = timeit(lambda: b.fill_(1.0)); print(f"fill 256MB        : {n*4/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
x = torch.randn(4, 512, 64, 64, device=dev)
t = timeit(lambda: torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"))
print(f"nearest x2 [4,512,64,64] (1:4 read:write, 168 MB): {x.numel()*4*5/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
xl = torch.randn(16, 512, 64, 64, device=dev)
t = timeit(lambda: torch.nn.functional.interpolate(xl, scale_factor=2, mode="nearest"))
print(f"nearest x2 [16,512,64,64] (671 MB): {xl.numel()*4*5/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200.op.upfirdn2d import upfirdn2d_raw
k1 = torch.tensor([1.0, 3.0, 3.0, 1.0], device=dev); k = torch.outer(k1, k1) / 16
t = timeit(lambda: upfirdn2d_raw(x, k, (2, 2), (1, 1), (2, 1, 2, 1)))
print(f"upfirdn2d up2 [4,512,64,64] (168 MB): {x.numel()*4*5/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
t = timeit(lambda: upfirdn2d_raw(xl, k, (2, 2), (1, 1), (2, 1, 2, 1)))
print(f"upfirdn2d up2 [16,512,64,64] (671 MB): {xl.numel()*4*5/t/1e9:7.0f} GB/s ({t*1e6:.1f} us)")
