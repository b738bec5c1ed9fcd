import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vspbfr_b200 import _lib
lib = _lib.load()
x = torch.randn(1, 1, 256, 256, device="cuda")
buf = (ctypes.c_ubyte * 128)()
f = lib.vsp_debug_tma_3d_f32
f.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
rc = f(x.data_ptr(), 256, 256, 1, 132, 19, 1, buf)
print("rc", rc, lib.vsp_last_error())
mine = bytes(buf)
print("mine ", mine.hex())
from cuda.bindings import driver as drv
err, = drv.cuInit(0)
res = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x.data_ptr(),
    [256, 256, 1], [256 * 4, 256 * 256 * 4], [132, 19, 1], [1, 1, 1],
    drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_NONE,
    drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B, drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
print("cuda-python err", res[0])
tm = res[1]
theirs = bytes(ctypes.string_at(int(tm.getPtr()), 128))
print("theirs", theirs.hex())
print("equal", mine == theirs)
