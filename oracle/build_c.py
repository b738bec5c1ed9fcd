"""Build recipe of the plain-C oracle (oracle/c/vsp_oracle.c -> oracle/_build/libvsp_oracle_c.so) and its ctypes binding.
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  `python -m oracle.build_c` or `build()` from __graft_entry__.build()."""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c", "vsp_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "libvsp_oracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        if os.path.exists(LIB):         # a prebuilt checker travelled with the tree: use it rather than fail the build step
            return LIB
        raise RuntimeError("oracle.build_c: no C compiler (gcc/cc) on PATH")
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + f".{os.getpid()}.tmp"   # atomic replace: several ranks / test workers may build at once
    subprocess.run([cc, "-O2", "-std=c99", "-shared", "-fPIC", "-o", tmp, SRC], check=True)
    os.replace(tmp, LIB)
    return LIB


def load():
    """ctypes handle with argument types set (builds on first use)."""
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        i64, f32p, ci, cf = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.oracle_upfirdn2d_out_size.restype = i64
        lib.oracle_upfirdn2d_out_size.argtypes = [i64, ci, ci, ci, ci, ci]
        lib.oracle_upfirdn2d_f32.restype = ci
        lib.oracle_upfirdn2d_f32.argtypes = [f32p, f32p, f32p, i64, i64, i64] + [ci] * 10
        lib.oracle_bias_act_f32.restype = ci
        lib.oracle_bias_act_f32.argtypes = [f32p, f32p, f32p, f32p, i64, i64, i64, ci, ci, cf, cf]
        _lib = lib
    return _lib


def upfirdn2d_c(x, kernel, up=1, down=1, pad=(0, 0)):
    """numpy front of oracle_upfirdn2d_f32: x [N,C,H,W] float32 -> [N,C,H',W'] (same argument forms as upfirdn2d_ref)."""
    import numpy as np

    x = np.ascontiguousarray(x, dtype=np.float32)
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    up_x, up_y = (up, up) if isinstance(up, int) else (int(up[0]), int(up[1]))
    down_x, down_y = (down, down) if isinstance(down, int) else (int(down[0]), int(down[1]))
    pad = tuple(int(p) for p in pad)
    px0, px1, py0, py1 = (pad[0], pad[1], pad[0], pad[1]) if len(pad) == 2 else pad
    n, c, h, w = x.shape
    lib = load()
    oh = lib.oracle_upfirdn2d_out_size(h, k.shape[0], up_y, down_y, py0, py1)
    ow = lib.oracle_upfirdn2d_out_size(w, k.shape[1], up_x, down_x, px0, px1)
    y = np.zeros((n, c, max(oh, 0), max(ow, 0)), dtype=np.float32)
    rc = lib.oracle_upfirdn2d_f32(x.ctypes.data, k.ctypes.data, y.ctypes.data, n * c, h, w, k.shape[0], k.shape[1], up_x, up_y,
                                  down_x, down_y, px0, px1, py0, py1)
    if rc:
        raise ValueError("oracle_upfirdn2d_f32: bad argument")
    return y


def bias_act_c(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """numpy front of oracle_bias_act_f32: bias broadcast on dim 1 (op/fused_bias_act_kernel.cu:84-88)."""
    import numpy as np

    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    b = np.ascontiguousarray(bias, dtype=np.float32) if bias is not None and np.size(bias) else None
    r = np.ascontiguousarray(ref, dtype=np.float32) if ref is not None and np.size(ref) else None
    step_b = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    size_b = int(b.size) if b is not None else 1
    rc = load().oracle_bias_act_f32(x.ctypes.data, b.ctypes.data if b is not None else None, r.ctypes.data if r is not None else None,
                                    y.ctypes.data, x.size, step_b, size_b, int(act), int(grad), float(alpha), float(scale))
    if rc:
        raise ValueError("oracle_bias_act_f32: bad argument")
    return y


if __name__ == "__main__":
    print(build(force=True))
