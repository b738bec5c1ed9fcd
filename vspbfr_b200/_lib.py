"""Build + ctypes binding of the C-ABI library (include/vsp_b200.h).

The shared object is built in-tree (``vspbfr_b200/libvsp_b200.so``) with
``nvcc -gencode arch=compute_100a,code=sm_100a``; nothing here imports torch
extensions or pybind — the boundary is plain C (SURVEY.md §8 b).  There is no
CPU fallback: if the library is missing and cannot be built, or a kernel call
returns non-zero, a ``RuntimeError`` is raised (the reference raises the same
type through TORCH_CHECK, op/upfirdn2d.cpp:9-15).
"""
from __future__ import annotations

import ctypes
import fcntl
import os
import shutil
import subprocess
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libvsp_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

SOURCES = [
    "c_api.cu",
    "upfirdn2d_sm100.cu",
    "bias_act_sm100.cu",
    "layout_sm100.cu",
    "conv_sm100.cu",
    "conv_ring_sm100.cu",
    "conv_up2h_sm100.cu",
    "wgrad_sm100.cu",
    "torgb_sm100.cu",
    "linear_sm100.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]

_lock = threading.Lock()
_lib = None


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "vsp_b200.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libvsp_b200.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    build_dir = os.path.join(_HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    # one builder at a time across threads AND processes (every torchrun rank calls load() on a stale checkout): the
    # losers of the race wait on the file lock, then find the library fresh and return
    with _lock, open(os.path.join(build_dir, ".lock"), "w") as lock_file:
        fcntl.flock(lock_file, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB_PATH
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            raise RuntimeError("vspbfr_b200: nvcc not found and libvsp_b200.so is missing or stale")
        objs = []
        procs = []
        compile_flags = [f for f in NVCC_FLAGS if f != "--shared"]
        for src in _sources():
            obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
            hdrs.append(os.path.join(INCLUDE, "vsp_b200.h"))
            if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                    and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in hdrs)):
                continue
            cmd = [nvcc, *compile_flags, "-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        log = []
        for src, p in procs:
            out, _ = p.communicate()
            log.append(out)
            if p.returncode != 0:
                raise RuntimeError(f"vspbfr_b200: nvcc failed on {src}:\n{out}")
        tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
        cmd = [nvcc, "--shared", "-o", tmp, *objs, "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"vspbfr_b200: link failed:\n{r.stdout}")
        os.replace(tmp, LIB_PATH)
        with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
        return LIB_PATH


class ConvEpilogue(Structure):
    """Mirror of ``vsp_conv_epilogue`` (include/vsp_b200.h)."""

    _fields_ = [
        ("row_scale", c_void_p),
        ("noise", c_void_p),
        ("noise_bstride", c_int64),
        ("noise_weight", c_float),
        ("noise_weight_dev", c_void_p),
        ("bias", c_void_p),
        ("pre_bias", c_void_p),
        ("pre_act", c_int),
        ("act", c_int),
        ("alpha", c_float),
        ("scale", c_float),
        ("residual", c_void_p),
        ("residual2", c_void_p),
        ("alpha_vec", c_void_p),
    ]


class LinearDesc(Structure):
    """Mirror of ``vsp_linear_desc`` (include/vsp_b200.h)."""

    _fields_ = [
        ("w", c_void_p),
        ("bias", c_void_p),
        ("x_off", c_int64),
        ("y_off", c_int64),
        ("in_dim", ctypes.c_int32),
        ("out_dim", ctypes.c_int32),
        ("wscale", c_float),
        ("bscale", c_float),
        ("x_bstride", c_int64),
        ("act", ctypes.c_int32),
        ("alpha", c_float),
        ("gain", c_float),
        ("pad_", ctypes.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/vsp_b200.h declares.
SIGNATURES = {
    "vsp_version": (c_int, []),
    "vsp_last_error": (c_char_p, []),
    "vsp_launch_count": (c_int64, []),
    "vsp_upfirdn2d_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                  c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int,
                                  c_void_p, c_int64, c_int, c_float, c_float, c_void_p]),
    "vsp_upfirdn2d_out_size": (c_int64, [c_int64, c_int, c_int, c_int, c_int, c_int]),
    "vsp_upfirdn2d_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                        c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int, c_int, POINTER(ConvEpilogue), c_void_p]),
    "vsp_bias_act_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                 c_int, c_int, c_float, c_float, c_void_p]),
    "vsp_bias_act_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                     c_float, c_float, c_void_p]),
    "vsp_nchw_f32_to_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vsp_nchw_f32_to_nhwc_split3_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "vsp_nhwc_bf16_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vsp_quantize_nchw_f32_to_hwc_u8": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_float, c_void_p]),
    "vsp_nchw_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "vsp_modulate_weights_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_int64, c_int64, c_int,
                                          c_float, c_float, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "vsp_weight_sumsq_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "vsp_conv2d_fprop_bf16": (c_int, [c_void_p, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                      c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int64, c_int64, POINTER(ConvEpilogue), c_void_p]),
    "vsp_conv2d_gather_bf16": (c_int, [c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                       c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), c_int, c_int64, c_int64,
                                       c_int, c_int64, c_int64, c_int, c_int, c_int, c_int64, c_int64,
                                       POINTER(ConvEpilogue), c_void_p]),
    "vsp_conv_transpose2d_s2_bf16": (c_int, [c_void_p, c_void_p, c_void_p,
                                             c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                             c_int, c_int, c_int, c_int64, c_int64, POINTER(ConvEpilogue), c_void_p]),
    "vsp_blur_sep_nhwc_bf16": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float), c_void_p, c_int64, c_int64, c_int64,
                                       c_int64, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(ConvEpilogue), c_void_p]),
    "vsp_scale_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "vsp_se_tail_nhwc_bf16": (c_int, [c_void_p] * 7 + [c_int64] * 7 + [c_void_p]),
    "vsp_conv1x1_wgrad_small_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vsp_conv2d_branches_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                         c_int, POINTER(c_int), c_int, c_int64, c_int64, POINTER(ConvEpilogue), c_void_p]),
    "vsp_grouped_linear_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "vsp_conv2d_up2_fused_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                          c_int64, c_int64, c_int64, POINTER(ConvEpilogue), c_void_p]),
    "vsp_conv2d_up2h_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                     c_int64, c_int64, c_int64, POINTER(c_float), POINTER(ConvEpilogue), c_void_p]),
    "vsp_torgb_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int64, c_float, c_void_p]),
    "vsp_torgb_pool2_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_float), c_void_p,
                                          c_int64, c_int64, c_int64, c_int64, c_float, c_void_p]),
    "vsp_conv2d_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vsp_nchw_f32_to_nhwc_bf16_dot": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_void_p]),
    "vsp_nhwc_bf16_to_nchw_f32_dot": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_void_p]),
}


def load():
    """dlopen the library (building it first if needed) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.vsp_version() != 1:
        raise RuntimeError(f"vspbfr_b200: ABI version mismatch ({lib.vsp_version()} != 1)")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vsp_last_error()
        raise RuntimeError(f"vsp_b200 {what}: {msg.decode() if msg else 'unknown error'}")


def launch_count() -> int:
    return int(load().vsp_launch_count())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_no_guard = _NoGuard()


def device_guard(device):
    """``with device_guard(t.device):`` — torch.cuda.device(...) only when ``device`` is not already current (the guard costs
    ~4 us of host time per call, a quarter of the launch path of the small operators)."""
    import torch

    if device.index is None or device.index == torch.cuda.current_device():
        return _no_guard
    return torch.cuda.device(device)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
