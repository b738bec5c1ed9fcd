#!/usr/bin/env python
"""Untracked scratch copy of the reference for GPU-side incumbents and the drop-in proof.

    python tools/make_incumbent.py        # in the build container (needs /root/reference)

Copies the reference tree (without images / git data) to ``baseline/_ref/VSPBFR`` — git-ignored, NOT gpurun-ignored, so it
travels to the GPU box — and JIT-builds ITS OWN ``op/`` CUDA extensions (op/upfirdn2d_kernel.cu, op/fused_bias_act_kernel.cu,
unmodified) for sm_100a into ``op/cache_*`` there, so that ``tools/bench_incumbent.py`` can time the reference's kernels and
models on the B200 and ``tests/test_dropin_gpu.py`` can run the reference's model files on top of this package's ``op``.
Nothing under ``vspbfr_b200/`` reads this directory; no reference source enters the git history."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref", "VSPBFR")


def main(force=False):
    if not os.path.isdir(REF):
        print("make_incumbent: /root/reference not present (GPU box) — nothing to do")
        return 0
    if os.path.isdir(DST) and not force and os.path.exists(os.path.join(DST, "op", "cache_upfirdn2d", "upfirdn2d.so")):
        print("make_incumbent: baseline/_ref/VSPBFR already built")
        return 0
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(REF, DST, ignore=shutil.ignore_patterns(".git", "imgs", "__pycache__", "cache_*", "*.pyc"))
    for d, _, files in os.walk(DST):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="4")
    code = "import sys; sys.path.insert(0, %r); import op; print('built', op.__file__)" % DST
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=DST, capture_output=True, text=True)
    print(r.stdout[-2000:], r.stderr[-3000:])
    return r.returncode


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
