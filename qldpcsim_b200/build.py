"""In-tree build of libqldpc_b200.so (hand-written CUDA for sm_100a, C ABI in include/qldpc_b200.h).

nvcc cross-compiles without a GPU.  -fmad=false: the min-sum and BP kernels are specified as individually
rounded IEEE operations (see csrc/ms_kernel.cuh); -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libqldpc_b200.so")
SRC = os.path.join(HERE, "csrc", "qldpc_api.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    out = [SRC, os.path.join(os.path.dirname(HERE), "include", "qldpc_b200.h")]
    d = os.path.join(HERE, "csrc")
    out += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cuh", ".cu", ".h", ".cpp"))]
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
