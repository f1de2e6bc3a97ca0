"""In-tree build of the sm_100a C-ABI library (nvcc cross-compiles without a GPU).

    python -m sg_pr_b200.build [--force]

Output: sg_pr_b200/libsgpr_b200.so (git-ignored; travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsgpr_b200.so")
SOURCES = ["api.cu", "train.cu"]
DEPS = ["api.cu", "train.cu", "train_kernels.cuh", "common.cuh", "embed_kernel.cuh", "head_kernels.cuh", "pack.hpp", "sortnet32.inc", "sortnet16.inc", "sortnet8.inc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(ROOT, "include", h) for h in ("sgpr_b200.h", "sgpr_b200_train.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
