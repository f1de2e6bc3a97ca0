"""TEST INFRASTRUCTURE: build the training kernels' source with the HOST compiler against tests/emu/cuda_emu.h.

    python tests/emu/build_emu.py        ->  tests/emu/libsgpr_emu.so   (git-ignored)

The result exports the whole C-ABI (include/sgpr_b200.h, include/sgpr_b200_train.h), executed by a thread-per-CUDA-thread emulator.
It exists to debug kernel logic on a machine without a GPU (tests/test_train_emu.py); the product never loads it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sg_pr_b200", "csrc")
LIB = os.path.join(HERE, "libsgpr_emu.so")
DEPS = [os.path.join(CSRC, f) for f in ("api.cu", "train.cu", "train_kernels.cuh", "embed_kernel.cuh", "head_kernels.cuh",
                                        "common.cuh", "pack.hpp")] + \
       [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "sgpr_b200_train.h"),
        os.path.join(ROOT, "include", "sgpr_b200.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DSGPR_EMU", "-x", "c++",
           os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "train.cu"), "-I", os.path.join(ROOT, "include"), "-o", LIB,
           "-latomic"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
