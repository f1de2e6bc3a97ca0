"""TEST INFRASTRUCTURE: build the kernels' source with the HOST compiler against tests/emu/cuda_emu.h.

    python tests/emu/build_emu.py        ->  tests/emu/libsgpr_emu.so   (git-ignored)

The result exports the whole C-ABI (include/sgpr_b200.h, include/sgpr_b200_train.h), executed by a thread-per-CUDA-thread
emulator.  It exists to debug kernel logic on a machine without a GPU (tests/test_train_emu.py, tests/test_eval_emu.py);
the product never loads it.  Same translation units as sg_pr_b200/build.py, minus the tcgen05 score-matrix kernel
(inline PTX; the emulated library keeps the fp32-FMA score-matrix kernel only).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sg_pr_b200", "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsgpr_emu.so")
DEPS = [os.path.join(CSRC, f) for f in ("api.cu", "train.cu", "embed_inst.cu", "train_inst.cu", "train_kernels.cuh",
                                        "embed_kernel.cuh", "head_kernels.cuh", "common.cuh", "pack.hpp", "topk_nth.cuh", "embed_tc_kernel.cuh", "tc_ops.cuh",
                                        "launchers.hpp")] + \
       [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "sgpr_b200_train.h"),
        os.path.join(ROOT, "include", "sgpr_b200.h"), os.path.abspath(__file__)]
UNITS = {"api.o": ("api.cu", []), "train.o": ("train.cu", [])}
for _npl in (1, 2, 4):
    UNITS[f"embed_npl{_npl}.o"] = ("embed_inst.cu", [f"-DSGPR_INST_NPL={_npl}"])
    UNITS[f"train_npl{_npl}.o"] = ("train_inst.cu", [f"-DSGPR_INST_NPL={_npl}"])
FLAGS = ["-std=c++20", "-O1", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-DSGPR_EMU"]


def _compile(item):
    obj, (src, defs) = item
    cmd = ["g++", *FLAGS, *defs, "-x", "c++", "-c", os.path.join(CSRC, src), "-I", os.path.join(ROOT, "include"),
           "-o", os.path.join(OBJ_DIR, obj)]
    return obj, subprocess.run(cmd, capture_output=True, text=True)


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        for obj, res in pool.map(_compile, UNITS.items()):
            if res.returncode != 0:
                raise RuntimeError(f"emulator build failed for {obj}:\n" + res.stdout + res.stderr)
    res = subprocess.run(["g++", "-shared", "-pthread", "-o", LIB, *[os.path.join(OBJ_DIR, o) for o in sorted(UNITS)], "-latomic"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
