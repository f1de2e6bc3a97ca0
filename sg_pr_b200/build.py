"""In-tree build of the sm_100a C-ABI library (nvcc cross-compiles without a GPU).

    python -m sg_pr_b200.build [--force] [-v]

Output: sg_pr_b200/libsgpr_b200.so (git-ignored; travels to the GPU box with the gpurun snapshot).  Every translation
unit is compiled to its own object under sg_pr_b200/build/ (in parallel, only when one of its dependencies is newer)
and the objects are linked with a static cudart: the library depends on neither torch nor a CUDA toolkit at run time.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ_DIR = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libsgpr_b200.so")
HEADERS = [os.path.join(ROOT, "include", h) for h in ("sgpr_b200.h", "sgpr_b200_train.h")]
COMMON = ["common.cuh", "launchers.hpp"]
EMBED_DEPS = COMMON + ["embed_kernel.cuh", "embed_tc_kernel.cuh", "tc_ops.cuh", "topk_nth.cuh", "sortnet32.inc", "sortnet16.inc", "sortnet8.inc"]
TRAIN_DEPS = EMBED_DEPS + ["train_kernels.cuh"]
# object name -> (source, extra defines, dependencies)
UNITS = {
    "api.o": ("api.cu", [], EMBED_DEPS + ["head_kernels.cuh", "pack.hpp"]),
    "scoremat_umma.o": ("scoremat_umma.cu", [], COMMON + ["scoremat_umma.cuh"]),
    "train.o": ("train.cu", [], TRAIN_DEPS),
}
for _npl in (1, 2, 4):
    UNITS[f"embed_npl{_npl}.o"] = ("embed_inst.cu", [f"-DSGPR_INST_NPL={_npl}"], EMBED_DEPS)
    UNITS[f"train_npl{_npl}.o"] = ("train_inst.cu", [f"-DSGPR_INST_NPL={_npl}"], TRAIN_DEPS)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550"]
# experiments: SGPR_EXTRA_NVCC_FLAGS="-DSGPR_UMMA_EPI_WARPS=16" SGPR_BUILD_TAG=epi16 builds libsgpr_b200_epi16.so beside
# the product library (select it at run time with SGPR_B200_LIB=...); the product build sets neither
EXTRA = os.environ.get("SGPR_EXTRA_NVCC_FLAGS", "").split()
TAG = os.environ.get("SGPR_BUILD_TAG", "")
if TAG:
    OBJ_DIR = os.path.join(PKG, "build", "variant_" + TAG)
    LIB = os.path.join(ROOT, "tools", "variants", f"libsgpr_b200_{TAG}.so")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _units():
    return {o: u for o, u in UNITS.items() if os.path.exists(os.path.join(CSRC, u[0]))}


def _unit_stale(obj: str, unit) -> bool:
    path = os.path.join(OBJ_DIR, obj)
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, unit[0])] + [os.path.join(CSRC, d) for d in unit[2]] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_unit_stale(o, u) or os.path.getmtime(os.path.join(OBJ_DIR, o)) > t for o, u in _units().items())


def _compile(obj: str, unit, verbose: bool):
    cmd = [_nvcc(), *NVCC_FLAGS, *EXTRA, *unit[1], "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, unit[0]),
           "-o", os.path.join(OBJ_DIR, obj)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return obj, cmd, res


def build(force: bool = False, verbose: bool = False) -> str:
    units = _units()
    todo = {o: u for o, u in units.items() if force or _unit_stale(o, u)}
    if not todo and os.path.exists(LIB) and not stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(todo) or 1, os.cpu_count() or 4)) as pool:
        for obj, cmd, res in pool.map(lambda it: _compile(it[0], it[1], verbose), todo.items()):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {obj} ({res.returncode}):\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(f"==== {obj}\n{res.stderr}")
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB,
            *[os.path.join(OBJ_DIR, o) for o in sorted(units)]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
