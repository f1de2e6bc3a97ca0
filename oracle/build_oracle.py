"""TEST INFRASTRUCTURE: compile the C++ piece of the oracle (oracle/topk_ref.cpp -> oracle/_build/libtopk_ref.so).

    python oracle/build_oracle.py [--force]

The reference itself is Python, so there is nothing to compile into oracle/_ref; this library only wraps the real
libstdc++ std::nth_element behind ATen's CPU topk comparator (the un-vendored dependency that decides k-NN ties,
/root/reference/dgcnn.py:19)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "topk_ref.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtopk_ref.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", LIB], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
