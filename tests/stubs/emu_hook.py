"""TEST-ONLY HOOK: run a script against the CPU emulator build of the kernels (tests/emu) instead of a GPU.

    python -c "import tests.stubs.emu_hook as h; h.run('/root/reference/eval_pair.py')"

Used by tests/test_reference_scripts.py to execute the reference's UNMODIFIED eval_pair.py / eval_batch.py in the build
container (no GPU there).  It injects the emulator library into sg_pr_b200.engine.Engine and points SG at the CPU; nothing
in the product imports this module, and the product library never loads the emulator by itself."""
import ctypes
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def install():
    import torch
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from sg_pr_b200 import _lib, engine, sg_net
    from tests.emu import build_emu
    lib = _lib.bind(ctypes.CDLL(build_emu.build()), _lib.SYMBOLS)
    original = engine.Engine.__init__

    def emulated_init(self, device=0, lib=lib):
        original(self, device, lib=lib)

    engine.Engine.__init__ = emulated_init
    sg_net.SG._device = lambda self: torch.device("cpu")


def run(script: str):
    install()
    sys.argv = [script] + sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
