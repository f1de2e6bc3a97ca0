"""The training kernels' SOURCE (sg_pr_b200/csrc/train_kernels.cuh + train.cu) executed on the CPU by the
thread-per-CUDA-thread emulator of tests/emu, against the reference's golden training vectors.

This checks kernel LOGIC (indexing, barriers, the backward algebra, the optimiser) without a GPU; the `-m gpu` twin
(tests/test_gpu_train.py) runs the identical checks through the real library.  The emulator build is test
infrastructure: sg_pr_b200 never loads it."""
import ctypes as C

import pytest

from sg_pr_b200 import _lib
from sg_pr_b200.train_engine import TrainEngine, layout
from tests import train_checks as tc


@pytest.fixture(scope="module")
def emu_lib():
    from tests.emu import build_emu
    return _lib.bind(C.CDLL(build_emu.build()), dict(_lib.TRAIN_SYMBOLS, sgpr_last_error=(C.c_char_p, [])))


def test_layout_covers_the_state_dict(emu_lib, kitti_state):
    entries, n_params = layout(emu_lib)
    names = [n for n, _, _ in entries]
    floats = {k for k, v in kitti_state.items() if v.dtype.is_floating_point}
    assert set(names) == floats                                   # everything but the int64 num_batches_tracked
    end = 0
    for name, off, size in entries:                               # dense, in order, sizes as in the checkpoint
        assert off == end and size == kitti_state[name].numel(), name
        end = off + size
    assert end == emu_lib.sgpr_train_state_count()
    assert sum(s for _, _, s in entries[:n_params]) == emu_lib.sgpr_train_param_count() == 47985
    assert all("running_" not in n for n in names[:n_params]) and all("running_" in n for n in names[n_params:])


@pytest.mark.parametrize("mirrored", [False, True])
def test_emulated_gradients_match_reference(emu_lib, mirrored):
    g, sd = tc.load_case("n32_k10")
    eng = TrainEngine(lib=emu_lib)
    tc.check_gradients(eng, g, sd, "cpu", pred_tol=1e-5, mirrored=mirrored)
    eng.close()


def test_emulated_two_optimiser_steps_match_reference(emu_lib):
    g, sd = tc.load_case("n32_k10")
    eng = TrainEngine(lib=emu_lib)
    tc.check_two_steps(eng, g, sd, "cpu", pred_tol=1e-5, mirrored=True)
    eng.close()


def test_emulated_split_forward_backward(emu_lib):
    g, sd = tc.load_case("n32_k10")
    eng = TrainEngine(lib=emu_lib)
    tc.check_split_forward_backward(eng, g, sd, "cpu", pred_tol=1e-5)
    import torch
    f1, t = torch.from_numpy(g["features_1"]), torch.from_numpy(g["target"])
    eng.step(f1, None, t, int(g["K"]), apply=False, mirrored=True)       # a fused step reuses the workspace ...
    with pytest.raises(_lib.SgprError, match="no sgpr_train_forward"):
        eng.backward(torch.zeros(8))                                      # ... so the old forward can no longer be differentiated
    eng.close()


def test_emulated_errors_are_loud(emu_lib):
    import torch
    g, sd = tc.load_case("n32_k10")
    eng = TrainEngine(lib=emu_lib)
    f1 = torch.from_numpy(g["features_1"])
    t = torch.from_numpy(g["target"])
    with pytest.raises(_lib.SgprError, match="set_state"):
        eng.step(f1, f1, t, 10)
    eng.set_state(sd)
    with pytest.raises(_lib.SgprError, match="topk"):
        eng.step(f1, f1, t, 33)
    with pytest.raises(ValueError):
        eng.step(f1, f1[:, :, :16].contiguous(), t, 10)
    with pytest.raises(_lib.SgprError, match="Adam"):
        eng.set_optimizer(-1.0)
    eng.close()


def test_emulated_batch_assembly_matches_host_augmentation(emu_lib, monkeypatch):
    from tests import assemble_checks as ac
    eng = TrainEngine(lib=emu_lib)
    ac.check_assemble(eng, "cpu", monkeypatch, M=8, N=32, P=5)
    eng.close()


def test_emulated_gradients_match_reference_at_64_nodes(emu_lib):
    """The headline shape (N = 64, k = 20: two 32-column blocks per distance row, the NPL = 2 kernels), mirrored step."""
    g, sd = tc.load_case("n64_k20")
    eng = TrainEngine(lib=emu_lib)
    tc.check_gradients(eng, g, sd, "cpu", pred_tol=5e-5, mirrored=True)
    eng.close()


def test_emulated_general_two_sided_batch(emu_lib, kitti_state):
    eng = TrainEngine(lib=emu_lib)
    tc.check_general_two_sided_batch(eng, kitti_state, "cpu", B=3, N=32, k=10)
    eng.close()
