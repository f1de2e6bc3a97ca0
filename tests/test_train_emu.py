"""The training kernels' SOURCE (sg_pr_b200/csrc/train_kernels.cuh + train.cu) executed on the CPU by the
thread-per-CUDA-thread emulator of tests/emu, against the reference's golden training vectors.

This checks kernel LOGIC (indexing, barriers, the backward algebra, the optimiser) without a GPU; the `-m gpu` twin
(tests/test_gpu_train.py) runs the identical checks through the real library.  The emulator build is test
infrastructure: sg_pr_b200 never loads it."""
import ctypes as C
import os

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


def test_emulated_process_batch_glue(emu_lib, golden_dir, tmp_path):
    """SGTrainer.process_batch(batch, training=True) — host path, mirrored step, state sync — with the emulated library
    standing in for the device: the step equals the oracle's on the batch the reference-shaped host code builds from the
    same seeds, and sync_model_from_device() brings parameters, running statistics and counters back into the module."""
    import random
    import numpy as np
    import torch
    from oracle import sgpr_oracle_train as ort
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SGTrainer, _DeviceAdam
    from sg_pr_b200.utils import process_pair
    from tests.helpers import write_fixture_tree
    root = str(tmp_path)
    cfg = write_fixture_tree(golden_dir, root)
    os.makedirs(f"{root}/lists", exist_ok=True)
    for seq in ("00", "08"):
        with open(f"{root}/lists/{seq}.txt", "w") as f:
            f.write("0.json 3.json\n0.json 250.json\n")
    args = sgpr_args().load(cfg)
    args.K, args.node_num, args.batch_size = 10, 64, 2
    trainer = SGTrainer(args, True)
    trainer.optimizer = _DeviceAdam(trainer)
    state0 = {k: v.detach().clone() for k, v in trainer.model.module.state_dict().items()}
    eng = TrainEngine(lib=emu_lib)                       # what _device_trainer() would create on a GPU box
    eng.set_state(state0, reset_optimizer=True)
    eng.set_optimizer(float(args.learning_rate), float(args.weight_decay))
    trainer._train_engine, trainer._unsynced_steps = eng, 0
    random.seed(5); np.random.seed(5)
    loss, pred, gt = trainer.process_batch(trainer.training_graphs, True)
    random.seed(5); np.random.seed(5)
    f1, tg = [], []
    for pair in trainer.training_graphs:                 # the reference-shaped host path on the same seeds
        d = trainer.transfer_to_torch(process_pair(pair), True)
        f1 += [d["features_1"], d["features_2"]]
        tg += [d["target"]] * 2
    f2 = [f1[i ^ 1] for i in range(len(f1))]
    work = {k: v.clone() for k, v in state0.items()}
    want = ort.train_step(work, torch.FloatTensor(np.array(f1)), torch.FloatTensor(np.array(f2)), torch.FloatTensor(tg), 10,
                          ort.new_adam_state(work), float(args.learning_rate), float(args.weight_decay))
    assert gt.tolist() == tg and np.abs(pred - want["pred"].numpy()).max() < 5e-4
    assert abs(loss - want["loss"]) < 1e-3 * max(1.0, want["loss"])
    trainer.sync_model_from_device()
    synced = trainer.model.module.state_dict()
    assert int(synced["dgcnn_conv_end.1.num_batches_tracked"]) == int(state0["dgcnn_conv_end.1.num_batches_tracked"]) + 2
    for name in ("dgcnn_s_conv1.1.running_mean", "dgcnn_conv_end.1.running_var"):
        ref = work[name]
        assert float((synced[name] - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), name
    assert not torch.equal(synced["attention.weight_matrix"], state0["attention.weight_matrix"])

    # fit()'s pipelining: naming the batch that follows makes process_batch build it while the current step runs on the
    # device — the same draws in the same order, hence the same batches, losses and stream positions as without it
    batches = [trainer.training_graphs, list(reversed(trainer.training_graphs)), trainer.training_graphs[:1]]

    def epoch(prefetch):
        eng.set_state(state0, reset_optimizer=True)
        random.seed(9); np.random.seed(9)
        out = []
        for i, b in enumerate(batches):
            if prefetch and i + 1 < len(batches):
                trainer._upcoming_batch = batches[i + 1]
            out.append(trainer.process_batch(b, True))
        assert getattr(trainer, "_prefetched", None) is None and getattr(trainer, "_upcoming_batch", None) is None
        return out, (np.random.rand(), random.random())

    plain, tail_plain = epoch(False)
    piped, tail_piped = epoch(True)
    assert tail_plain == tail_piped
    for x, y in zip(plain, piped):
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2])
    eng.close()


def test_emulated_train_forward_in_cpu_tie_mode_matches_oracle_on_dense_graphs(emu_lib, kitti_state):
    """Graphs without zero pads (the k-NN tie-rule regime, see tests/test_tie_rule.py): the train-mode forward with
    knn_ties="cpu" reproduces the reference's CPU predictions; the default rule does not (it is the CUDA reference)."""
    import torch
    from oracle import sgpr_oracle_train as ort
    from sg_pr_b200 import synth
    a, b = synth.make_pair_batch(3, 32, 10, seed=8, dense=True)
    f1 = torch.stack([a, b], dim=1).reshape(6, 15, 32).contiguous()
    f2 = torch.stack([b, a], dim=1).reshape(6, 15, 32).contiguous()
    state = {name: value.clone() for name, value in kitti_state.items()}    # forward_train updates BN buffers in place
    want = ort.forward_train(state, f1, f2, 10)[0].detach()
    eng = TrainEngine(lib=emu_lib)
    eng.set_state(kitti_state)
    eng.set_knn_ties("cpu")
    pred, _, _ = eng.forward(f1, f2, 10, update_running=False)
    assert float((pred - want).abs().max()) <= 2e-5
    eng.set_knn_ties("cuda")
    plain, _, _ = eng.forward(f1, f2, 10, update_running=False)
    assert float((plain - want).abs().max()) > 1e-4
    with pytest.raises(ValueError):
        eng.set_knn_ties("numpy")
    eng.close()


def test_backward_of_a_stale_forward_is_refused(emu_lib, kitti_state):
    """the engine keeps ONE forward's activations.  A second train-mode forward before backward() must make
    the first node's backward raise instead of silently differentiating the second forward."""
    import torch
    from sg_pr_b200 import synth
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SG
    args = sgpr_args()
    args.K, args.node_num = 10, 32
    model = SG(args, 12)
    model.load_state_dict(kitti_state)
    model._device = lambda: torch.device("cpu")
    model._autograd_engine = TrainEngine(lib=emu_lib)
    model.train()
    fa = synth.make_pair_batch(2, 32, 10, seed=1)
    fb = synth.make_pair_batch(2, 32, 10, seed=2)
    pa, _, _ = model({"features_1": fa[0], "features_2": fa[1]})
    pb, _, _ = model({"features_1": fb[0], "features_2": fb[1]})
    gen = model._autograd_engine.forward_generation()
    assert gen == 2
    with pytest.raises(RuntimeError, match="another train-mode forward"):
        pa.sum().backward()
    pb.sum().backward()                                  # the latest forward is still differentiable
    assert model.scoring_layer.weight.grad is not None
