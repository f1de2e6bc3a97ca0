"""The EVAL hot path's source (csrc/embed_kernel.cuh, head_kernels.cuh, api.cu) executed on the CPU by the
thread-per-CUDA-thread emulator of tests/emu (mbarriers and bulk copies included), against the golden vectors the
reference itself produced.  Kernel-logic coverage without a GPU; the `-m gpu` suite runs the same checks on the B200.
The emulator build is test infrastructure: sg_pr_b200 never loads it."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import sgpr_oracle as orc
from sg_pr_b200 import _lib, synth
from sg_pr_b200.engine import Engine

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def emu_engine(kitti_state):
    from tests.emu import build_emu
    lib = _lib.bind(C.CDLL(build_emu.build()), _lib.SYMBOLS)
    eng = Engine(lib=lib)
    eng.set_weights(kitti_state)
    yield eng
    eng.close()


def test_emulated_fused_kernel_matches_reference_scores(emu_engine):
    """Golden synthetic batch (N = 32, k = 10) produced by the unmodified reference: scores and attention within 1e-5."""
    g = np.load(os.path.join(GOLDEN, "ref_synth_n32_k10.npz"))
    f1, f2 = torch.from_numpy(g["features_1"])[:4], torch.from_numpy(g["features_2"])[:4]
    score, a1, a2 = emu_engine.forward_pairs(f1, f2, 10)
    assert float(np.abs(score.numpy() - g["score"][:4]).max()) <= 1e-5
    assert float(np.abs(a1.numpy().reshape(4, -1) - g["att_1"][:4].reshape(4, -1)).max()) <= 1e-5
    assert emu_engine.launch_count() >= 1


def test_emulated_knn_sets_and_stage_outputs_match_oracle(emu_engine, kitti_state):
    """Per-stage parity of the embed kernel (k-NN sets of all six layers, layer outputs, pooled vectors) at N = 64, k = 20,
    then the pair head through score_pairs and the score matrix."""
    graphs = synth.make_graphs(3, 64, 20, seed=17)
    got = emu_engine.embed(graphs, 20, want_att=True, want_emb=True, trace=True)
    want = orc.embed_graphs(graphs, 20, kitti_state, want_trace=True)
    for layer in range(6):
        code = orc.classify_knn_rows(want["knn_pd"][layer], want["knn_idx"][layer], got["knn"][:, layer].long(),
                                     want["layer_in"][layer])
        assert int((code > 0).sum()) == 0, f"layer {layer}: k-NN rows differ"
    want_pooled = want["pooled"].reshape(3, 32)
    assert float((got["pooled"] - want_pooled).abs().max()) <= 1e-5
    assert float((got["emb"] - want["emb"]).abs().max()) <= 1e-5
    idx = torch.tensor([[0, 1], [2, 0]], dtype=torch.int32)
    pairs = emu_engine.score_pairs(got["pooled"], idx)
    mat = emu_engine.score_matrix(got["pooled"], got["pooled"])
    assert float((mat[0, 1] - pairs[0]).abs()) <= 2e-6 and float((mat[2, 0] - pairs[1]).abs()) <= 2e-6
    ref = orc.score_matrix(want_pooled, want_pooled, kitti_state)
    assert float((mat - ref).abs().max()) <= 1e-5


def test_emulated_reference_fixture_pair(emu_engine):
    """data/0.json vs data/250.json at the shipped config (K = 10, node_num = 100: the NPL = 4 kernel) — BASELINE config 1,
    the pair eval_pair.py prints "Score: 1.3489922e-06" for — and one positive pair."""
    with np.load(os.path.join(GOLDEN, "ref_fixture_pairs.npz")) as z:
        for a, b in (("0", "250"), ("0", "3")):
            p = f"K10_N100_{a}_{b}_"
            f1, f2 = torch.from_numpy(z[p + "features_1"]), torch.from_numpy(z[p + "features_2"])
            score, a1, a2 = emu_engine.forward_pairs(f1, f2, 10)
            assert abs(float(score[0]) - float(z[p + "score"][0])) <= 1e-5, (a, b)
            assert float(np.abs(a2.numpy() - z[p + "att_2"]).max()) <= 1e-5


def test_emulated_persistent_launch_and_edge_cases(emu_engine, kitti_state):
    """More graphs than resident CTAs (the emulated device has 4 SMs): heaviest-first work queue (sgpr_order_kernel);
    k = N, k = 1 and an all-pad graph; batch independence."""
    f1, f2 = synth.make_pair_batch(9, 16, 10, seed=4)                 # 18 graphs > 8 resident CTAs
    score, _, _ = emu_engine.forward_pairs(f1, f2, 10)
    want = orc.forward_pairs(f1, f2, 10, kitti_state)["score"]
    assert float((score - want).abs().max()) <= 1e-5
    one, _, _ = emu_engine.forward_pairs(f1[3:4], f2[3:4], 10)
    assert torch.equal(one[0], score[3])                                # a pair's score does not depend on its batch
    g = synth.make_graphs(2, 16, 4, seed=2)
    g[1] = 0.0                                                          # a graph of zero pads only
    for k in (16, 1):
        got = emu_engine.embed(g, k)["pooled"]
        ref = orc.embed_graphs(g, k, kitti_state)["pooled"].reshape(2, 32)
        assert float((got - ref).abs().max()) <= 1e-5, k
    with pytest.raises(_lib.SgprError, match="topk"):
        emu_engine.embed(g, 17)


@pytest.mark.parametrize("tag,tol", [("n64_k20", 1e-5), ("n16_k10", 1e-5)])
def test_emulated_synthetic_goldens(emu_engine, tag, tol):
    """Reference-produced synthetic batches at the headline shape (N = 64, k = 20) and the smallest one."""
    g = np.load(os.path.join(GOLDEN, f"ref_synth_{tag}.npz"))
    f1, f2 = torch.from_numpy(g["features_1"]), torch.from_numpy(g["features_2"])
    score, a1, a2 = emu_engine.forward_pairs(f1, f2, int(g["K"]))
    assert float(np.abs(score.numpy() - g["score"]).max()) <= tol
    assert float(np.abs(a1.numpy().reshape(len(f1), -1) - g["att_1"].reshape(len(f1), -1)).max()) <= tol
    assert float(np.abs(a2.numpy().reshape(len(f1), -1) - g["att_2"].reshape(len(f1), -1)).max()) <= tol


def test_emulated_host_entry_point_matches_device_entry(emu_engine):
    """sgpr_forward_pairs_host (staged copies + kernel + copies back) == sgpr_forward_pairs, bit for bit."""
    f1, f2 = synth.make_pair_batch(3, 32, 10, seed=12)
    d_score, d_a1, d_a2 = emu_engine.forward_pairs(f1, f2, 10)
    h_score, h_a1, h_a2 = emu_engine.forward_pairs_host(f1, f2, 10)
    assert torch.equal(h_score, d_score) and torch.equal(h_a1, d_a1) and torch.equal(h_a2, d_a2)
    empty, _, _ = emu_engine.forward_pairs(f1[:0], f2[:0], 10)
    assert empty.shape == (0,)


def test_emulated_tensor_core_variant_of_the_fused_kernel(kitti_state, monkeypatch):
    """csrc/embed_tc_kernel.cuh (SGPR_EMBED_TC=1): Gram / GEMM of the 64-channel layers as 3xTF32 UMMAs — here with the
    emulator's model of tcgen05 / TMEM (tests/emu + csrc/tc_ops.cuh), which checks the operand-plane swizzle, the TMEM
    column map and the lane / row assignments.  Same scores as the FFMA kernel and the oracle."""
    from tests.emu import build_emu
    lib = _lib.bind(C.CDLL(build_emu.build()), _lib.SYMBOLS)
    monkeypatch.setenv("SGPR_EMBED_TC", "1")
    tc = Engine(lib=lib)
    monkeypatch.delenv("SGPR_EMBED_TC")
    ff = Engine(lib=lib)
    tc.set_weights(kitti_state)
    ff.set_weights(kitti_state)
    for n, k in ((40, 10), (64, 20)):
        f1, f2 = synth.make_pair_batch(2, n, k, seed=5)
        a, b = tc.forward_pairs(f1, f2, k), ff.forward_pairs(f1, f2, k)
        want = orc.forward_pairs(f1, f2, k, kitti_state)
        assert float((a[0] - b[0]).abs().max()) <= 5e-6 and float((a[0] - want["score"]).abs().max()) <= 1e-5
        assert float((a[1] - want["att_1"]).abs().max()) <= 1e-5
    e = tc.embed(f1, 20, want_emb=True)
    assert float((e["emb"] - want["emb_1"]).abs().max()) <= 2e-5
    tc.close()
    ff.close()


def test_emulated_branch_split_launch_is_bit_identical(kitti_state, monkeypatch):
    """EmbedArgs::split (small launches: one EdgeConv branch of a graph per work unit, the second arriver merges) against
    whole-graph units (SGPR_NO_SPLIT=1): every output bit for bit, pairs / embed / compact records, both tie rules.  The
    emulated device has 8 resident CTA slots, so G <= 5 graphs launch split and larger batches do not."""
    from sg_pr_b200.engine import compact_graphs
    from tests.emu import build_emu
    lib = _lib.bind(C.CDLL(build_emu.build()), _lib.SYMBOLS)
    split = Engine(lib=lib)
    monkeypatch.setenv("SGPR_NO_SPLIT", "1")
    whole = Engine(lib=lib)
    monkeypatch.delenv("SGPR_NO_SPLIT")
    for e in (split, whole):
        e.set_weights(kitti_state)
    for b, n, k in ((1, 40, 10), (2, 64, 20), (2, 100, 10)):
        f1, f2 = synth.make_pair_batch(b, n, k, seed=20 + b)
        got, ref = split.forward_pairs(f1, f2, k), whole.forward_pairs(f1, f2, k)
        assert all(torch.equal(x, y) for x, y in zip(got, ref))
        want = orc.forward_pairs(f1, f2, k, kitti_state)
        assert float((got[0] - want["score"]).abs().max()) <= 1e-5
        again = split.forward_pairs_compact(compact_graphs(f1), compact_graphs(f2), n, k)
        assert torch.equal(again[0], ref[0])
    g = synth.make_graphs(5, 64, 20, seed=3)
    a, b = split.embed(g, 20, want_att=True, want_emb=True), whole.embed(g, 20, want_att=True, want_emb=True)
    assert all(torch.equal(a[key], b[key]) for key in ("pooled", "att", "emb"))
    for e in (split, whole):
        e.set_knn_ties("cpu")
    f1, f2 = synth.make_pair_batch(2, 32, 10, seed=9, dense=True)
    assert torch.equal(split.forward_pairs(f1, f2, 10)[0], whole.forward_pairs(f1, f2, 10)[0])
    split.close()
    whole.close()


def test_emulated_random_shapes_vs_oracle(kitti_state):
    """A seeded random walk over (node_num, k, batch, density): odd node counts on both sides of the 32 / 64 row-tiling
    boundaries, k from 1 up, dense batches under the CPU tie rule — every score and attention value within 1e-5 of the
    oracle, compact records bit-identical to the one-hot blocks.  Small batches launch branch-split, larger ones do not."""
    import random
    from sg_pr_b200.engine import compact_graphs
    from tests.emu import build_emu
    eng = Engine(lib=_lib.bind(C.CDLL(build_emu.build()), _lib.SYMBOLS))
    eng.set_weights(kitti_state)
    rnd = random.Random(11)
    for it in range(12):
        n = rnd.choice([2, 3, 5, 17, 31, 33, 47, 63, 65, 96])
        k = rnd.randint(1, max(1, min(n - 1, 24)))
        b = rnd.choice([1, 2, 3, 5])
        dense = rnd.random() < 0.25
        try:
            f1, f2 = synth.make_pair_batch(b, n, k, seed=it, dense=dense)
        except ValueError:                                   # no room for k pads: only a dense graph has this (n, k)
            f1, f2 = synth.make_pair_batch(b, n, k, seed=it, dense=True)
            dense = True
        eng.set_knn_ties("cpu" if dense else "cuda")         # the oracle is the reference on the CPU
        got = eng.forward_pairs(f1, f2, k)
        want = orc.forward_pairs(f1, f2, k, kitti_state)
        assert float((got[0] - want["score"]).abs().max()) <= 1e-5, (n, k, b, dense)
        assert float((got[1] - want["att_1"]).abs().max()) <= 1e-5 and float((got[2] - want["att_2"]).abs().max()) <= 1e-5
        assert torch.equal(eng.forward_pairs_compact(compact_graphs(f1), compact_graphs(f2), n, k)[0], got[0])
    eng.close()
