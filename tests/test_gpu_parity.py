"""GPU parity tests (run with `-m gpu` on the B200 box).  Every call goes through the C-ABI (sg_pr_b200.engine →
libsgpr_b200.so); the checker is the oracle (oracle/sgpr_oracle.py, run live on the host CPU) and the committed
golden vectors the reference itself produced (tests/golden/).  Tolerance: 1e-5 abs fp32 on scores (BASELINE.json),
k-NN index sets exact up to exact ties."""
import os

import numpy as np
import pytest
import torch

from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5
FEAT_ATOL, FEAT_RTOL = 2e-5, 1e-5


@pytest.fixture(scope="module")
def eng(kitti_state):
    from sg_pr_b200.engine import Engine
    e = Engine(0)
    e.set_weights(kitti_state)
    yield e
    e.close()


def _cuda(t):
    return t.cuda(non_blocking=False)


def test_library_is_loaded_native():
    """The CUDA path is the one that runs: the in-tree .so must be mapped into this process."""
    from sg_pr_b200 import _lib
    _lib.load()
    with open("/proc/self/maps") as f:
        assert "libsgpr_b200.so" in f.read()


FIXTURE_PAIRS = [("0", "250"), ("0", "3"), ("3", "0"), ("0", "0"), ("250", "0"), ("3", "250")]


@pytest.mark.parametrize("K,N", [(10, 100), (20, 64)])
def test_fixture_pairs_match_reference_golden(eng, golden_dir, K, N):
    with np.load(os.path.join(golden_dir, "ref_fixture_pairs.npz")) as z:
        for a, b in FIXTURE_PAIRS:
            p = f"K{K}_N{N}_{a}_{b}_"
            f1, f2 = torch.from_numpy(z[p + "features_1"]), torch.from_numpy(z[p + "features_2"])
            score, a1, a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), K)
            assert abs(float(score[0]) - float(z[p + "score"][0])) <= SCORE_TOL, (K, N, a, b)
            assert np.abs(a1.cpu().numpy() - z[p + "att_1"]).max() <= SCORE_TOL
            assert np.abs(a2.cpu().numpy() - z[p + "att_2"]).max() <= SCORE_TOL


@pytest.mark.parametrize("tag", ["n64_k20", "n100_k10", "n32_k10", "n128_k20", "n16_k10"])
def test_synthetic_golden(eng, golden_dir, tag):
    with np.load(os.path.join(golden_dir, f"ref_synth_{tag}.npz")) as z:
        K = int(z["K"])
        f1, f2 = torch.from_numpy(z["features_1"]), torch.from_numpy(z["features_2"])
        score, a1, a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), K)
        assert np.abs(score.cpu().numpy() - z["score"]).max() <= SCORE_TOL
        assert np.abs(a1.cpu().numpy() - z["att_1"]).max() <= SCORE_TOL
        assert np.abs(a2.cpu().numpy() - z["att_2"]).max() <= SCORE_TOL
        emb = eng.embed(_cuda(f1), K, want_emb=True)["emb"].cpu().numpy()
        np.testing.assert_allclose(emb, z["emb_1"], atol=FEAT_ATOL, rtol=FEAT_RTOL)


def test_dense_golden_in_cpu_tie_mode(eng, golden_dir):
    """Graphs WITHOUT zero pads: k-NN ties on the one-hot branch fall between different nodes, so the tie rule of topk
    (dgcnn.py:19) decides the score.  The golden batch was produced by the reference on a CPU; `knn_ties="cpu"` restates
    ATen's CPU rule (csrc/topk_nth.cuh) and must reproduce it to 1e-5 with identical k-NN sets."""
    with np.load(os.path.join(golden_dir, "ref_synth_n64_k20_dense.npz")) as z:
        K = int(z["K"])
        f1, f2 = torch.from_numpy(z["features_1"]), torch.from_numpy(z["features_2"])
        eng.set_knn_ties("cpu")
        try:
            score, a1, a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), K)
            got = eng.embed(_cuda(f1), K, want_emb=True, trace=True)
        finally:
            eng.set_knn_ties("cuda")
        assert np.abs(score.cpu().numpy() - z["score"]).max() <= SCORE_TOL
        assert np.abs(a1.cpu().numpy() - z["att_1"]).max() <= SCORE_TOL
        assert np.abs(a2.cpu().numpy() - z["att_2"]).max() <= SCORE_TOL
        np.testing.assert_allclose(got["emb"].cpu().numpy(), z["emb_1"], atol=FEAT_ATOL, rtol=FEAT_RTOL)
        knn = got["knn"].cpu().numpy().astype(np.int64)
        for layer in range(6):
            assert np.array_equal(np.sort(knn[:, layer], axis=-1), np.sort(z[f"knn_idx_1_{layer}"], axis=-1)), layer
        # and the default rule is measurably NOT the CPU reference here (that is what the mode is for)
        plain, _, _ = eng.forward_pairs(_cuda(f1), _cuda(f2), K)
        assert np.abs(plain.cpu().numpy() - z["score"]).max() > 1e-3


@pytest.mark.parametrize("n,k", [(64, 20), (32, 10), (100, 10), (128, 20), (16, 10)])
def test_cpu_tie_mode_dense_vs_oracle(eng, kitti_state, n, k):
    """`knn_ties="cpu"` on dense batches of every tile size vs the live CPU oracle: 1e-5, near ties explained one by one."""
    from tests.helpers import assert_scores_match_or_near_tie
    f1, f2 = synth.make_pair_batch(24, n, k, seed=300 + n, dense=True)
    want = orc.forward_pairs(f1, f2, k, kitti_state)
    eng.set_knn_ties("cpu")
    try:
        score, _, _ = eng.forward_pairs(_cuda(f1), _cuda(f2), k)
        assert_scores_match_or_near_tie(eng, kitti_state, f1, f2, k, score, want["score"], SCORE_TOL)
    finally:
        eng.set_knn_ties("cuda")


@pytest.mark.parametrize("n,k", [(64, 20), (32, 10), (100, 10)])
def test_default_tie_rule_is_the_reference_on_cuda(eng, kitti_state, n, k):
    """The default rule (ties to the lowest index) against the reference's module code run on THIS GPU (ATen CUDA topk,
    TF32 off) on dense graphs: every k-NN row selects the same nodes (class 0) or sits on a rounding-level near tie, the
    pooled vectors agree to 1e-5 — while the same reference run on the CPU is ~0.1 away (tools/tie_rule_probe.py)."""
    from tests.helpers import reference_on_cuda, tie_divergences
    g = synth.make_graphs(32, n, k, seed=500 + n, dense=True)
    ref = reference_on_cuda(kitti_state, g, k)
    first = tie_divergences(eng, kitti_state, g, k, want=ref)
    assert all(c == 2 for _, codes in first.values() for c in codes), f"non-near-tie divergence from the CUDA reference: {first}"
    assert len(first) <= 2
    diverged = sorted({b for b, _ in first})
    keep = torch.ones(32, dtype=torch.bool)
    keep[diverged] = False
    pooled = eng.embed(_cuda(g), k)["pooled"].cpu()
    want = ref["pooled"].squeeze(-1)
    np.testing.assert_allclose(pooled[keep].numpy(), want[keep].numpy(), atol=5e-5, rtol=1e-5)
    cpu = orc.embed_graphs(g, k, kitti_state)["pooled"].squeeze(-1)
    assert float((cpu - want).abs().max()) > 1e-3          # the reference disagrees with itself across devices here


@pytest.mark.parametrize("tag", ["3_20_08", "10_20_05"])
def test_other_checkpoints(golden_dir, tag):
    from sg_pr_b200.engine import Engine
    e = Engine(0)
    e.set_weights(orc.load_state_npz(os.path.join(golden_dir, f"model_{tag}.npz")))
    with np.load(os.path.join(golden_dir, "ref_ckpt_scores.npz")) as z:
        score, a1, _ = e.forward_pairs(_cuda(torch.from_numpy(z["features_1"])), _cuda(torch.from_numpy(z["features_2"])), 20)
        assert np.abs(score.cpu().numpy() - z[f"score_{tag}"]).max() <= SCORE_TOL
        assert np.abs(a1.cpu().numpy() - z[f"att_1_{tag}"]).max() <= SCORE_TOL
    e.close()


@pytest.mark.parametrize("n,k", [(64, 20), (100, 10), (128, 20), (32, 10), (30, 7), (16, 10)])
def test_stage_parity_vs_oracle(eng, kitti_state, n, k):
    """Per-stage comparison: k-NN sets of all 6 EdgeConv layers, each layer's output, node embeddings, attention."""
    g = synth.make_graphs(6, n, k, seed=11)
    want = orc.embed_graphs(g, k, kitti_state, want_trace=True)
    got = eng.embed(_cuda(g), k, want_att=True, want_emb=True, trace=True)
    knn = got["knn"].cpu().long()
    layers = got["layers"].cpu()
    for layer in range(6):
        ok = orc.knn_sets_equivalent(want["knn_pd"][layer], want["knn_idx"][layer], knn[:, layer], want["layer_in"][layer])
        assert bool(ok.all()), f"k-NN set mismatch (not a tie) layer {layer}: rows {(~ok).nonzero()[:5].tolist()}"
        ref = want["layer_out"][layer].permute(0, 2, 1)                # [M, N, C']
        np.testing.assert_allclose(layers[:, layer, :, :ref.shape[2]].numpy(), ref.numpy(), atol=FEAT_ATOL, rtol=FEAT_RTOL)
    np.testing.assert_allclose(got["emb"].cpu().numpy(), want["emb"].numpy(), atol=FEAT_ATOL, rtol=FEAT_RTOL)
    np.testing.assert_allclose(got["att"].cpu().numpy(), want["att"].numpy(), atol=SCORE_TOL)
    np.testing.assert_allclose(got["pooled"].cpu().numpy(), want["pooled"].squeeze(-1).numpy(), atol=5e-5, rtol=1e-5)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_headline_batch_scores(eng, kitti_state, seed):
    """BASELINE config 2 shape: batch 128, 64-node graphs, k=20 — every score within 1e-5 of the oracle, except pairs
    where a k-NN near-tie (reference gap < 2e-6 relative) resolved the other way; those must be explained one by one."""
    f1, f2 = synth.make_pair_batch(128, 64, 20, seed=seed)
    want = orc.forward_pairs(f1, f2, 20, kitti_state)
    score, a1, a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), 20)
    from tests.helpers import assert_scores_match_or_near_tie
    bad = assert_scores_match_or_near_tie(eng, kitti_state, f1, f2, 20, score, want["score"], SCORE_TOL)
    good = torch.ones(128, dtype=torch.bool)
    good[bad] = False
    assert float((a1.cpu() - want["att_1"]).abs()[good].max()) <= SCORE_TOL
    assert float((a2.cpu() - want["att_2"]).abs()[good].max()) <= SCORE_TOL


def test_batch_independence_and_launch_determinism(eng):
    """A graph's embedding does not depend on its batch or position and repeats bit-exactly."""
    g = _cuda(synth.make_graphs(300, 64, 20, seed=3))
    full = eng.embed(g, 20)["pooled"]
    again = eng.embed(g, 20)["pooled"]
    assert torch.equal(full, again)
    solo = eng.embed(g[17:18].contiguous(), 20)["pooled"]
    assert torch.equal(solo[0], full[17])
    rev = eng.embed(g.flip(0).contiguous(), 20)["pooled"].flip(0)
    assert torch.equal(rev, full)


def test_pad_collapse_is_bit_exact(eng, kitti_state, monkeypatch):
    """Collapsing the trailing all-zero pad nodes into one row (csrc/embed_kernel.cuh) must not change a single bit
    versus processing every node: compare against a context created with SGPR_NO_DEDUP=1."""
    from sg_pr_b200.engine import Engine
    monkeypatch.setenv("SGPR_NO_DEDUP", "1")
    plain = Engine(0)
    monkeypatch.delenv("SGPR_NO_DEDUP")
    plain.set_weights(kitti_state)
    for n, k, dense in ((64, 20, False), (100, 10, False), (32, 10, False), (64, 20, True)):
        g = synth.make_graphs(40, n, k, seed=31, dense=dense)
        g[3].zero_()                       # an all-pad graph
        g[4, :, n // 2:].zero_()           # pads start mid-way
        g[5, :, 1:].zero_()                # a single real node
        a = eng.embed(_cuda(g), k, want_att=True, want_emb=True)
        b = plain.embed(_cuda(g), k, want_att=True, want_emb=True)
        for key in ("pooled", "att", "emb"):
            assert torch.equal(a[key], b[key]), (n, k, dense, key)
    plain.close()


def test_pairs_equal_embed_plus_head(eng):
    """forward_pairs == embed + score_pairs == score_matrix entry, bit for bit in the pair head's own arithmetic."""
    f1, f2 = synth.make_pair_batch(40, 64, 20, seed=5)
    score, _, _ = eng.forward_pairs(_cuda(f1), _cuda(f2), 20)
    pooled = eng.embed(_cuda(torch.cat([f1, f2])), 20)["pooled"]
    idx = torch.stack([torch.arange(40), torch.arange(40) + 40], dim=1).cuda()
    assert torch.equal(eng.score_pairs(pooled, idx), score)
    mat = eng.score_matrix(pooled[:40].contiguous(), pooled[40:].contiguous())
    assert float((mat.diagonal() - score).abs().max()) <= 2e-6
    # the score is NOT symmetric in its arguments (layers_batch.py:78-81)
    swapped, _, _ = eng.forward_pairs(_cuda(f2), _cuda(f1), 20)
    assert float((swapped - score).abs().max()) > 1e-4


def test_score_matrix_vs_oracle(eng, kitti_state):
    g = synth.make_graphs(70, 64, 20, seed=9)
    pooled = eng.embed(_cuda(g), 20)["pooled"]
    mat = eng.score_matrix(pooled[:33].contiguous(), pooled).cpu()
    want = orc.score_matrix(pooled[:33].cpu(), pooled.cpu(), kitti_state)
    assert float((mat - want).abs().max()) <= SCORE_TOL
    # ragged tile edges: 1 row, 1 column, and a strided output view
    one = eng.score_matrix(pooled[5:6].contiguous(), pooled[7:8].contiguous())
    assert abs(float(one[0, 0]) - float(want[5, 7])) <= SCORE_TOL
    big = torch.zeros(33, 100, device="cuda")
    eng.score_matrix(pooled[:33].contiguous(), pooled, out=big[:, 10:80])
    assert torch.equal(big[:, 10:80].cpu(), mat) and float(big[:, :10].abs().max()) == 0.0


def test_compact_input_is_bit_identical(eng):
    """13-byte-per-node records (xyz + uint8 label) expanded in shared memory give the SAME bits as the
    one-hot [15, N] blocks — device-resident, and pinned host records read in place over PCIe."""
    from sg_pr_b200.engine import compact_graphs
    for n, k, b in ((64, 20, 300), (100, 10, 40), (16, 10, 33), (128, 20, 20), (30, 7, 9)):
        f1, f2 = synth.make_pair_batch(b, n, k, seed=40 + n)
        f2[0].zero_()                                   # an all-pad graph
        c1, c2 = compact_graphs(f1), compact_graphs(f2)
        assert c1.shape == (b, eng.compact_stride(n)) and c1.shape[1] * 4 < 15 * n * 4
        want = eng.forward_pairs(_cuda(f1), _cuda(f2), k)
        got = eng.forward_pairs_compact(_cuda(c1), _cuda(c2), n, k)
        pinned = eng.forward_pairs_compact(c1.pin_memory(), c2.pin_memory(), n, k)
        for w, g, p in zip(want, got, pinned):
            assert torch.equal(w, g) and torch.equal(w, p), (n, k)
        e0 = eng.embed(_cuda(f1), k, want_att=True)
        e1 = eng.embed_compact(_cuda(c1), n, k, want_att=True)
        assert torch.equal(e0["pooled"], e1["pooled"]) and torch.equal(e0["att"], e1["att"])
    with pytest.raises(ValueError):
        compact_graphs(torch.rand(2, 15, 64))           # label rows that are not one-hot have no compact form
    with pytest.raises(ValueError):
        eng.forward_pairs_compact(torch.zeros(2, 100, dtype=torch.uint8, device="cuda"),
                                  torch.zeros(2, 100, dtype=torch.uint8, device="cuda"), 64, 20)


@pytest.mark.parametrize("env", [{}, {"SGPR_SCOREMAT_V2": "1"}, {"SGPR_SCOREMAT_FFMA": "1"}])
def test_score_matrix_kernel_variants_agree(kitti_state, monkeypatch, env):
    """The three score-matrix kernels — tcgen05 3xTF32 bilinear + FFMA2 epilogue (default), the same with FC1 on the tensor
    cores from a TMEM-resident A operand (SGPR_SCOREMAT_V2), and the fp32-FMA kernel (SGPR_SCOREMAT_FFMA) — against the
    oracle at ragged shapes, and a 1000 x 1000 block against the fused pair kernel's own head."""
    from sg_pr_b200.engine import Engine
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    e = Engine(0)
    for key in env:
        monkeypatch.delenv(key)
    e.set_weights(kitti_state)
    g = synth.make_graphs(1000, 64, 20, seed=19)
    pooled = e.embed(_cuda(g), 20)["pooled"]
    for r, m in ((1, 1), (7, 129), (17, 300), (130, 257)):
        mat = e.score_matrix(pooled[:r].contiguous(), pooled[5:5 + m].contiguous()).cpu()
        want = orc.score_matrix(pooled[:r].cpu(), pooled[5:5 + m].cpu(), kitti_state)
        assert float((mat - want).abs().max()) <= SCORE_TOL, (env, r, m)
    big = e.score_matrix(pooled, pooled)
    idx = synth.make_sequence_pairs(1000, 4096, seed=2).cuda()
    pairs = e.score_pairs(pooled, idx)
    assert float((big[idx[:, 0], idx[:, 1]] - pairs).abs().max()) <= 4e-6, env
    e.close()


def test_tensor_core_variant_of_the_fused_kernel(kitti_state, golden_dir, monkeypatch):
    """csrc/embed_tc_kernel.cuh (opt-in, SGPR_EMBED_TC=1): the 64-channel Gram / GEMM tiles as 3xTF32 tcgen05 UMMAs with the
    weights as a TMEM-resident A operand.  Same parity bar as the default kernel: golden batch, live oracle with near ties
    explained, compact input, persistent launch."""
    from sg_pr_b200.engine import Engine, compact_graphs
    from tests.helpers import assert_scores_match_or_near_tie
    monkeypatch.setenv("SGPR_EMBED_TC", "1")
    tc = Engine(0)
    monkeypatch.delenv("SGPR_EMBED_TC")
    tc.set_weights(kitti_state)
    with np.load(os.path.join(golden_dir, "ref_synth_n64_k20.npz")) as z:
        f1, f2 = torch.from_numpy(z["features_1"]), torch.from_numpy(z["features_2"])
        score, a1, _ = tc.forward_pairs(_cuda(f1), _cuda(f2), 20)
        assert np.abs(score.cpu().numpy() - z["score"]).max() <= SCORE_TOL
        assert np.abs(a1.cpu().numpy() - z["att_1"]).max() <= SCORE_TOL
    for n, k, b in ((64, 20, 128), (40, 10, 64), (33, 8, 40)):
        f1, f2 = synth.make_pair_batch(b, n, k, seed=70 + n)
        want = orc.forward_pairs(f1, f2, k, kitti_state)
        score, _, _ = tc.forward_pairs(_cuda(f1), _cuda(f2), k)
        assert_scores_match_or_near_tie(tc, kitti_state, f1, f2, k, score, want["score"], SCORE_TOL)
        again, _, _ = tc.forward_pairs_compact(_cuda(compact_graphs(f1)), _cuda(compact_graphs(f2)), n, k)
        assert torch.equal(again, score)
    g = _cuda(synth.make_graphs(700, 64, 20, seed=3))          # persistent launch, heaviest-first order
    full = tc.embed(g, 20)["pooled"]
    assert torch.equal(tc.embed(g[17:18].contiguous(), 20)["pooled"][0], full[17])
    tc.close()


def test_branch_split_launch_is_bit_identical(kitti_state, monkeypatch):
    """Small launches run one EdgeConv branch of a graph per work unit (EmbedArgs::split; api.cu splits while G <= 2/3 of the
    resident CTA slots, 197 on a B200).  Against whole-graph units (SGPR_NO_SPLIT=1): every output bit for bit — across the
    threshold, for the three row-tiling widths, compact records, the embed entry and the CPU tie rule."""
    from sg_pr_b200.engine import Engine, compact_graphs
    split = Engine(0)
    monkeypatch.setenv("SGPR_NO_SPLIT", "1")
    whole = Engine(0)
    monkeypatch.delenv("SGPR_NO_SPLIT")
    for e in (split, whole):
        e.set_weights(kitti_state)
    for b, n, k in ((1, 64, 20), (16, 64, 20), (37, 64, 20), (64, 64, 20), (98, 64, 20), (99, 64, 20), (24, 32, 10), (12, 128, 20)):
        f1, f2 = (_cuda(t) for t in synth.make_pair_batch(b, n, k, seed=300 + b))
        for _ in range(2):                                   # twice: the arrival counters must be back at zero
            got, ref = split.forward_pairs(f1, f2, k), whole.forward_pairs(f1, f2, k)
            assert all(torch.equal(x, y) for x, y in zip(got, ref)), (b, n, k)
        again = split.forward_pairs_compact(_cuda(compact_graphs(f1.cpu())), _cuda(compact_graphs(f2.cpu())), n, k)
        assert torch.equal(again[0], ref[0])
    g = _cuda(synth.make_graphs(150, 64, 20, seed=5))
    a, b = split.embed(g, 20, want_att=True, want_emb=True), whole.embed(g, 20, want_att=True, want_emb=True)
    assert all(torch.equal(a[key], b[key]) for key in ("pooled", "att", "emb"))
    for e in (split, whole):
        e.set_knn_ties("cpu")
    f1, f2 = (_cuda(t) for t in synth.make_pair_batch(8, 64, 20, seed=9, dense=True))
    assert torch.equal(split.forward_pairs(f1, f2, 20)[0], whole.forward_pairs(f1, f2, 20)[0])
    split.close()
    whole.close()


def test_host_entry_point_matches_device(eng):
    f1, f2 = synth.make_pair_batch(64, 64, 20, seed=21)
    d_score, d_a1, d_a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), 20)
    h_score, h_a1, h_a2 = eng.forward_pairs_host(f1, f2, 20)
    assert torch.equal(h_score, d_score.cpu()) and torch.equal(h_a1, d_a1.cpu()) and torch.equal(h_a2, d_a2.cpu())
    # pinned buffers: the kernel reads the inputs and writes the scores in place (zero-copy) — same bits
    out = (torch.empty(64).pin_memory(), torch.empty(64, 64, 1).pin_memory(), torch.empty(64, 64, 1).pin_memory())
    p_score, p_a1, p_a2 = eng.forward_pairs_host(f1.pin_memory(), f2.pin_memory(), 20, out=out)
    assert torch.equal(p_score, h_score) and torch.equal(p_a1, h_a1) and torch.equal(p_a2, h_a2)
    z_score, z_a1, _ = eng.forward_pairs(f1.pin_memory(), f2.pin_memory(), 20)      # device outputs, pinned inputs
    assert torch.equal(z_score.cpu(), h_score) and torch.equal(z_a1.cpu(), h_a1)
    mixed, _, _ = eng.forward_pairs_host(f1.pin_memory(), f2, 20, want_att=False)     # one pageable side -> staged path
    assert torch.equal(mixed, h_score)


def test_edge_cases(eng, kitti_state):
    # empty batch
    s, a1, a2 = eng.forward_pairs(torch.zeros(0, 15, 64, device="cuda"), torch.zeros(0, 15, 64, device="cuda"), 20)
    assert s.shape == (0,) and a1.shape == (0, 64, 1)
    # single pair, k == N, k == 1, all-zero graph (every node a pad)
    for n, k in ((8, 8), (64, 1), (64, 64), (5, 3)):
        f1, f2 = synth.make_pair_batch(2, n, 1, seed=n)
        f2[1].zero_()
        want = orc.forward_pairs(f1, f2, k, kitti_state)
        score, a1, a2 = eng.forward_pairs(_cuda(f1), _cuda(f2), k)
        if k in (1, n):      # no selection ambiguity at all when k == N; k == 1 picks self (distance 0)
            assert float((score.cpu() - want["score"]).abs().max()) <= SCORE_TOL, (n, k)
        assert bool(torch.isfinite(score).all()) and float(score.min()) >= 0 and float(score.max()) <= 1
    # few pads (0 < #pads < k): between the KITTI shape and dense — both tie modes well-formed, cpu mode = CPU oracle
    f1, f2 = synth.make_pair_batch(6, 64, 20, seed=2, dense=True)
    f1[:, :, 58:] = 0.0
    f2[:, :, 50:] = 0.0
    want = orc.forward_pairs(f1, f2, 20, kitti_state)
    eng.set_knn_ties("cpu")
    try:
        score, _, _ = eng.forward_pairs(_cuda(f1), _cuda(f2), 20)
    finally:
        eng.set_knn_ties("cuda")
    assert float((score.cpu() - want["score"]).abs().max()) <= SCORE_TOL


def test_errors_are_loud(eng):
    from sg_pr_b200._lib import SgprError
    f = torch.zeros(2, 15, 64, device="cuda")
    with pytest.raises(SgprError):
        eng.forward_pairs(f, f, 65)            # k > N: topk raises in the reference (dgcnn.py:19)
    with pytest.raises(SgprError):
        eng.forward_pairs(f, f, 0)
    big = torch.zeros(1, 15, 129, device="cuda")
    with pytest.raises(SgprError):
        eng.forward_pairs(big, big, 10)        # beyond the shared-memory tiling
    with pytest.raises(ValueError):
        eng.forward_pairs(torch.zeros(2, 14, 64, device="cuda"), torch.zeros(2, 14, 64, device="cuda"), 10)
    from sg_pr_b200.engine import Engine
    fresh = Engine(0)
    with pytest.raises(SgprError):
        fresh.forward_pairs(f, f, 10)          # no weights yet
    fresh.close()
