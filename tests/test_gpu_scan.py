"""GPU test of the embed-once / score-matrix sequence scan through the engine (single rank; the row-block sharding
is covered on CPU by tests/test_scan_gloo.py and on N GPUs by tools/scan_bench.py)."""
import pytest
import torch

from oracle import sgpr_oracle as orc
from sg_pr_b200 import scan, synth

pytestmark = pytest.mark.gpu


def test_scan_matches_pairwise_forward(kitti_state):
    from sg_pr_b200.engine import Engine
    eng = Engine(0)
    eng.set_weights(kitti_state)
    m = 150
    graphs = synth.make_graphs(m, 64, 20, seed=13)
    sc = scan.SequenceScanner(eng)
    mat, (lo, hi) = sc.scan(graphs, 20)
    assert (lo, hi) == (0, m) and mat.shape == (m, m)
    # spot-check 64 ordered pairs against the fused pair kernel and the oracle
    idx = synth.make_sequence_pairs(m, 64, seed=1)
    f1, f2 = graphs[idx[:, 0]].cuda(), graphs[idx[:, 1]].cuda()
    fused, _, _ = eng.forward_pairs(f1, f2, 20)
    picked = mat[idx[:, 0].cuda(), idx[:, 1].cuda()]
    assert float((picked - fused).abs().max()) <= 2e-6
    want = orc.forward_pairs(graphs[idx[:, 0]], graphs[idx[:, 1]], 20, kitti_state)["score"]
    from tests.helpers import assert_scores_match_or_near_tie
    assert_scores_match_or_near_tie(eng, kitti_state, graphs[idx[:, 0]], graphs[idx[:, 1]], 20, picked, want, 1e-5, max_bad=4)
    # world-size-1 NCCL path (the logic-only variant of the multi-GPU scan)
    vals, nbr, _ = sc.top_matches(graphs, 20, per_row=3, exclude_window=10)
    assert vals.shape == (m, 3) and int(nbr[-1].max()) <= m - 1 - 10
    eng.close()
