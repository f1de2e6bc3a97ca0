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


def _nccl_worker(rank, world, port, m, out_dir, exchange="nccl"):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from sg_pr_b200.engine import Engine
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        eng = Engine(rank)
        eng.set_weights(orc.load_state_npz(os.path.join(root, "tests", "golden", "model_kitti.npz")))
        graphs = synth.make_graphs(m, 64, 20, seed=13).cuda()
        sc = scan.SequenceScanner(eng, rank, world)
        full, (lo, hi) = sc.scan(graphs, 20, exchange=exchange)
        if exchange == "peer":                      # a second scan into the same peer-mapped buffers (reuse protocol)
            full, (lo, hi) = sc.scan(graphs, 20, exchange=exchange)
        torch.cuda.synchronize()
        torch.save({"full": full.cpu(), "lo": lo, "hi": hi}, os.path.join(out_dir, f"r{rank}.pt"))
        sc.close()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m,exchange", [(256, "nccl"), (301, "nccl"), (256, "peer"), (301, "peer")])
def test_two_rank_nccl_scan_is_bit_equal_to_single_gpu(kitti_state, tmp_path, m, exchange):
    """Config 4 on real ranks: two processes, one GPU each, NCCL — the sharded [M, M] matrix every rank ends up with is
    bit-equal to the single-GPU scan (the row block a rank computes does not depend on how many ranks there are).
    m = 256 / 301: equal row blocks (one in-place all-gather) / unequal ones (pad path); exchange "peer": the score kernel
    stores straight into both ranks' matrices over NVLink peer memory (CUDA IPC), no score collective."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    from sg_pr_b200.engine import Engine
    eng = Engine(0)
    eng.set_weights(kitti_state)
    single, _ = scan.SequenceScanner(eng).scan(synth.make_graphs(m, 64, 20, seed=13).cuda(), 20)
    single = single.cpu()
    eng.close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, m, str(tmp_path), exchange), nprocs=2, join=True)
    for r in range(2):
        o = torch.load(tmp_path / f"r{r}.pt")
        assert o["full"].shape == (m, m)
        assert torch.equal(o["full"], single), f"rank {r}: sharded scan differs from the single-GPU scan"
