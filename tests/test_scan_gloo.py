"""World-size-2 (gloo, CPU) test of the row-block sharding + all-gather logic of the all-pairs sequence scan
(sg_pr_b200/scan.py).  The compute callables are the ORACLE's here (this is a test of the host-side sharding logic;
the CUDA engine is exercised by tests/test_gpu_scan.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sg_pr_b200 import scan, synth


def test_row_blocks_partition_everything():
    for m in (0, 1, 7, 8, 9, 4000, 4001):
        for world in (1, 2, 3, 8):
            blocks = [scan.row_block(m, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sgpr_oracle as orc
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sd = orc.load_state_npz(os.path.join(root, "tests", "golden", "model_kitti.npz"))
        graphs = synth.make_graphs(m, 32, 10, seed=3)
        embed = lambda g, k: orc.embed_graphs(g, k, sd)["pooled"].squeeze(-1) if g.shape[0] else torch.empty(0, 32)
        score = lambda rows, cols, out=None: orc.score_matrix(rows, cols, sd) if rows.shape[0] else torch.empty(0, cols.shape[0])
        full, (lo, hi) = scan.scan_all_pairs(graphs, 10, embed, score, rank, world)
        lo2, hi2 = scan.row_block(m, rank, world)
        local, _ = scan.scan_all_pairs(graphs[lo2:hi2], 10, embed, score, rank, world, gather_scores=False,
                                       graphs_are_local=True)
        torch.save({"full": full, "lo": lo, "hi": hi, "local": local}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


# unequal blocks (pad path) and equal blocks (direct, one collective) on two ranks; three ranks with a ragged last block
@pytest.mark.parametrize("world,m", [(2, 9), (2, 12), (3, 10)])
def test_sharded_scan_matches_single_rank(tmp_path, world, m):
    from oracle import sgpr_oracle as orc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = orc.load_state_npz(os.path.join(root, "tests", "golden", "model_kitti.npz"))
    graphs = synth.make_graphs(m, 32, 10, seed=3)
    pooled = orc.embed_graphs(graphs, 10, sd)["pooled"].squeeze(-1)
    want = orc.score_matrix(pooled, pooled, sd)
    single, _ = scan.scan_all_pairs(graphs, 10, lambda g, k: orc.embed_graphs(g, k, sd)["pooled"].squeeze(-1),
                                    lambda r, c, out=None: orc.score_matrix(r, c, sd))
    assert torch.equal(single, want)
    mp.spawn(_worker, args=(world, _free_port(), m, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for r, o in enumerate(outs):
        assert o["full"].shape == (m, m)
        assert torch.allclose(o["full"], want, atol=1e-6), r       # every rank ends up with the whole matrix
        assert torch.allclose(o["local"], want[o["lo"]:o["hi"]], atol=1e-6)
    assert (outs[0]["lo"], outs[-1]["hi"]) == (0, m) and all(outs[r]["hi"] == outs[r + 1]["lo"] for r in range(world - 1))
