#!/usr/bin/env python
"""All-pairs sequence scan benchmark (BASELINE config 4): M-graph synthetic sequence, M x M ordered pairs, row-block
sharded over the ranks of one box with NCCL all-gathers (pooled vectors, then score rows).

    python tools/scan_bench.py --graphs 4000                                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/scan_bench.py --graphs 4000

Prints one JSON line on rank 0: ordered pairs/s for the whole scan (embed + head + collectives), device-timed, max over
ranks, plus the split between the phases and a parity spot-check against the fused pair kernel."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sg_pr_b200 import scan, synth
from sg_pr_b200.engine import Engine
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=4000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
with np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz")) as z:
    state = {k: torch.from_numpy(z[k].copy()) for k in z.files}
eng = Engine(local); eng.set_weights(state)
sc = scan.SequenceScanner(eng, rank, world)
M, N, K = args.graphs, 64, 20
graphs = synth.make_graphs(M, N, K, seed=42).to(dev)

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

for _ in range(args.warmup): mat, (lo, hi) = sc.scan(graphs, K)
sync()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(args.steps): mat, (lo, hi) = sc.scan(graphs, K)
ev[1].record(); sync()
total_ms = ev[0].elapsed_time(ev[1]) / args.steps
# phase split (rank-local, single shot each)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
sync(); e[0].record(); pooled = eng.embed(graphs[lo:hi].contiguous(), K)["pooled"]; e[1].record()
allp = scan._all_gather_rows(pooled, M, world); e[2].record()
blk = eng.score_matrix(pooled, allp); e[3].record(); sync()
t = torch.tensor([total_ms, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    idx = synth.make_sequence_pairs(M, 256, seed=9)
    fused, _, _ = eng.forward_pairs(graphs[idx[:, 0].to(dev)], graphs[idx[:, 1].to(dev)], K)
    err = float((mat[idx[:, 0].to(dev), idx[:, 1].to(dev)] - fused).abs().max())
    total_ms, embed_ms, gather_ms, head_ms = (float(x) for x in t.tolist())
    print(json.dumps({"metric": "ordered graph-pairs/sec, all-pairs sequence scan", "value": M * M / (total_ms * 1e-3),
                      "unit": "graph-pairs/s", "n_gpus": world, "graphs": M, "node_num": N, "k": K, "ms_per_scan": total_ms,
                      "phases_ms": {"embed_row_block": embed_ms, "allgather_pooled": gather_ms, "score_row_block": head_ms,
                                    "allgather_scores_and_rest": max(0.0, total_ms - embed_ms - gather_ms - head_ms)},
                      "score_matrix_bytes": M * M * 4, "max_abs_diff_vs_fused_pair_kernel": err}))
if world > 1: dist.destroy_process_group()
