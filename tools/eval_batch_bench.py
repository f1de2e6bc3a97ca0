#!/usr/bin/env python
"""eval_batch.py-shaped end-to-end benchmark: a sequence of M graph JSON files, a pair list of P
[path_a, path_b] lines, batches of 128 through `SGTrainer.eval_batch_pair` — the loop of eval_batch.py:30-36.

Arms: (1) embed-once + GraphStore (default), (2) GraphStore + fused pair kernel per batch (embed_cache off),
      (3) the reference-shaped host prep (process_pair + transfer_to_torch per pair, sg_net.py:503-519) + fused kernel.
Prints one JSON line per arm."""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sg_pr_b200 import synth
from sg_pr_b200.parser_sg import sgpr_args
from sg_pr_b200.sg_net import SGTrainer
from sg_pr_b200.utils import process_pair
from tests.helpers import write_fixture_tree

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=1000)
ap.add_argument("--pairs", type=int, default=12800)
args = ap.parse_args()
root = tempfile.mkdtemp(prefix="sgpr_evalbench_")
cfg = write_fixture_tree(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"), root)
g = synth.make_graphs(args.graphs, 64, 20, seed=7)
os.makedirs(f"{root}/seq", exist_ok=True)
rng = np.random.default_rng(0)
for i in range(args.graphs):
    n_real = int((g[i, 3:].sum(0) > 0).sum())
    pose = [0.0] * 12
    pose[3], pose[11] = float(rng.uniform(0, 3000)), float(rng.uniform(0, 3000))
    with open(f"{root}/seq/{i}.json", "w") as f:
        json.dump({"centers": g[i, :3, :n_real].T.tolist(), "nodes": g[i, 3:, :n_real].argmax(0).tolist(), "pose": pose}, f)
a = sgpr_args().load(cfg)
a.K, a.node_num, a.batch_size, a.p_thresh = 20, 64, 128, 3
trainer = SGTrainer(a, False)
trainer.model.eval()
pairs = []
while len(pairs) < args.pairs:                       # keep only pairs the reference accepts (d <= 3 m or d >= 20 m)
    i, j = rng.integers(0, args.graphs, 2)
    pa, pb = f"{root}/seq/{i}.json", f"{root}/seq/{j}.json"
    d = trainer._store().distance(pa, pb)
    if d <= 3 or d >= 20:
        pairs.append([pa, pb])
batches = [pairs[i:i + 128] for i in range(0, len(pairs), 128)]

def run(label, fn, fresh_store=True):
    if fresh_store:
        trainer._graph_store = None; trainer._emb = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = [fn(b) for b in batches]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pred = np.concatenate([o[0] for o in out])
    print(json.dumps({"arm": label, "pairs": len(pairs), "graphs": args.graphs, "seconds": dt, "pairs_per_s": len(pairs) / dt}), flush=True)
    return pred

trainer.embed_cache = True
p1 = run("embed-once + GraphStore (cold: parses every file once)", trainer.eval_batch_pair)
p1b = run("embed-once + GraphStore (warm)", trainer.eval_batch_pair, fresh_store=False)
trainer.embed_cache = False
p2 = run("GraphStore + fused pair kernel per batch", trainer.eval_batch_pair)

def reference_shaped(batch):                          # sg_net.py:503-525 as written: re-read + re-pad both files of every pair
    f1, f2, tg = [], [], []
    for pair in batch:
        data = trainer.transfer_to_torch(process_pair(pair), False)
        f1.append(data["features_1"]); f2.append(data["features_2"]); tg.append(data["target"])
    data = {"features_1": torch.FloatTensor(np.array(f1)), "features_2": torch.FloatTensor(np.array(f2))}
    with torch.no_grad():
        pred, _, _ = trainer.model(data)
    return pred.cpu().numpy().reshape(-1), np.array(tg)
p3 = run("reference-shaped host prep + fused pair kernel", reference_shaped)
print(json.dumps({"max_abs_diff_embed_once_vs_fused": float(np.abs(p1 - p2).max()), "max_abs_diff_vs_reference_shaped_prep": float(np.abs(p1 - p3).max())}))
