#!/usr/bin/env python
"""One call through each launch shape of the fused kernel (branch-split sole / shared, whole-graph, persistent) and the
tcgen05 score matrix, checked against the oracle — meant to run under compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
for b, what in ((1, "split, one unit per SM"), (50, "split, shared SMs"), (110, "whole graphs"), (160, "persistent")):
    f1, f2 = synth.make_pair_batch(b, 64, 20, seed=b)
    got = eng.forward_pairs(f1.cuda(), f2.cuda(), 20)[0].cpu()
    want = orc.forward_pairs(f1[:4], f2[:4], 20, sd)["score"]
    print(what, "max |dscore| on 4 pairs", float((got[:4] - want).abs().max()), flush=True)
g = synth.make_graphs(90, 64, 20, seed=2).cuda()
pooled = eng.embed(g, 20)["pooled"]
m = eng.score_matrix(pooled[:40], pooled)
torch.cuda.synchronize()
print("score matrix", tuple(m.shape), float(m.sum()))
