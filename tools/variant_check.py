#!/usr/bin/env python
"""A build variant (SGPR_B200_LIB) against the oracle on small batches — a quick numerical gate for experiments."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
for n, k, b in ((64, 20, 48), (40, 10, 16), (33, 8, 8), (100, 10, 8)):
    f1, f2 = synth.make_pair_batch(b, n, k, seed=5 + b)
    got = eng.forward_pairs(f1.cuda(), f2.cuda(), k)
    want = orc.forward_pairs(f1, f2, k, sd)
    print(json.dumps({"lib": os.environ.get("SGPR_B200_LIB", "default"), "N": n, "k": k, "B": b,
                      "score_vs_oracle": float((got[0].cpu() - want["score"]).abs().max()),
                      "att_vs_oracle": float((got[1].cpu() - want["att_1"]).abs().max())}), flush=True)
