#!/usr/bin/env python
"""A few launches of the branch-split fused kernel (B pairs, default 16) for an ncu capture:
ncu --set full --clock-control none --import-source on -k regex:sgpr_embed -s 3 -c 1 -o out python tools/split_profile.py 16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for s in range(6):
    f1, f2 = synth.make_pair_batch(B, 64, 20, seed=s)
    eng.forward_pairs(f1.cuda(), f2.cuda(), 20)
torch.cuda.synchronize()
