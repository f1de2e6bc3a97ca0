#!/usr/bin/env python
"""A few 4000 x 4000 score-matrix calls for ncu (tools/summarize_ncu.py reads the report)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
g = torch.Generator().manual_seed(0)
pooled = (torch.randn(4000, 32, generator=g) * 8).cuda()
out = torch.empty(4000, 4000, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng.score_matrix(pooled, pooled, out=out)
torch.cuda.synchronize()
print("done", float(out.sum()))
