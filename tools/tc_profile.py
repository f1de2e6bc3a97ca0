import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
f1, f2 = synth.make_pair_batch(128, 64, 20, seed=1)
f1, f2 = f1.cuda(), f2.cuda()
for _ in range(4): eng.forward_pairs(f1, f2, 20)
torch.cuda.synchronize()
