#!/usr/bin/env python
"""Launch time of the fused kernel vs graph size: every graph of the batch with the same number of real nodes
(KITTI-shape otherwise), B = 16 (one CTA per SM) and B = 128 (two CTAs on 108 of the 148 SMs).  Tells how much of the
B = 128 launch is the size mix of the two graphs that happen to share an SM."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0)
eng.set_weights(sd)


def batch(B, n_real, seed):
    g = synth.make_graphs(2 * B, 64, 20, seed=seed, dense=True)
    if n_real is None:
        g = synth.make_graphs(2 * B, 64, 20, seed=seed)
    else:
        g[:, :, n_real:] = 0.0
    return g[:B].contiguous().cuda(), g[B:].contiguous().cuda()


def timed(B, n_real):
    sets = [batch(B, n_real, s) for s in range(48)]
    for i in range(10):
        eng.forward_pairs(*sets[i % 48], 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        eng.forward_pairs(*sets[i % 48], 20)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / 200 * 1e3, 2)


for n in (None, 25, 30, 31, 35, 39, 40, 44):
    print(json.dumps({"n_real": "U{25..44}" if n is None else n, "B16_us": timed(16, n), "B128_us": timed(128, n), "B148_us": timed(148, n)}), flush=True)
