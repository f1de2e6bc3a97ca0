#!/usr/bin/env python
"""Fused-kernel time at the batch sizes that matter for the headline (B = 16..256, N = 64, k = 20), CUDA-event timed,
device-resident inputs, rotating over more input sets than fit in L2.  One JSON line.  SGPR_B200_LIB selects a variant."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
out = {"lib": os.environ.get("SGPR_B200_LIB", "default")}
ref = None
for B in (16, 64, 74, 80, 96, 112, 128, 148, 256, 512):
    sets = max(2, min(64, 140_000_000 // (2 * B * 15 * 64 * 4)))
    data = [tuple(t.cuda() for t in synth.make_pair_batch(B, 64, 20, seed=s)) for s in range(sets)]
    for i in range(10): eng.forward_pairs(*data[i % sets], 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 200
    e0.record()
    for i in range(iters): eng.forward_pairs(*data[i % sets], 20)
    e1.record(); torch.cuda.synchronize()
    out[f"B{B}_us"] = round(e0.elapsed_time(e1) / iters * 1e3, 2)
    if B == 128:
        s, _, _ = eng.forward_pairs(*data[0], 20)
        out["checksum_B128"] = float(s.double().sum())
print(json.dumps(out))
