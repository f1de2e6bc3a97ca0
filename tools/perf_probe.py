"""Kernel-time probe: fused forward at several batch sizes / input kinds, CUDA-event timed (device-resident inputs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)

def timeit(f1, f2, k, iters=200):
    for _ in range(10): eng.forward_pairs(f1, f2, k, want_att=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): eng.forward_pairs(f1, f2, k, want_att=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

out = []
for n, k in ((64, 20),):
    for B in (16, 32, 64, 74, 128, 148, 256, 512, 1024, 4096):
        f1, f2 = synth.make_pair_batch(B, n, k, seed=1)
        us = timeit(f1.cuda(), f2.cuda(), k)
        out.append({"N": n, "k": k, "B": B, "kind": "kitti", "us": round(us, 2), "pairs_per_s": round(B / us * 1e6)})
    f1, f2 = synth.make_pair_batch(128, n, k, seed=1, dense=True)
    us = timeit(f1.cuda(), f2.cuda(), k)
    out.append({"N": n, "k": k, "B": 128, "kind": "dense", "us": round(us, 2), "pairs_per_s": round(128 / us * 1e6)})
for n, k, B in ((16, 10, 512), (32, 10, 512), (64, 10, 512), (64, 20, 512), (100, 10, 128), (128, 10, 512), (128, 20, 512)):
    f1, f2 = synth.make_pair_batch(B, n, k, seed=2)
    us = timeit(f1.cuda(), f2.cuda(), k, iters=50)
    out.append({"N": n, "k": k, "B": B, "kind": "kitti", "us": round(us, 2), "pairs_per_s": round(B / us * 1e6)})
for r in out: print(json.dumps(r))
