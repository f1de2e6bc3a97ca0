#!/usr/bin/env python
"""Tensor-core variant of the fused kernel (SGPR_EMBED_TC=1) vs the FFMA kernel and the oracle, plus launch times."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine
sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
os.environ["SGPR_EMBED_TC"] = "1"
tc = Engine(0); tc.set_weights(sd)
del os.environ["SGPR_EMBED_TC"]
ff = Engine(0); ff.set_weights(sd)
for n, k, b in ((64, 20, 64), (40, 10, 32), (33, 8, 16), (64, 20, 300)):
    f1, f2 = synth.make_pair_batch(b, n, k, seed=5 + b)
    a = tc.forward_pairs(f1.cuda(), f2.cuda(), k); torch.cuda.synchronize()
    r = ff.forward_pairs(f1.cuda(), f2.cuda(), k); torch.cuda.synchronize()
    rec = {"N": n, "k": k, "B": b, "tc_vs_ffma_score": float((a[0] - r[0]).abs().max()), "tc_vs_ffma_att": float((a[1] - r[1]).abs().max()),
           "pairs_over_1e-5": int(((a[0] - r[0]).abs() > 1e-5).sum())}
    if b <= 64:
        want = orc.forward_pairs(f1, f2, k, sd)
        rec["tc_vs_oracle"] = float((a[0].cpu() - want["score"]).abs().max())
    print(json.dumps(rec), flush=True)
for name, e in (("ffma", ff), ("tc", tc)):
    out = {"kernel": name}
    for B in (16, 74, 128, 256, 512):
        sets = max(2, min(64, 140_000_000 // (2 * B * 15 * 64 * 4)))
        data = [tuple(t.cuda() for t in synth.make_pair_batch(B, 64, 20, seed=s)) for s in range(sets)]
        for i in range(10): e.forward_pairs(*data[i % sets], 20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): e.forward_pairs(*data[i % sets], 20)
        e1.record(); torch.cuda.synchronize()
        out[f"B{B}_us"] = round(e0.elapsed_time(e1) / 200 * 1e3, 2)
    print(json.dumps(out), flush=True)
