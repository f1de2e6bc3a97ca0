#!/usr/bin/env python
"""tcgen05 score-matrix kernel vs the fp32-FMA kernel and the oracle, plus timings.  Run on the B200 under `timeout`."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgpr_oracle as orc
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
eng = Engine(0); eng.set_weights(sd)
os.environ["SGPR_SCOREMAT_FFMA"] = "1"
ffma = Engine(0); ffma.set_weights(sd)
del os.environ["SGPR_SCOREMAT_FFMA"]
graphs = synth.make_graphs(4000, 64, 20, seed=4)
pooled = torch.cat([eng.embed(graphs[i:i + 1000].cuda(), 20)["pooled"] for i in range(0, 4000, 1000)])
out = []
for r, m in ((1, 1), (16, 128), (33, 70), (5, 300), (250, 1000), (4000, 4000)):
    rows, cols = pooled[:r].contiguous(), pooled[:m].contiguous()
    a = eng.score_matrix(rows, cols); torch.cuda.synchronize()
    b = ffma.score_matrix(rows, cols); torch.cuda.synchronize()
    rec = {"R": r, "M": m, "umma_vs_ffma_max_abs": float((a - b).abs().max())}
    if r * m <= 300000:
        want = orc.score_matrix(rows.cpu(), cols.cpu(), sd)
        rec["umma_vs_oracle_max_abs"] = float((a.cpu() - want).abs().max())
        rec["ffma_vs_oracle_max_abs"] = float((b.cpu() - want).abs().max())
    bad = (a - b).abs() > 1e-5
    if bool(bad.any()):
        idx = bad.nonzero()[:8].tolist()
        rec["first_bad"] = [(i, j, float(a[i, j]), float(b[i, j])) for i, j in idx]
        rec["bad_count"] = int(bad.sum())
    for name, e in (("umma", eng), ("ffma", ffma)):
        buf = torch.empty(r, m, device="cuda")
        for _ in range(3): e.score_matrix(rows, cols, out=buf)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10): e.score_matrix(rows, cols, out=buf)
        ev1.record(); torch.cuda.synchronize()
        rec[name + "_ms"] = ev0.elapsed_time(ev1) / 10
    out.append(rec)
    print(json.dumps(rec), flush=True)
