#!/usr/bin/env python
"""Training-step parity at scale: many synthetic batches, each run through sgpr_train_step (gradients only) and through
the oracle's autograd on the host CPU from the same state; reports the distribution of prediction and gradient
deviations.  A k-NN near-tie flip in any graph moves the whole batch through the batch statistics (training
parity bar), so deviations are reported per batch together with how many k-NN rows differ.

    python tools/train_parity_report.py [--batches 12] [--listed 32] [--nodes 64] [--k 20] > profiles/<tag>_train_parity_report.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sgpr_oracle as orc
from oracle import sgpr_oracle_train as ort
from oracle.make_golden_train import train_batch
from sg_pr_b200.train_engine import TrainEngine

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, default=12)
ap.add_argument("--listed", type=int, default=32)
ap.add_argument("--nodes", type=int, default=64)
ap.add_argument("--k", type=int, default=20)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sd = orc.load_state_npz(os.path.join(root, "tests", "golden", "model_kitti.npz"))
eng = TrainEngine(0)
rows, t0 = [], time.time()
for b in range(a.batches):
    f1, f2, target = train_batch(a.listed, a.nodes, a.k, seed=1000 + b)
    work = {n: v.clone() for n, v in sd.items()}
    want = ort.train_step(work, f1, f2, target, a.k, ort.new_adam_state(work), 0.0, 0.0, want_trace=True)
    eng.set_state(sd)
    _, pred = eng.step(f1.cuda(), None, target.cuda(), a.k, apply=False, mirrored=True)
    grads = eng.grads()
    G, N, k = f1.shape[0], a.nodes, a.k
    # classify every k-NN row against the oracle's (oracle.classify_knn_rows): 0 same nodes, 1 exact-tie swap,
    # 2 near tie, 3 mismatch.  Only the FIRST differing layer of a branch is meaningful: later layers of that branch see
    # different inputs.  Side 1 of the oracle == the mirrored step's only side.
    tr = want["aux"]["trace_1"]
    counts, first = [0, 0, 0, 0], {}
    for layer in range(6):
        got = torch.from_numpy(eng.debug_read("idx", layer, (G, N, k), np.uint8).astype(np.int64))
        code = orc.classify_knn_rows(tr["knn_pd"][layer], tr["knn_idx"][layer], got, tr["layer_in"][layer])
        for g_, i_ in (code > 0).nonzero().tolist():
            key = (g_, layer // 3)
            if key not in first:
                first[key] = layer
            if first[key] == layer:
                counts[int(code[g_, i_])] += 1
        counts[0] += int((code == 0).sum())
    # the same step by the oracle in float64: the distance between the reference's own fp32 and fp64 results is the noise
    # floor of this comparison (max / LeakyReLU / ReLU kinks route a gradient differently when a value sits within rounding
    # of a tie or of zero)
    w64 = {n: (v.double() if v.dtype.is_floating_point else v.clone()) for n, v in sd.items()}
    ref64 = ort.train_step(w64, f1.double(), f2.double(), target.double(), a.k, ort.new_adam_state(w64), 0.0, 0.0)
    rel, rel64, floor = {}, {}, {}
    for name, ref in want["grads"].items():
        scale = max(float(ref.abs().max()), 1e-12)
        r64 = ref64["grads"][name]
        rel[name] = float((grads[name].reshape(ref.shape) - ref).abs().max()) / scale
        rel64[name] = float((grads[name].reshape(ref.shape).double() - r64).abs().max()) / scale
        floor[name] = float((ref.double() - r64).abs().max()) / scale
    rows.append({"batch": b, "max_abs_dpred": float((pred.cpu() - want["pred"]).abs().max()),
                 "knn_first_divergences": {"exact_tie_swap": counts[1], "near_tie": counts[2], "mismatch": counts[3]},
                 "knn_rows_total": 6 * G * N,
                 "max_rel_dgrad": max(rel.values()), "worst_grad": max(rel, key=rel.get),
                 "max_rel_dgrad_vs_oracle_fp64": max(rel64.values()),
                 "oracle_fp32_vs_fp64_max_rel_dgrad": max(floor.values()),
                 "oracle_fp32_vs_fp64_max_abs_dpred": float((want["pred"].double() - ref64["pred"]).abs().max()),
                 "max_abs_dpred_vs_oracle_fp64": float((pred.cpu().double() - ref64["pred"]).abs().max())})
clean = [r for r in rows if sum(r["knn_first_divergences"].values()) == 0]
out = {"what": "sgpr_train_step (mirrored, gradients only) vs oracle autograd on the host CPU, same state",
       "config": {"listed_pairs": a.listed, "forward_pairs": 2 * a.listed, "nodes": a.nodes, "k": a.k, "batches": a.batches},
       "batches_without_knn_difference": len(clean),
       "max_abs_dpred_clean": max((r["max_abs_dpred"] for r in clean), default=None),
       "max_rel_dgrad_clean": max((r["max_rel_dgrad"] for r in clean), default=None),
       "max_abs_dpred_all": max(r["max_abs_dpred"] for r in rows), "max_rel_dgrad_all": max(r["max_rel_dgrad"] for r in rows),
       "noise_floor_oracle_fp32_vs_fp64": {"max_abs_dpred": max(r["oracle_fp32_vs_fp64_max_abs_dpred"] for r in rows),
                                           "max_rel_dgrad": max(r["oracle_fp32_vs_fp64_max_rel_dgrad"] for r in rows),
                                           "median_rel_dgrad": float(np.median([r["oracle_fp32_vs_fp64_max_rel_dgrad"] for r in rows]))},
       "kernel_vs_oracle_fp64": {"max_abs_dpred": max(r["max_abs_dpred_vs_oracle_fp64"] for r in rows),
                                 "max_rel_dgrad": max(r["max_rel_dgrad_vs_oracle_fp64"] for r in rows),
                                 "median_rel_dgrad": float(np.median([r["max_rel_dgrad_vs_oracle_fp64"] for r in rows]))},
       "kernel_vs_oracle_fp32_median_rel_dgrad": float(np.median([r["max_rel_dgrad"] for r in rows])),
       "knn_first_divergences_total": {key: sum(r["knn_first_divergences"][key] for r in rows)
                                       for key in ("exact_tie_swap", "near_tie", "mismatch")},
       "note": "a batch is 'clean' when every k-NN row selects the same nodes as the oracle (up to bit-identical "
               "duplicates); one swap / near tie in any graph moves the whole batch through the batch statistics",
       "seconds": round(time.time() - t0, 1), "per_batch": rows}
print(json.dumps(out, indent=1))
