#!/usr/bin/env python
"""Which k-NN tie rule does the reference follow on ITS native device?  (VERDICT r01 task 1a)

`dgcnn.knn` ends in `pd.topk(k)` (/root/reference/dgcnn.py:19).  On the one-hot semantic branch >= 88 % of the rows hold
an exact tie at the k-th value, so the rule that breaks it decides the score whenever a graph has fewer than k zero pads.
ATen's CPU topk breaks ties as an artefact of std::nth_element (emulated bit-for-bit by `knn_ties="cpu"`); this probe
records what ATen's CUDA topk does on the B200 and which rule each mode of the fused kernel therefore matches.

  A. `torch.topk` on cuda vs three candidate rules, per row, on tie-heavy inputs (one-hot distances, small-integer
     rows, duplicated dense rows): "lowest index first", "highest index first", "CPU nth_element order".
  B. the reference's module code as stock PyTorch ops on the B200 (TF32 off) vs the fused kernel in each tie mode, on
     DENSE synthetic batches (no pads: the regime where the rule matters) and on KITTI-shape batches (pads >= k: where
     it cannot matter): per-layer k-NN rows classified with oracle.classify_knn_rows, scores compared.
  C. the same kernel modes vs the CPU oracle (= the reference on CPU).

Prints one JSON document (committed as profiles/r02_tie_rule_probe.json).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import sgpr_oracle as orc
from sg_pr_b200 import dgcnn as our_dgcnn
from sg_pr_b200 import synth
from sg_pr_b200.engine import Engine

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = torch.device("cuda", 0)
SHAPES = ((16, 10), (32, 10), (64, 10), (64, 20), (100, 10), (128, 20))


def rule_sets(pd: torch.Tensor, k: int):
    """Per-row index sets [rows, k] (sorted ascending) under 'lowest index first' and 'highest index first'."""
    n = pd.shape[-1]
    ar = torch.arange(n)
    # stable sort by value descending: ties keep ascending index -> lowest-first; reversed input -> highest-first
    lo = torch.sort(pd, dim=-1, descending=True, stable=True)[1][..., :k]
    hi = (n - 1) - torch.sort(pd.flip(-1), dim=-1, descending=True, stable=True)[1][..., :k]
    del ar
    return lo.sort(dim=-1)[0], hi.sort(dim=-1)[0]


def part_a():
    out = []
    g = torch.Generator().manual_seed(1)
    for n, k in SHAPES:
        cases = {}
        feats = synth.make_graphs(64, n, k, seed=5, dense=True)[:, 3:, :]            # one-hot rows: pd in {0, -2}
        cases["one_hot_dense"] = orc.pairwise_neg_sqdist(feats).reshape(-1, n)
        feats = synth.make_graphs(64, n, min(k, n // 2), seed=6, dense=False)[:, 3:, :] if n - k >= 1 else feats
        cases["one_hot_padded"] = orc.pairwise_neg_sqdist(feats).reshape(-1, n)
        cases["small_int"] = -torch.randint(0, 4, (2048, n), generator=g).float()
        base = torch.randn(2048, max(2, n // 4), generator=g)
        cases["dup_dense"] = base[:, torch.randint(0, base.shape[1], (n,), generator=g)]
        for name, pd in cases.items():
            cuda_idx = pd.to(DEV).topk(k, dim=-1)[1].cpu()
            cpu_idx = pd.topk(k, dim=-1)[1]
            lo, hi = rule_sets(pd, k)
            cs = cuda_idx.sort(dim=-1)[0]
            thr = pd.sort(dim=-1, descending=True)[0][:, k - 1: k]
            tied_rows = ((pd == thr).sum(-1) > 1) & ((pd >= thr).sum(-1) > k)       # the tie actually crosses the cut
            # within ties, is the CUDA *output order* ascending in index? (sorted=True output; informative)
            vals = torch.gather(pd, -1, cuda_idx)
            asc = ((vals[:, 1:] != vals[:, :-1]) | (cuda_idx[:, 1:] > cuda_idx[:, :-1])).all(-1)
            out.append({"N": n, "k": k, "case": name, "rows": int(pd.shape[0]), "rows_tie_crosses_cut": int(tied_rows.sum()),
                        "cuda_eq_lowest_index_first": int((cs == lo).all(-1).sum()),
                        "cuda_eq_highest_index_first": int((cs == hi).all(-1).sum()),
                        "cuda_eq_cpu_topk": int((cs == cpu_idx.sort(dim=-1)[0]).all(-1).sum()),
                        "cpu_eq_lowest_index_first": int((cpu_idx.sort(dim=-1)[0] == lo).all(-1).sum()),
                        "cuda_output_order_ties_ascending": int(asc.sum())})
    return out


_rec = {}


def recording_knn(x, k):
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    idx = pd.topk(k=k, dim=-1)[1]
    _rec.setdefault("pd", []).append(pd.cpu())
    _rec.setdefault("idx", []).append(idx.cpu())
    _rec.setdefault("x", []).append(x.cpu())
    return idx


def classify(trace_ref, got_knn):
    """trace_ref: dict(pd, idx, x) lists of 6 layers; got_knn [B,6,N,k] -> per-code first-divergence row counts."""
    b = got_knn.shape[0]
    counts = {"rows": 0, "same": 0, "exact_tie_swap": 0, "near_tie": 0, "mismatch": 0}
    seen = torch.zeros(b, 2, dtype=torch.bool)
    for layer in range(6):
        code = orc.classify_knn_rows(trace_ref["pd"][layer], trace_ref["idx"][layer], got_knn[:, layer].long(),
                                     trace_ref["x"][layer])
        br = layer // 3
        fresh = ~seen[:, br]
        counts["rows"] += int(fresh.sum()) * code.shape[1]
        for c, name in ((0, "same"), (1, "exact_tie_swap"), (2, "near_tie"), (3, "mismatch")):
            counts[name] += int((code[fresh] == c).sum())
        seen[:, br] |= (code > 0).any(dim=1)
    return counts


def part_bc(sd):
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SG
    from sg_pr_b200.torch_baseline import forward_torch
    eng = Engine(0)
    eng.set_weights(sd)
    modes = ["cuda"] + (["cpu"] if hasattr(eng, "set_knn_ties") else [])
    our_dgcnn.knn = recording_knn
    out = []
    for n, k, dense in ((64, 20, True), (64, 20, False), (32, 10, True), (100, 10, True), (128, 20, True)):
        margs = sgpr_args()
        margs.K, margs.node_num, margs.gpu, margs.cuda = k, n, 0, "0"
        model = SG(margs, 12)
        model.load_state_dict(sd)
        model.cuda(0).eval()
        b = 64
        f1, f2 = synth.make_pair_batch(b, n, k, seed=77 + n, dense=dense)
        # ---- reference module code on the B200 ----
        _rec.clear()
        with torch.no_grad():
            gpu_score, _, _ = forward_torch(model, f1.to(DEV), f2.to(DEV))
        gpu_score = gpu_score.cpu()
        tr_gpu = [{key: _rec[key][s * 6:(s + 1) * 6] for key in ("pd", "idx", "x")} for s in (0, 1)]
        # ---- reference on CPU (oracle) ----
        want = orc.forward_pairs(f1, f2, k, sd, want_trace=True)
        tr_cpu = [{"pd": want[f"knn_pd_{s}"], "idx": want[f"knn_idx_{s}"], "x": want[f"layer_in_{s}"]} for s in ("1", "2")]
        row = {"N": n, "k": k, "dense": dense, "pairs": b,
               "ref_gpu_vs_ref_cpu_max_abs_dscore": float((gpu_score - want["score"]).abs().max())}
        for mode in modes:
            if hasattr(eng, "set_knn_ties"):
                eng.set_knn_ties(mode)
            score, _, _ = eng.forward_pairs(f1.to(DEV), f2.to(DEV), k)
            score = score.cpu()
            knn = [eng.embed(f.to(DEV), k, trace=True)["knn"].cpu() for f in (f1, f2)]
            for ref_name, ref_score, traces in (("ref_gpu", gpu_score, tr_gpu), ("ref_cpu", want["score"], tr_cpu)):
                cls = {"rows": 0, "same": 0, "exact_tie_swap": 0, "near_tie": 0, "mismatch": 0}
                for s in (0, 1):
                    c = classify(traces[s], knn[s])
                    for key in cls:
                        cls[key] += c[key]
                err = (score - ref_score).abs()
                row[f"kernel[{mode}]_vs_{ref_name}"] = {"max_abs_dscore": float(err.max()),
                                                        "pairs_within_1e-5": int((err <= 1e-5).sum()), "knn_rows": cls}
        out.append(row)
    if hasattr(eng, "set_knn_ties"):
        eng.set_knn_ties("cuda")
    eng.close()
    return out


def main():
    sd = orc.load_state_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_kitti.npz"))
    doc = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "A_topk_rule": part_a(), "B_C_model": part_bc(sd)}
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
