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
# second baseline (BASELINE.md §3 "optional"): the same math as ~200 stock PyTorch launches per forward ON THE B200 — the
# unfused path the reference would run on this GPU (our differentiable eval-mode path: sg_pr_b200.torch_baseline.forward_torch under no_grad)
try:
    from sg_pr_b200.parser_sg import sgpr_args
    from sg_pr_b200.sg_net import SG
    from sg_pr_b200.torch_baseline import forward_torch
    margs = sgpr_args(); margs.K, margs.node_num, margs.gpu, margs.cuda = 20, 64, 0, "0"
    model = SG(margs, 12); model.load_state_dict(sd); model.cuda(0).eval()
    f1, f2 = synth.make_pair_batch(128, 64, 20, seed=1)
    f1, f2 = f1.cuda(), f2.cuda()
    with torch.no_grad():
        for _ in range(5): ref_score, _, _ = forward_torch(model, f1, f2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): forward_torch(model, f1, f2)
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fused, _, _ = eng.forward_pairs(f1, f2, 20)
    out.append({"N": 64, "k": 20, "B": 128, "kind": "kitti", "impl": "unfused stock PyTorch ops on the same B200 (eager, fp32)",
                "us": round(us, 1), "pairs_per_s": round(128 / us * 1e6),
                "max_abs_diff_vs_fused_kernel": float((ref_score - fused).abs().max())})
except Exception as ex:  # pragma: no cover
    out.append({"impl": "unfused stock PyTorch ops", "error": repr(ex)})
for r in out: print(json.dumps(r))
