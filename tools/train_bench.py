"""Training-step benchmark (BASELINE config 3): `listed` synthetic KITTI-shape pairs -> 2*listed forward pairs (both
orders, sg_net.py:324-331), N nodes, k neighbours.  Times the device step (sgpr_train_step: 11 launches) with CUDA
events, inputs resident in HBM, and — optionally — the same step as stock PyTorch ops on the same GPU (the reference's
own module code path: sg_pr_b200.torch_baseline.forward_torch + loss.backward() + torch.optim.Adam).

    python tools/train_bench.py [--listed 128] [--nodes 64] [--k 20] [--steps 50] [--warmup 5] [--torch-baseline]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--listed", type=int, default=128)
    ap.add_argument("--nodes", type=int, default=64)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--torch-baseline", action="store_true")
    ap.add_argument("--two-sided", action="store_true", help="run both sides (default: mirrored, what process_batch does)")
    ap.add_argument("--cpu-baseline", action="store_true", help="one oracle step on the host CPU (slow: several seconds)")
    a = ap.parse_args()

    from oracle import sgpr_oracle as orc                      # only to read the committed checkpoint fixture
    from oracle.make_golden_train import train_batch
    from sg_pr_b200.train_engine import TrainEngine
    sd = orc.load_state_npz(os.path.join(ROOT, "tests", "golden", "model_kitti.npz"))
    f1, f2, target = train_batch(a.listed, a.nodes, a.k, seed=1)
    f1, f2, target = f1.cuda(), f2.cuda(), target.cuda()
    eng = TrainEngine(0)
    eng.set_state(sd)
    eng.set_optimizer(1e-3, 5e-4)
    mirrored = not a.two_sided
    for _ in range(a.warmup):
        eng.step(f1, f2, target, a.k, mirrored=mirrored)
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, pred = eng.step(f1, f2, target, a.k, mirrored=mirrored)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out = {"what": "sgpr_train_step (mirrored)" if mirrored else "sgpr_train_step (two-sided)", "listed_pairs": a.listed, "forward_pairs": 2 * a.listed, "nodes": a.nodes, "k": a.k,
           "ms_per_step": round(ms, 4), "listed_pairs_per_s": round(a.listed / ms * 1e3, 1),
           "launches_per_step": (eng.launch_count() - l0) // a.steps, "loss_after": float(loss)}
    print(json.dumps(out), flush=True)

    if a.torch_baseline:
        from sg_pr_b200.parser_sg import sgpr_args
        from sg_pr_b200.sg_net import SG
        from sg_pr_b200.torch_baseline import forward_torch
        args = sgpr_args()
        args.K, args.node_num, args.gpu, args.cuda = a.k, a.nodes, 0, "0"
        model = SG(args, 12)
        model.load_state_dict(sd)
        model.cuda().train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)

        def step():
            opt.zero_grad()
            pred, _, _ = forward_torch(model, f1, f2)
            loss = torch.mean(torch.nn.functional.binary_cross_entropy(pred, target))
            loss.backward()
            opt.step()
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = max(3, a.steps // 5)
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        tms = e0.elapsed_time(e1) / n
        print(json.dumps({"what": "stock PyTorch ops on the same GPU (fp32, TF32 off)", "ms_per_step": round(tms, 3),
                          "listed_pairs_per_s": round(a.listed / tms * 1e3, 1), "speedup": round(tms / ms, 1)}), flush=True)

    if a.cpu_baseline:
        from oracle import sgpr_oracle_train as ort
        sdc = {n: v.clone() for n, v in sd.items()}
        adam = ort.new_adam_state(sdc)
        c1, c2, ct = f1.cpu(), f2.cpu(), target.cpu()
        t0 = time.perf_counter()
        ort.train_step(sdc, c1, c2, ct, a.k, adam, 1e-3, 5e-4)
        dt = time.perf_counter() - t0
        print(json.dumps({"what": "oracle train step on the host CPU", "threads": torch.get_num_threads(),
                          "ms_per_step": round(dt * 1e3, 1), "listed_pairs_per_s": round(a.listed / dt, 2)}), flush=True)


if __name__ == "__main__":
    main()
