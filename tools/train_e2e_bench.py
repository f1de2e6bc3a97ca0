#!/usr/bin/env python
"""main_sg.py-shaped end-to-end training benchmark (BASELINE config 3): a directory of graph JSON files, a pair list,
batches of `--batch` listed pairs through `SGTrainer.process_batch(batch, True)` — the inner loop of fit()
(sg_net.py:355-364) — for one epoch.

Arms (one JSON line each; every arm trains on the device through sgpr_train_step):
  1. reference-shaped host prep: both JSON files re-read per listed pair, Python one-hot loop, per-pair numpy
     augmentation (what the reference's process_batch does on the host, sg_net.py:316-331)
  2. this repo's default host path: parsed files cached, vectorised one-hot, features_1 only (mirrored step);
     same RNG call order as the reference
  3. device_augment: graphs uploaded once, sgpr_train_assemble builds the augmented batch in HBM
Optionally (--cpu-steps K) the reference's own training step on the host CPU for K batches (oracle port)."""
import argparse, contextlib, io, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sg_pr_b200 import synth
from sg_pr_b200.parser_sg import sgpr_args
from sg_pr_b200.sg_net import SGTrainer, _DeviceAdam
from sg_pr_b200 import utils as U
from tests.helpers import write_fixture_tree

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=1000)
ap.add_argument("--pairs", type=int, default=2560)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--cpu-steps", type=int, default=0)
args = ap.parse_args()
root = tempfile.mkdtemp(prefix="sgpr_trainbench_")
cfg = write_fixture_tree(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"), root)
g = synth.make_graphs(args.graphs, 64, 20, seed=7)
os.makedirs(f"{root}/seq", exist_ok=True)
rng = np.random.default_rng(0)
poses = rng.uniform(0, 300, (args.graphs, 2))
for i in range(args.graphs):
    n_real = int((g[i, 3:].sum(0) > 0).sum())
    pose = [0.0] * 12
    pose[3], pose[11] = float(poses[i, 0]), float(poses[i, 1])
    with open(f"{root}/seq/{i}.json", "w") as f:
        json.dump({"centers": g[i, :3, :n_real].T.tolist(), "nodes": g[i, 3:, :n_real].argmax(0).tolist(), "pose": pose}, f)
pairs = []
while len(pairs) < args.pairs:                       # half positives (d <= 3 m: here the graph with itself), half d >= 20 m
    i, j = rng.integers(0, args.graphs, 2)
    if len(pairs) % 2 == 0:
        j = i
    if i == j or np.hypot(*(poses[i] - poses[j])) >= 20:
        pairs.append([f"{root}/seq/{i}.json", f"{root}/seq/{j}.json"])
os.makedirs(f"{root}/lists", exist_ok=True)
for seq in ("00", "08"):
    with open(f"{root}/lists/{seq}.txt", "w") as f:
        f.writelines(f"{os.path.basename(a)} {os.path.basename(b)}\n" for a, b in pairs)
batches = [pairs[i:i + args.batch] for i in range(0, len(pairs), args.batch)]


def make_trainer(**kw):
    a = sgpr_args().load(cfg)
    a.K, a.node_num, a.batch_size, a.p_thresh = 20, 64, args.batch, 3
    a.graph_pairs_dir, a.pair_list_dir = f"{root}/seq", f"{root}/lists"
    for k, v in kw.items():
        setattr(a, k, v)
    torch.manual_seed(0)                               # every arm starts from the same random initialisation
    with contextlib.redirect_stdout(io.StringIO()):
        t = SGTrainer(a, True)
    t.optimizer = _DeviceAdam(t)
    return t


def run(label, trainer, fn, warm=False, pipelined=False):
    fn(batches[0])                                     # warm-up: context, workspace, first upload
    if warm:                                           # epoch 2 onwards: every file parsed / every graph uploaded already
        for b in batches:
            fn(b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    losses = []
    for i, b in enumerate(batches):
        if pipelined and i + 1 < len(batches):         # what fit() does: the next batch is built while this step runs
            trainer._upcoming_batch = batches[i + 1]
        losses.append(fn(b)[0])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"arm": label, "listed_pairs": len(pairs), "batch": args.batch, "seconds": round(dt, 4),
                      "listed_pairs_per_s": round(len(pairs) / dt, 1), "ms_per_batch": round(dt / len(batches) * 1e3, 3),
                      "first_loss": round(losses[0], 4), "last_loss": round(losses[-1], 4),
                      "device_launches": trainer._train_engine.launch_count()}), flush=True)


t1 = make_trainer()
def reference_shaped(batch):                           # host side exactly as sg_net.py:316-331 does it
    eng = t1._device_trainer()
    f1, tg = [], []
    for pair in batch:
        data = U.process_pair(pair)                    # re-reads both files
        n1, c1 = t1._fit_node_count(data["nodes_1"], data["centers_1"])
        n2, c2 = t1._fit_node_count(data["nodes_2"], data["centers_2"])
        x1, x2 = c1[None].copy(), c2[None].copy()
        import random
        if random.random() > 0.5:
            x1[:, :, 0] = -x1[:, :, 0]; x2[:, :, 0] = -x2[:, :, 0]
        x1, x2 = t1.augment_data(x1), t1.augment_data(x2)
        def onehot(nodes):                             # the reference's Python loop (sg_net.py:270-275)
            out = np.zeros((len(nodes), 12))
            for r, v in enumerate(nodes):
                if v != -1: out[r, int(v)] = 1.0
            return out
        a = np.squeeze(np.concatenate((x1, onehot(n1)[None]), axis=2).transpose(0, 2, 1))
        b = np.squeeze(np.concatenate((x2, onehot(n2)[None]), axis=2).transpose(0, 2, 1))
        f1 += [a, b]
        tg += [1.0 if data["distance"] <= 3 else 0.0] * 2
    f2 = [f1[i ^ 1] for i in range(len(f1))]
    d1, d2 = torch.FloatTensor(np.array(f1)), torch.FloatTensor(np.array(f2))
    loss, pred = eng.step(d1.cuda(), d2.cuda(), torch.FloatTensor(tg).cuda(), 20, apply=True, mirrored=False)
    return loss.item(), pred.cpu().numpy()
run("reference-shaped host prep (re-read JSON, python one-hot, both sides) + device step", t1, reference_shaped)

t2 = make_trainer()
run("default, first epoch: cached parse + vectorised one-hot + mirrored device step (reference RNG order)", t2, lambda b: t2.process_batch(b, True))
run("default, later epochs", t2, lambda b: t2.process_batch(b, True), warm=True)
run("default, later epochs, as fit() drives it (next batch's host prep overlaps the device step)", t2,
    lambda b: t2.process_batch(b, True), warm=True, pipelined=True)

t3 = make_trainer(device_augment=True, augment_seed=1)
run("device_augment, first epoch (parses + uploads every graph once): sgpr_train_assemble + mirrored device step", t3, lambda b: t3.process_batch(b, True))
run("device_augment, later epochs", t3, lambda b: t3.process_batch(b, True), warm=True)

if args.cpu_steps:
    from oracle import sgpr_oracle_train as ort
    sd = {k: v.detach().cpu().clone() for k, v in t2.model.module.state_dict().items()}
    adam = ort.new_adam_state(sd)
    t0 = time.perf_counter()
    for b in batches[:args.cpu_steps]:
        f1, tg = [], []
        for pair in b:
            d = t2.transfer_to_torch(U.process_pair(pair), True)
            f1 += [d["features_1"], d["features_2"]]; tg += [d["target"]] * 2
        f2 = [f1[i ^ 1] for i in range(len(f1))]
        ort.train_step(sd, torch.FloatTensor(np.array(f1)), torch.FloatTensor(np.array(f2)), torch.FloatTensor(tg), 20, adam, 1e-3, 5e-4)
    dt = time.perf_counter() - t0
    print(json.dumps({"arm": "reference training step on the host CPU (oracle port, autograd + Adam)", "threads": torch.get_num_threads(),
                      "batches": args.cpu_steps, "listed_pairs_per_s": round(args.cpu_steps * args.batch / dt, 2),
                      "ms_per_batch": round(dt / args.cpu_steps * 1e3, 1)}), flush=True)
