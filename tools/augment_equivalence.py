#!/usr/bin/env python
"""Do the host-augmented (numpy, reference call order) and device-augmented (sgpr_train_assemble, Philox) training
paths learn alike?  Same synthetic dataset, same initial weights per seed, `--epochs` epochs each; prints the mean loss
of every epoch per arm and seed.  (The augmentation ARITHMETIC is pinned by tests/assemble_checks.py; this is the
distribution-level sanity check.)"""
import argparse, contextlib, io, json, os, random, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sg_pr_b200 import synth
from sg_pr_b200.parser_sg import sgpr_args
from sg_pr_b200.sg_net import SGTrainer, _DeviceAdam
from tests.helpers import write_fixture_tree

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=400)
ap.add_argument("--pairs", type=int, default=1280)
ap.add_argument("--epochs", type=int, default=4)
ap.add_argument("--seeds", type=int, default=3)
args = ap.parse_args()
root = tempfile.mkdtemp(prefix="sgpr_augeq_")
cfg = write_fixture_tree(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"), root)
g = synth.make_graphs(args.graphs, 64, 20, seed=7)
os.makedirs(f"{root}/seq", exist_ok=True)
rng = np.random.default_rng(0)
poses = rng.uniform(0, 300, (args.graphs, 2))
for i in range(args.graphs):
    n_real = int((g[i, 3:].sum(0) > 0).sum())
    pose = [0.0] * 12
    pose[3], pose[11] = float(poses[i, 0]), float(poses[i, 1])
    json.dump({"centers": g[i, :3, :n_real].T.tolist(), "nodes": g[i, 3:, :n_real].argmax(0).tolist(), "pose": pose},
              open(f"{root}/seq/{i}.json", "w"))
pairs = []
while len(pairs) < args.pairs:
    i, j = rng.integers(0, args.graphs, 2)
    if len(pairs) % 2 == 0:
        j = i
    if i == j or np.hypot(*(poses[i] - poses[j])) >= 20:
        pairs.append([f"{root}/seq/{i}.json", f"{root}/seq/{j}.json"])
os.makedirs(f"{root}/lists", exist_ok=True)
for seq in ("00", "08"):
    open(f"{root}/lists/{seq}.txt", "w").writelines(f"{os.path.basename(a)} {os.path.basename(b)}\n" for a, b in pairs)
batches = [pairs[i:i + 128] for i in range(0, len(pairs), 128)]
for seed in range(args.seeds):
    for device in (False, True):
        a = sgpr_args().load(cfg)
        a.K, a.node_num, a.batch_size, a.p_thresh = 20, 64, 128, 3
        a.device_augment, a.augment_seed = device, seed
        torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            t = SGTrainer(a, True)
        t.optimizer = _DeviceAdam(t)
        curve = [round(float(np.mean([t.process_batch(b, True)[0] for b in batches])), 4) for _ in range(args.epochs)]
        print(json.dumps({"seed": seed, "arm": "device_augment" if device else "host augment (reference order)",
                          "mean_loss_per_epoch": curve}), flush=True)
