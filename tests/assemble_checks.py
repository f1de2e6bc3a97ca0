"""Shared check of the device batch assembly + augmentation (sgpr_train_assemble): the kernel's random draws are fed to
the reference-shaped host code path (SGTrainer.transfer_to_torch(training=True) -> utils.py augmentations) in the order
that code consumes them, and the resulting features must match the kernel's output."""
import random
import types

import numpy as np
import torch


def host_replay(graphs, pair_idx, draws, jitter, monkeypatch):
    """features_1 rows [2P,15,N] built by the host path with the kernel's draws substituted for numpy's / random's."""
    from sg_pr_b200.sg_net import SGTrainer
    N = graphs.shape[2]
    trainer = SGTrainer.__new__(SGTrainer)
    trainer.args = types.SimpleNamespace(node_num=N, p_thresh=3)
    trainer.global_labels = {i: i for i in range(12)}
    trainer.number_of_labels = 12
    queue = {"uniform": [], "randn": [], "random": []}
    monkeypatch.setattr(np.random, "uniform", lambda *a, **k: queue["uniform"].pop(0))
    monkeypatch.setattr(np.random, "randn", lambda *a: queue["randn"].pop(0))
    monkeypatch.setattr(random, "random", lambda: queue["random"].pop(0))
    out = []
    for p, (ga, gb) in enumerate(pair_idx.tolist()):
        data = {"distance": 0.0}
        for side, g in ((1, ga), (2, gb)):
            blk = graphs[g].double().numpy()
            onehot = blk[3:]
            nodes = np.where(onehot.sum(0) > 0, onehot.argmax(0), -1).astype(np.float64)
            data[f"nodes_{side}"], data[f"centers_{side}"] = nodes.tolist(), blk[:3].T.copy()
        queue["random"].append(float(draws[2 * p, 0]))                       # sg_net.py:288 random.random() > 0.5
        for slot in (2 * p, 2 * p + 1):                                       # augment_data(xyz_1) then augment_data(xyz_2)
            d = draws[slot].double().numpy()
            queue["uniform"] += [float(d[1]), np.array([d[2]]), np.array([[d[6], d[7], d[8]]])]   # rotate, scale, shift
            queue["randn"] += [jitter[slot].double().numpy()[None], np.array([d[3], d[4], d[5]])]  # jitter, perturbation
        new = trainer.transfer_to_torch(data, True)
        out += [new["features_1"], new["features_2"]]
    assert not queue["uniform"] and not queue["randn"] and not queue["random"]     # every draw was consumed, in order
    return torch.from_numpy(np.array(out)).float()


def check_assemble(eng, device, monkeypatch, M=12, N=64, P=9, seed=1234, step=7):
    from sg_pr_b200 import synth
    a, b = synth.make_pair_batch(M // 2, N, 20, seed=21)
    graphs = torch.cat([a, b])
    gen = torch.Generator().manual_seed(3)
    pair_idx = torch.randint(0, M, (P, 2), generator=gen, dtype=torch.int32)
    out, draws, jitter = eng.assemble(graphs.to(device), pair_idx.to(device), seed, step, want_draws=True)
    out, draws, jitter = out.cpu(), draws.cpu(), jitter.cpu()
    want = host_replay(graphs, pair_idx, draws, jitter, monkeypatch)
    # label rows are copied, coordinates go through ~10 fp32 operations at metre scale (the host path uses float64)
    assert torch.equal(out[:, 3:], want[:, 3:])
    err = (out[:, :3] - want[:, :3]).abs()
    assert float(err.max()) <= 5e-5, float(err.max())
    # reproducible from (seed, step); different steps give different draws; both graphs of a pair share the flip draw
    again = eng.assemble(graphs.to(device), pair_idx.to(device), seed, step).cpu()
    assert torch.equal(again, out)
    other = eng.assemble(graphs.to(device), pair_idx.to(device), seed, step + 1).cpu()
    assert not torch.equal(other[:, :3], out[:, :3])
    assert torch.equal(draws[0::2, 0], draws[1::2, 0])
    return draws, jitter


def check_draw_statistics(draws, jitter):
    """Coarse distribution checks of the Philox draws (many slots): ranges and first two moments."""
    d = draws.double()
    n = d.shape[0]
    tol = 6.0 / np.sqrt(n)
    assert 0 <= float(d[:, 0].min()) and float(d[:, 0].max()) < 1 and abs(float(d[:, 0].mean()) - 0.5) < tol
    assert 0 <= float(d[:, 1].min()) and float(d[:, 1].max()) < 1 and abs(float(d[:, 1].mean()) - 0.5) < tol
    assert 0.8 <= float(d[:, 2].min()) and float(d[:, 2].max()) < 1.25 and abs(float(d[:, 2].mean()) - 1.025) < tol
    assert abs(float(d[:, 3:6].mean())) < tol and abs(float(d[:, 3:6].std()) - 1.0) < tol
    assert -0.3 <= float(d[:, 6:9].min()) and float(d[:, 6:9].max()) < 0.3 and abs(float(d[:, 6:9].mean())) < tol
    j = jitter.double().reshape(-1)
    assert abs(float(j.mean())) < 0.01 and abs(float(j.std()) - 1.0) < 0.01
    assert abs(float((j.abs() > 1.96).double().mean()) - 0.05) < 0.005
