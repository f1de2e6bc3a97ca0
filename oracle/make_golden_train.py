"""Generate tests/golden/ref_train_<cfg>.npz: two optimiser steps of the UNMODIFIED reference model on CPU.

Run in the build container only:   python oracle/make_golden_train.py
TEST INFRASTRUCTURE — never imported by the product.

What is replayed is exactly `SGTrainer.process_batch(batch, training=True)` (sg_net.py:312-343) from the point where
the batch tensors exist — the host-side augmentation before that point draws from numpy's global RNG
(sg_net.py:223-230, 288-294) and is not part of the device path: `optimizer.zero_grad()`, `model(data)` in train
mode (batch-statistics BatchNorm, each side normalised separately because `dgcnn_conv_pass` is called once per side,
sg_net.py:123-124), `mean(binary_cross_entropy)`, `backward`, `Adam(lr, weight_decay).step()` (sg_net.py:351-352).
The batch has process_batch's structure: every listed pair (a, b) appears as (a, b) and (b, a) with the same target.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from sg_pr_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def train_batch(listed: int, N: int, K: int, seed: int):
    """process_batch's doubling (sg_net.py:324-331) of `listed` synthetic pairs; alternating targets 1/0."""
    a, b = synth.make_pair_batch(listed, N, K, seed=seed)
    f1 = torch.stack([a, b], dim=1).reshape(2 * listed, *a.shape[1:]).contiguous()
    f2 = torch.stack([b, a], dim=1).reshape(2 * listed, *a.shape[1:]).contiguous()
    target = torch.tensor([float(i % 2 == 0) for i in range(listed)]).repeat_interleave(2)
    return f1, f2, target


def main():
    os.makedirs(OUT, exist_ok=True)
    for tag, N, K, listed in (("n32_k10", 32, 10, 4), ("n64_k20", 64, 20, 8)):
        torch.manual_seed(0)
        trainer = ref_shim.reference_trainer(K=K, node_num=N)
        model = trainer.model
        module = ref_shim.reference_module(trainer)
        lr, wd = float(trainer.args.learning_rate), float(trainer.args.weight_decay)
        optim = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)          # sg_net.py:351-352
        model.train()
        f1, f2, target = train_batch(listed, N, K, seed=77)
        out = {"features_1": f1.numpy(), "features_2": f2.numpy(), "target": target.numpy(),
               "K": np.int64(K), "N": np.int64(N), "lr": np.float64(lr), "weight_decay": np.float64(wd)}
        for step in (1, 2):
            optim.zero_grad()
            pred, _, _ = model({"features_1": f1, "features_2": f2, "target": target})
            loss = torch.mean(torch.nn.functional.binary_cross_entropy(pred, target))
            loss.backward()
            if step == 1:
                for name, p in module.named_parameters():
                    out["grad1." + name] = p.grad.detach().numpy().copy()
            optim.step()
            out[f"pred{step}"] = pred.detach().numpy().copy()
            out[f"loss{step}"] = np.float64(loss.item())
            for name, v in module.state_dict().items():
                out[f"state{step}." + name] = v.detach().numpy().copy()
            print(f"{tag} step {step}: loss {loss.item():.9g}  pred[:4] {pred[:4].detach().numpy()}")
        np.savez_compressed(os.path.join(OUT, f"ref_train_{tag}.npz"), **out)


if __name__ == "__main__":
    main()
