"""CPU oracle for one SG_PR TRAINING step.  TEST INFRASTRUCTURE ONLY (same import rules as sgpr_oracle.py).

Restates, in functional stock-PyTorch CPU ops with autograd, what `SGTrainer.process_batch(batch, training=True)`
does once the batch tensors exist (/root/reference/sg_net.py:332-338):

    model.train(); optimizer.zero_grad(); prediction = model(data)            sg_net.py:112-138 in train mode
    loss = mean(binary_cross_entropy(prediction, target))                      sg_net.py:335
    loss.backward(); Adam(lr, weight_decay).step()                             sg_net.py:337-338, 351-352

Train mode changes exactly one thing in the forward: the seven BatchNorm layers (sg_net.py:52-76) normalise with the
statistics of the current batch — per channel over B*N*k edge activations (BatchNorm2d) or B*N node activations
(BatchNorm1d), biased variance — and fold them into the running statistics with momentum 0.1 (unbiased variance).
`dgcnn_conv_pass` is called once per side (sg_net.py:123-124), so each side is its own BatchNorm batch and the running
statistics are updated twice per step, side 1 first.  k-NN indices carry no gradient (`topk` indices, dgcnn.py:19).

Pinning: `oracle/make_golden_train.py` runs the unmodified reference for two steps and commits inputs, predictions,
losses, step-1 gradients and the full state after each step under tests/golden/ref_train_*.npz;
tests/test_oracle_golden.py::test_train_step_matches_reference replays them.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import sgpr_oracle as orc

BN_MOMENTUM = 0.1      # nn.BatchNorm default, sg_net.py:52
BN_LAYERS = tuple(l + ".1" for l in orc.XYZ_LAYERS + orc.SEM_LAYERS) + ("dgcnn_conv_end.1",)


def is_param(name: str) -> bool:
    return not (name.endswith("running_mean") or name.endswith("running_var") or name.endswith("num_batches_tracked"))


def _bn_train(y: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """Batch-statistics BatchNorm; updates sd[prefix.running_*] and num_batches_tracked in place like nn.BatchNorm."""
    out = F.batch_norm(y, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                       sd[prefix + ".bias"], True, BN_MOMENTUM, orc.BN_EPS)
    if prefix + ".num_batches_tracked" in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def _edgeconv_train(x, k, sd, layer, trace=None):
    idx, pd = orc.knn_indices(x.detach(), k)
    e = orc.edge_tensor(x, idx)
    y = F.conv2d(e, sd[layer + ".0.weight"])
    if trace is not None:
        trace.setdefault("knn_idx", []).append(idx)
        trace.setdefault("knn_pd", []).append(pd)
        trace.setdefault("layer_in", []).append(x.detach())
        trace.setdefault("bn_mean", []).append(y.detach().mean(dim=(0, 2, 3)))
        trace.setdefault("bn_var", []).append(y.detach().var(dim=(0, 2, 3), unbiased=False))
    out = F.leaky_relu(_bn_train(y, sd, layer + ".1"), orc.LRELU_SLOPE).max(dim=-1)[0]
    if trace is not None:
        trace.setdefault("layer_out", []).append(out.detach())
    return out


def node_embeddings_train(feat, k, sd, trace=None):
    """sg_net.py:79-110 in train mode: [B, 3+L, N] -> [B, N, filters_3]."""
    xyz, sem = feat[:, :3, :], feat[:, 3:, :]
    for layer in orc.XYZ_LAYERS:
        xyz = _edgeconv_train(xyz, k, sd, layer, trace)
    for layer in orc.SEM_LAYERS:
        sem = _edgeconv_train(sem, k, sd, layer, trace)
    x = F.conv1d(torch.cat((xyz, sem), dim=1), sd["dgcnn_conv_end.0.weight"])
    if trace is not None:
        trace["end_mean"] = x.detach().mean(dim=(0, 2))
        trace["end_var"] = x.detach().var(dim=(0, 2), unbiased=False)
    x = F.leaky_relu(_bn_train(x, sd, "dgcnn_conv_end.1"), orc.LRELU_SLOPE)
    return x.permute(0, 2, 1)


def forward_train(sd, f1, f2, k, want_trace: bool = False):
    """`SG.forward` in train mode on a state dict whose BN buffers are updated in place.  Returns (score, aux)."""
    t1, t2 = ({} if want_trace else None), ({} if want_trace else None)
    e1 = node_embeddings_train(f1, k, sd, t1)
    e2 = node_embeddings_train(f2, k, sd, t2)
    p1, a1 = orc.attention_pool(e1, sd)
    p2, a2 = orc.attention_pool(e2, sd)
    ntn = orc.ntn_vector(p1, p2, sd)
    score = orc.score_head(ntn, sd)
    return score, {"emb_1": e1, "emb_2": e2, "pooled_1": p1, "pooled_2": p2, "att_1": a1, "att_2": a2, "ntn": ntn,
                   "trace_1": t1, "trace_2": t2}


def new_adam_state(sd) -> dict:
    return {"step": 0, "m": {n: torch.zeros_like(v) for n, v in sd.items() if is_param(n)},
            "v": {n: torch.zeros_like(v) for n, v in sd.items() if is_param(n)}}


def train_step(sd: Dict[str, torch.Tensor], f1, f2, target, k: int, adam: dict, lr: float, weight_decay: float,
               betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, want_trace: bool = False):
    """One optimiser step, in place on `sd` (parameters and BatchNorm buffers) and `adam`.
    Returns {"loss", "pred", "grads": {name: tensor}, "aux"}."""
    work = {}
    for name, value in sd.items():
        work[name] = value.detach().clone().requires_grad_(True) if is_param(name) else value
    pred, aux = forward_train(work, f1, f2, k, want_trace)
    loss = torch.mean(F.binary_cross_entropy(pred, target))
    names = [n for n in sd if is_param(n)]
    grads = dict(zip(names, torch.autograd.grad(loss, [work[n] for n in names])))
    # torch.optim.Adam (sg_net.py:351-352): L2 weight decay added to the gradient, bias-corrected moments
    adam["step"] += 1
    t = adam["step"]
    b1, b2 = betas
    with torch.no_grad():
        for n in names:
            g = grads[n] + weight_decay * sd[n]
            adam["m"][n].mul_(b1).add_(g, alpha=1 - b1)
            adam["v"][n].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (adam["v"][n].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            sd[n].addcdiv_(adam["m"][n], denom, value=-lr / (1 - b1 ** t))
    return {"loss": float(loss.detach()), "pred": pred.detach(), "grads": grads, "aux": aux}
