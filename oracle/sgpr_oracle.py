"""CPU oracle for the SG_PR pairwise graph-similarity forward.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module — and only as the checker / the timed CPU baseline.  The product (`sg_pr_b200`) never
imports it and has no CPU path of its own.

What it is: a functional fp32 restatement, in stock PyTorch CPU ops, of the reference's eval-mode hot path
`SG.forward` (/root/reference/sg_net.py:112-138), written in the *reference's* arithmetic form — the
`[B, 2C, N, k]` edge tensor is materialised and pushed through a 1x1 convolution exactly as the reference does
— so that its rounding behaviour is the reference's (same ATen kernels: MKL sgemm, oneDNN conv, CPU topk).
It works on a plain `{name: tensor}` state dict (checkpoint keys with the `module.` prefix stripped,
sg_net.py:168-174) and exposes every intermediate the parity tests compare (k-NN distances and index sets
per EdgeConv layer, per-layer features, node embeddings, attention scores, pooled vectors, NTN vector).

Pinning: the reference ships no tests and no expected outputs, so this oracle is pinned
against outputs of the reference ITSELF, produced in the build container by `oracle/make_golden.py`
(which imports /root/reference through `oracle/ref_shim.py`) and committed under `tests/golden/`;
`tests/test_oracle_golden.py` replays them on every run.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d / BatchNorm1d default used at sg_net.py:52-76
LRELU_SLOPE = 0.2      # nn.LeakyReLU(negative_slope=0.2), sg_net.py:53-76

XYZ_LAYERS = ("dgcnn_s_conv1", "dgcnn_s_conv2", "dgcnn_s_conv3")   # sg_net.py:50,58,66
SEM_LAYERS = ("dgcnn_f_conv1", "dgcnn_f_conv2", "dgcnn_f_conv3")   # sg_net.py:54,62,70


def strip_module_prefix(state_dict) -> Dict[str, torch.Tensor]:
    """sg_net.py:168-172 — checkpoints are DataParallel state dicts; drop the 7-char 'module.' prefix."""
    out = {}
    for name, value in state_dict.items():
        out[name[7:] if name.startswith("module.") else name] = torch.as_tensor(value)
    return out


def pairwise_neg_sqdist(x: torch.Tensor) -> torch.Tensor:
    """dgcnn.py:15-17 — `pd[b,i,j] = -xx[j] - inner[i,j] - xx[i]`, inner = -2 x_i.x_j;  x is [B, C, N]."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def knn_indices(x: torch.Tensor, k: int):
    """dgcnn.py:14-20 — k largest of pd per row (nearest, self included).  Returns (idx [B,N,k], pd [B,N,N])."""
    pd = pairwise_neg_sqdist(x)
    return pd.topk(k=k, dim=-1)[1], pd


def edge_tensor(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """dgcnn.py:34-47 — gather neighbours, build cat(nbr - ctr, ctr) → [B, 2C, N, k] (a permuted view)."""
    b, c, n = x.shape
    k = idx.shape[-1]
    rows = x.transpose(2, 1).contiguous().view(b * n, c)
    flat = (idx + torch.arange(b).view(-1, 1, 1) * n).view(-1)
    nbr = rows[flat, :].view(b, n, k, c)
    ctr = rows.view(b, n, 1, c).repeat(1, 1, k, 1)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


def bn_eval(y: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """nn.BatchNorm{1,2}d in eval mode (running statistics), sg_net.py:52 etc."""
    return F.batch_norm(y, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, BN_EPS)


def edgeconv(x: torch.Tensor, k: int, sd, layer: str, trace: dict | None = None) -> torch.Tensor:
    """One dynamic EdgeConv layer: sg_net.py:84-86 (= get_graph_feature → conv1x1+BN+LReLU → max over k)."""
    idx, pd = knn_indices(x, k)
    e = edge_tensor(x, idx)
    y = F.conv2d(e, sd[layer + ".0.weight"])
    y = F.leaky_relu(bn_eval(y, sd, layer + ".1"), LRELU_SLOPE)
    out = y.max(dim=-1, keepdim=False)[0]
    if trace is not None:
        trace.setdefault("knn_idx", []).append(idx)
        trace.setdefault("knn_pd", []).append(pd)
        trace.setdefault("layer_in", []).append(x)
        trace.setdefault("layer_out", []).append(out)
    return out


def node_embeddings(feat: torch.Tensor, k: int, sd, trace: dict | None = None) -> torch.Tensor:
    """sg_net.py:79-110 `dgcnn_conv_pass`: [B, 3+L, N] → [B, N, filters_3]."""
    xyz = feat[:, :3, :]
    sem = feat[:, 3:, :]
    for layer in XYZ_LAYERS:
        xyz = edgeconv(xyz, k, sd, layer, trace)
    for layer in SEM_LAYERS:
        sem = edgeconv(sem, k, sd, layer, trace)
    x = torch.cat((xyz, sem), dim=1)
    x = F.conv1d(x, sd["dgcnn_conv_end.0.weight"])
    x = F.leaky_relu(bn_eval(x, sd, "dgcnn_conv_end.1"), LRELU_SLOPE)
    return x.permute(0, 2, 1)


def attention_pool(emb: torch.Tensor, sd):
    """layers_batch.py:28-39 — returns (pooled [B, F, 1], sigmoid scores [B, N, 1])."""
    b = emb.shape[0]
    w = sd["attention.weight_matrix"]
    ctx = torch.tanh(torch.mean(torch.matmul(emb, w), dim=1))
    att = torch.sigmoid(torch.matmul(emb, ctx.view(b, -1, 1)))
    pooled = torch.matmul(emb.permute(0, 2, 1), att)
    return pooled, att


def ntn_vector(e1: torch.Tensor, e2: torch.Tensor, sd) -> torch.Tensor:
    """layers_batch.py:70-83 — relu(e1' W e2 + V [e1; e2] + b) → [B, T, 1]."""
    b, f, _ = e1.shape
    w = sd["tensor_network.weight_matrix"]
    t = w.shape[2]
    s = torch.matmul(e1.permute(0, 2, 1), w.view(f, -1)).view(b, f, t)
    s = torch.matmul(s.permute(0, 2, 1), e2)
    blk = torch.matmul(sd["tensor_network.weight_matrix_block"], torch.cat((e1, e2), dim=1))
    return F.relu(s + blk + sd["tensor_network.bias"])


def score_head(ntn: torch.Tensor, sd) -> torch.Tensor:
    """sg_net.py:131-136 — sigmoid(Linear(relu(Linear(ntn')))) → [B]."""
    h = F.relu(F.linear(ntn.permute(0, 2, 1), sd["fully_connected_first.weight"], sd["fully_connected_first.bias"]))
    return torch.sigmoid(F.linear(h, sd["scoring_layer.weight"], sd["scoring_layer.bias"])).reshape(-1)


@torch.no_grad()
def embed_graphs(feat: torch.Tensor, k: int, sd, want_trace: bool = False) -> dict:
    """Per-graph half of the path (sg_net.py:123 + :126): node embeddings → attention pooling."""
    trace = {} if want_trace else None
    emb = node_embeddings(feat, k, sd, trace)
    pooled, att = attention_pool(emb, sd)
    out = {"emb": emb, "pooled": pooled, "att": att}
    if trace is not None:
        out.update(trace)
    return out


@torch.no_grad()
def forward_pairs(f1: torch.Tensor, f2: torch.Tensor, k: int, sd, want_trace: bool = False) -> dict:
    """sg_net.py:112-138 `SG.forward` in eval mode. Returns score [B], att_1/att_2 [B, N, 1] + intermediates."""
    g1 = embed_graphs(f1, k, sd, want_trace)
    g2 = embed_graphs(f2, k, sd, want_trace)
    ntn = ntn_vector(g1["pooled"], g2["pooled"], sd)
    out = {"score": score_head(ntn, sd), "att_1": g1["att"], "att_2": g2["att"], "ntn": ntn,
           "pooled_1": g1["pooled"], "pooled_2": g2["pooled"], "emb_1": g1["emb"], "emb_2": g2["emb"]}
    if want_trace:
        for side, g in (("1", g1), ("2", g2)):
            for key in ("knn_idx", "knn_pd", "layer_in", "layer_out"):
                out[f"{key}_{side}"] = g[key]
    return out


@torch.no_grad()
def score_matrix(pooled_rows: torch.Tensor, pooled_cols: torch.Tensor, sd) -> torch.Tensor:
    """All ordered pairs (row graph = side 1, column graph = side 2) through layers_batch.py:70-83 and
    sg_net.py:131-136.  pooled_* are [R, F] / [M, F]; returns [R, M]."""
    r, f = pooled_rows.shape
    m = pooled_cols.shape[0]
    e1 = pooled_rows.view(r, 1, f, 1).expand(r, m, f, 1).reshape(r * m, f, 1)
    e2 = pooled_cols.view(1, m, f, 1).expand(r, m, f, 1).reshape(r * m, f, 1)
    return score_head(ntn_vector(e1, e2, sd), sd).view(r, m)


def knn_sets_equivalent(pd_ref: torch.Tensor, idx_ref: torch.Tensor, idx_test: torch.Tensor,
                        x_in: torch.Tensor | None = None) -> torch.Tensor:
    """Per-row verdict that two k-NN index sets are the same selection up to exact ties (SURVEY §7 hard parts 1-2).
    Two rows agree when either
      (a) they pick the same multiset of reference distances, or
      (b) (needs the layer input `x_in` [B, C, N]) they pick the same multiset of NODE VECTORS — indices are mapped
          to the first node with a bit-identical feature vector; zero pads and same-label nodes are such
          duplicates, and the reference itself chooses among them by sgemm rounding noise.
    Returns bool [B, N]."""
    a = torch.gather(pd_ref, -1, idx_ref.long()).sort(dim=-1)[0]
    b = torch.gather(pd_ref, -1, idx_test.long()).sort(dim=-1)[0]
    same = (a == b).all(dim=-1)
    if x_in is not None:
        rows = x_in.transpose(2, 1)
        dup = (rows[:, :, None, :] == rows[:, None, :, :]).all(dim=-1)
        canon = dup.to(torch.uint8).argmax(dim=-1)                       # first identical node
        ca = torch.gather(canon[:, None, :].expand(-1, idx_ref.shape[1], -1), -1, idx_ref.long()).sort(dim=-1)[0]
        cb = torch.gather(canon[:, None, :].expand(-1, idx_test.shape[1], -1), -1, idx_test.long()).sort(dim=-1)[0]
        same = same | (ca == cb).all(dim=-1)
    return same


NEAR_TIE = 1e-6   # see classify_knn_rows


def classify_knn_rows(pd_ref: torch.Tensor, idx_ref: torch.Tensor, idx_test: torch.Tensor, x_in: torch.Tensor,
                      near: float = NEAR_TIE) -> torch.Tensor:
    """Per-row verdict [B, N] on a k-NN selection against the reference's (SURVEY §7 hard parts 1-2):
         0  same nodes — identical sets up to bit-identical duplicates (zero pads, same-label nodes);
         1  exact-tie swap — different nodes whose REFERENCE distances are equal in fp32 (topk's choice is arbitrary);
         2  near tie — the two selections, sorted by reference distance, differ element-wise by at most
            `near` * (xx_i + max_j xx_j), i.e. the swapped columns sit inside the rounding noise of
            pd = 2 x_i.x_j - xx_j - xx_i  (dgcnn.py:15-17; observed swaps: <= 1.5e-7 of that scale);
         3  mismatch — anything else (a real disagreement, or the downstream effect of an earlier 1/2 in the same branch).
    Verdicts 1 and 2 change the score (by up to ~1e-2) although neither side is wrong: the reference's own CPU and GPU
    paths, or fp32 and fp64, disagree on exactly these rows."""
    rows = x_in.transpose(2, 1)
    dup = (rows[:, :, None, :] == rows[:, None, :, :]).all(dim=-1)
    canon = dup.to(torch.uint8).argmax(dim=-1)
    ca = torch.gather(canon[:, None, :].expand(-1, idx_ref.shape[1], -1), -1, idx_ref.long()).sort(dim=-1)[0]
    cb = torch.gather(canon[:, None, :].expand(-1, idx_test.shape[1], -1), -1, idx_test.long()).sort(dim=-1)[0]
    same = (ca == cb).all(dim=-1)
    ref_vals = torch.gather(pd_ref, -1, idx_ref.long()).sort(dim=-1)[0]
    test_vals = torch.gather(pd_ref, -1, idx_test.long()).sort(dim=-1)[0]
    depth = (ref_vals - test_vals).abs().max(dim=-1)[0]
    same_pd = (ref_vals == test_vals).all(dim=-1)
    xx = (x_in * x_in).sum(dim=1)
    scale = xx + xx.max(dim=-1, keepdim=True)[0]
    code = torch.full(same.shape, 3, dtype=torch.int64)
    code[depth <= near * scale] = 2
    code[same_pd] = 1
    code[same] = 0
    return code


def load_state_npz(path: str) -> Dict[str, torch.Tensor]:
    """Read a checkpoint stored as .npz (tests/golden/*.npz; written by oracle/make_golden.py)."""
    import numpy as np
    with np.load(path) as z:
        return {name: torch.from_numpy(z[name].copy()) for name in z.files}
