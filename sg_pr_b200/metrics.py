"""Precision-recall / F1max of large score sets on the device.

`eval_batch.py:70-86` and `SGTrainer.score` (sg_net.py:414-418) call sklearn's `precision_recall_curve` on Python
lists; after an all-pairs scan that is 10^6-10^7 scores and becomes the long pole.  `pr_curve` reproduces sklearn's
definition (thresholds = distinct score values in increasing order, precision/recall evaluated at `score >= t`, the
curve stopped at full recall and closed with the (precision=1, recall=0) point) with one sort + cumsum in torch, on
whatever device the scores live on.  tests/test_metrics.py compares it with sklearn, ties included.
"""
from __future__ import annotations

from typing import Tuple

import torch


def pr_curve(labels: torch.Tensor, scores: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(precision, recall, thresholds) exactly as sklearn.metrics.precision_recall_curve(labels, scores) returns them."""
    scores = scores.reshape(-1).to(torch.float64)
    labels = labels.reshape(-1).to(torch.float64)
    order = torch.argsort(scores, descending=True, stable=True)
    s, y = scores[order], labels[order]
    distinct = torch.nonzero(s[1:] != s[:-1]).reshape(-1)
    idx = torch.cat([distinct, torch.tensor([s.numel() - 1], device=s.device)])
    tps = torch.cumsum(y, 0)[idx]
    fps = 1 + idx.to(torch.float64) - tps
    thr = s[idx]
    precision = tps / (tps + fps)
    precision = torch.nan_to_num(precision, nan=0.0)
    recall = tps / tps[-1] if tps[-1] > 0 else torch.ones_like(tps)
    # sklearn (>= 1.1) keeps every threshold; reverse so recall decreases, then append the (1, 0) end point
    rev = torch.arange(idx.numel() - 1, -1, -1, device=s.device)
    one = torch.ones(1, dtype=torch.float64, device=s.device)
    return torch.cat([precision[rev], one]), torch.cat([recall[rev], 0 * one]), thr[rev]


def roc_curve(labels: torch.Tensor, scores: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(fpr, tpr, thresholds) as sklearn.metrics.roc_curve(labels, scores, drop_intermediate=False) returns them: one point
    per distinct score in decreasing order, preceded by the (0, 0) point with threshold +inf."""
    scores = scores.reshape(-1).to(torch.float64)
    labels = labels.reshape(-1).to(torch.float64)
    order = torch.argsort(scores, descending=True, stable=True)
    s, y = scores[order], labels[order]
    distinct = torch.nonzero(s[1:] != s[:-1]).reshape(-1)
    idx = torch.cat([distinct, torch.tensor([s.numel() - 1], device=s.device)])
    tps = torch.cumsum(y, 0)[idx]
    fps = 1 + idx.to(torch.float64) - tps
    zero = torch.zeros(1, dtype=torch.float64, device=s.device)
    tps, fps = torch.cat([zero, tps]), torch.cat([zero, fps])
    thr = torch.cat([torch.full((1,), float("inf"), dtype=torch.float64, device=s.device), s[idx]])
    fpr = fps / fps[-1] if fps[-1] > 0 else torch.full_like(fps, float("nan"))
    tpr = tps / tps[-1] if tps[-1] > 0 else torch.full_like(tps, float("nan"))
    return fpr, tpr, thr


def auc(x: torch.Tensor, y: torch.Tensor) -> float:
    """Trapezoidal area under a curve given by increasing x (sklearn.metrics.auc)."""
    return float(torch.trapezoid(y.to(torch.float64), x.to(torch.float64)))


def f1_max(labels: torch.Tensor, scores: torch.Tensor) -> float:
    """max over the PR curve of 2PR/(P+R), NaN -> 0 — the number eval_batch.py writes to <seq>_DL_F1_max.txt."""
    p, r, _ = pr_curve(labels, scores)
    f1 = torch.nan_to_num(2 * p * r / (p + r), nan=0.0)
    return float(f1.max())
