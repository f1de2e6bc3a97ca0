"""Device PR-curve / F1max against sklearn (the reference's metric, eval_batch.py:70-86, sg_net.py:414-418)."""
import numpy as np
import pytest
import torch
from sklearn import metrics as skm

from sg_pr_b200 import metrics


@pytest.mark.parametrize("seed,n,ties", [(0, 1000, False), (1, 5000, True), (2, 17, True), (3, 200, False)])
def test_pr_curve_matches_sklearn(seed, n, ties):
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.3).astype(np.float64)
    y[0] = 1.0
    s = rng.random(n)
    if ties:
        s = np.round(s, 2)
    p0, r0, t0 = skm.precision_recall_curve(y, s)
    p1, r1, t1 = metrics.pr_curve(torch.from_numpy(y), torch.from_numpy(s))
    np.testing.assert_allclose(p1.numpy(), p0, rtol=1e-12, atol=0)
    np.testing.assert_allclose(r1.numpy(), r0, rtol=1e-12, atol=0)
    np.testing.assert_allclose(t1.numpy(), t0, rtol=0, atol=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.max(np.nan_to_num(2 * p0 * r0 / (p0 + r0)))
    assert abs(metrics.f1_max(torch.from_numpy(y), torch.from_numpy(s)) - want) <= 1e-12
