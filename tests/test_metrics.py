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


@pytest.mark.parametrize("seed,n", [(5, 2000), (6, 300)])
def test_roc_curve_matches_sklearn(seed, n):
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.4).astype(np.float64)
    y[:2] = (1.0, 0.0)
    s = np.round(rng.random(n), 2)
    f0, t0, th0 = skm.roc_curve(y, s, drop_intermediate=False)
    f1, t1, th1 = metrics.roc_curve(torch.from_numpy(y), torch.from_numpy(s))
    np.testing.assert_allclose(f1.numpy(), f0, rtol=1e-12)
    np.testing.assert_allclose(t1.numpy(), t0, rtol=1e-12)
    np.testing.assert_allclose(th1.numpy()[1:], th0[1:], rtol=0, atol=0)
    assert abs(metrics.auc(f1, t1) - skm.auc(f0, t0)) <= 1e-12
    # eval_batch.py:50-51 uses the default drop_intermediate=True: fewer points, the same area
    fd, td, _ = skm.roc_curve(y, s)
    assert abs(metrics.auc(f1, t1) - skm.auc(fd, td)) <= 1e-12


@pytest.mark.gpu
def test_device_metrics_at_scan_scale_match_sklearn():
    """SURVEY §8 f4 on the device: 4 M scores with heavy ties (fp32 sigmoid outputs saturate at 0 / 1 after an all-pairs
    scan), PR curve / F1max / ROC AUC computed on the GPU vs sklearn on the host."""
    g = torch.Generator().manual_seed(3)
    n = 4_000_000
    y = (torch.rand(n, generator=g) < 0.02).double()
    s = torch.sigmoid(torch.randn(n, generator=g) * 6 + (y * 4 - 2)).float()
    s[::7] = 0.0
    s[::11] = 1.0
    p0, r0, t0 = skm.precision_recall_curve(y.numpy(), s.numpy())
    p1, r1, t1 = metrics.pr_curve(y.cuda(), s.cuda())
    assert p1.is_cuda and p1.shape[0] == p0.shape[0]
    np.testing.assert_allclose(p1.cpu().numpy(), p0, rtol=1e-12)
    np.testing.assert_allclose(r1.cpu().numpy(), r0, rtol=1e-12)
    np.testing.assert_array_equal(t1.cpu().numpy(), t0.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.max(np.nan_to_num(2 * p0 * r0 / (p0 + r0)))
    assert abs(metrics.f1_max(y.cuda(), s.cuda()) - want) <= 1e-12
    f0, tp0, _ = skm.roc_curve(y.numpy(), s.numpy())
    f1, tp1, _ = metrics.roc_curve(y.cuda(), s.cuda())
    assert abs(metrics.auc(f1, tp1) - skm.auc(f0, tp0)) <= 1e-10
