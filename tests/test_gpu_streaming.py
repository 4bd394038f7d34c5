"""Minibatch streaming and on-device minibatch selection (bayesic_b200/streaming.py):
a pass over host-resident data in double-buffered chunks gives the same statistics as the resident
pass (both within tolerance of the float64 oracle), including ragged last chunks and empty input;
``gather_rows`` is exact (a copy)."""
import numpy as np
import pytest

from bayesic_b200 import stats, streaming
from oracle import closed_forms as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,d,chunk,pinned', [(50000, 256, 8192, False), (20001, 64, 4096, True), (300, 256, 1000, False)])
def test_streamed_regression_statistics(n, d, chunk, pinned):
    import torch
    rng = np.random.RandomState(n)
    X = rng.randn(n, d).astype(np.float32)
    y = rng.randn(n).astype(np.float32)
    hx, hy = torch.from_numpy(X), torch.from_numpy(y)
    if pinned:
        hx, hy = hx.pin_memory(), hy.pin_memory()
    xtx, xty, yty = streaming.streamed_pass(stats.regression_suffstats, (hx, hy), chunk)
    want = O.regression_suffstats(X, y)
    scale = np.abs(want[0]).max()
    np.testing.assert_allclose(xtx.cpu().numpy(), want[0], rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(xty.cpu().numpy(), want[1], rtol=1e-4, atol=2e-5 * np.sqrt(scale * want[2]))
    np.testing.assert_allclose(float(yty.reshape(-1)[0]), want[2], rtol=1e-5)
    rx, ry, ryy = stats.regression_suffstats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    np.testing.assert_allclose(xtx.cpu().numpy(), rx.double().cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)


def test_streamed_logistic_pass_and_single_output():
    import torch
    rng = np.random.RandomState(5)
    n, d, s = 30000, 256, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    W = torch.from_numpy((rng.randn(s, d) / np.sqrt(d)).astype(np.float32)).cuda()
    ll, G = streaming.streamed_pass(lambda xc, yc: stats.logistic_reparam_stats(xc, yc, W), (X, y), 7000)
    rl, rG = stats.logistic_reparam_stats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), W)
    np.testing.assert_allclose(ll.cpu().numpy(), rl.cpu().numpy(), rtol=1e-5)
    scale = float(rG.abs().max())
    np.testing.assert_allclose(G.cpu().numpy(), rG.cpu().numpy(), rtol=1e-4, atol=1e-4 * scale)
    # a pass with one output comes back as one tensor
    s2 = streaming.streamed_pass(lambda xc: stats.gaussian_suffstats(xc)[2], (X[:, :64].copy(),), 9999)
    np.testing.assert_allclose(s2.cpu().numpy(), X[:, :64].astype(np.float64).T @ X[:, :64].astype(np.float64),
                               rtol=1e-4, atol=1e-4 * n)


def test_streamed_pass_edge_cases():
    import torch
    out = streaming.streamed_pass(stats.regression_suffstats, (np.zeros((0, 32), np.float32), np.zeros(0, np.float32)), 128)
    assert all(float(o.abs().sum()) == 0.0 for o in out) and tuple(out[0].shape) == (32, 32)
    with pytest.raises(ValueError):
        streaming.streamed_pass(stats.regression_suffstats, (np.zeros((4, 8), np.float32), np.zeros(5, np.float32)), 2)
    with pytest.raises(TypeError):
        streaming.streamed_pass(stats.regression_suffstats, (np.zeros((4, 8)), np.zeros(4)), 2)
    with pytest.raises(ValueError):
        streaming.streamed_pass(stats.regression_suffstats, (np.zeros((4, 8), np.float32), np.zeros(4, np.float32)), 0)
    del torch


@pytest.mark.parametrize('d', [1, 7, 64, 1024])
def test_gather_rows_is_exact_and_flags_bad_indices(d):
    import torch
    g = torch.Generator(device='cuda').manual_seed(d)
    X = torch.randn(5000, d, device='cuda', generator=g)
    idx = torch.randint(0, 5000, (777,), device='cuda', generator=g)
    got = streaming.gather_rows(X, idx)
    assert torch.equal(got, X[idx])
    assert streaming.gather_rows(X, idx[:0]).shape == (0, d)
    bad = idx.clone()
    bad[3], bad[10] = 5000, -1
    with pytest.raises(IndexError):
        streaming.gather_rows(X, bad)
    rows, count = streaming.gather_rows(X, bad, check=False)
    assert int(count.item()) == 2 and bool(torch.isnan(rows[3]).all()) and torch.equal(rows[4], X[idx[4]])
    with pytest.raises(TypeError):
        streaming.gather_rows(X, idx.int())
