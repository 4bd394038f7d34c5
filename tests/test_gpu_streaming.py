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


def test_p2p_allreduce_protocol_two_ranks_on_one_gpu():
    """bb_allreduce_sum_p2p with both "ranks" on one device (two buffers, two flag arrays, two
    streams): the flag protocol, the double-buffered slots across several epochs, the rank-ordered
    sum and the fused expected log-likelihood.  Real NVLink peers are exercised by bench.py --gpus N."""
    import ctypes
    import torch
    from bayesic_b200.backend import library as L
    from bayesic_b200.parallel import PackedStats
    lib = L.load()
    d, world = 16, 2
    layout = PackedStats.gaussian(d)
    count, stride = layout.numel, (layout.numel + 31) // 32 * 32
    dev = torch.device('cuda')
    bufs = [torch.zeros(2 * stride, dtype=torch.float64, device=dev) for _ in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    buf_ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    flag_ptrs = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
    outs = [torch.zeros(count, dtype=torch.float64, device=dev) for _ in range(world)]
    status = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
    elbo = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.RandomState(0)
    e_lambda = torch.from_numpy(np.eye(d) + 0.1 * rng.randn(d, d)).to(dev)
    e_lambda_mu = torch.from_numpy(rng.randn(d)).to(dev)
    torch.cuda.synchronize()
    for epoch in range(1, 6):
        parts = [rng.randn(count) for _ in range(world)]
        for r in range(world):
            parts[r][-1] = 1000.0 + r                  # row counts
            lo = (epoch & 1) * stride
            bufs[r][lo:lo + count].copy_(torch.from_numpy(parts[r]))
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                L.check(lib.bb_allreduce_sum_p2p(buf_ptrs.data_ptr(), flag_ptrs.data_ptr(), r, world, count, stride,
                                                 epoch, 2000.0, outs[r].data_ptr(), status[r].data_ptr(),
                                                 e_lambda.data_ptr(), e_lambda_mu.data_ptr(), 0.7, -1.3, d,
                                                 elbo[r].data_ptr(), ctypes.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        want = parts[0] + parts[1]
        s2, s1, n = want[:d * d].reshape(d, d), want[d * d:d * d + d], want[-1]
        want_elbo = O.gaussian_expected_loglik(n, s1, s2, e_lambda.cpu().numpy(), e_lambda_mu.cpu().numpy(), 0.7, -1.3)
        for r in range(world):
            assert int(status[r].item()) == 0
            np.testing.assert_array_equal(outs[r].cpu().numpy(), want)          # rank-ordered: bit-identical
            np.testing.assert_allclose(float(elbo[r].item()), want_elbo, rtol=1e-12)
    # a peer that never arrives is reported, not waited for forever
    with torch.cuda.stream(streams[0]):
        L.check(lib.bb_allreduce_sum_p2p(buf_ptrs.data_ptr(), flag_ptrs.data_ptr(), 0, world, count, stride, 6, 5.0,
                                         outs[0].data_ptr(), status[0].data_ptr(), None, None, 0.0, 0.0, 0, None,
                                         ctypes.c_void_p(streams[0].cuda_stream)))
    torch.cuda.synchronize()
    assert int(status[0].item()) == 2                                           # 1 + rank of the missing peer
    with pytest.raises(ValueError):
        L.check(lib.bb_allreduce_sum_p2p(buf_ptrs.data_ptr(), flag_ptrs.data_ptr(), 3, world, count, stride, 7, 5.0,
                                         outs[0].data_ptr(), status[0].data_ptr(), None, None, 0.0, 0.0, 0, None, None))
