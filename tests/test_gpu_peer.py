"""The library's own multi-GPU combine, exercised with both "ranks" on ONE device (two sets of
buffers, two handles, two streams; the kernels of the two ranks run concurrently and talk through
the same flag protocol they use over NVLink): ``bb_comm_*`` (two-shot all-reduce of a packed float64
payload, csrc/p2p_reduce.cu) and ``bb_gaussian_pass_*`` (statistics + exchange + expected
log-likelihood in ONE launch, csrc/suffstats_sm100.cu).  Real NVLink peers are exercised by
``bench.py --gpus N`` and ``tests/multi_gpu_check.py`` under torchrun.

Reference for what is combined: the iid-summed statistics of bayesic/distribution/base.py:328-332."""
import ctypes

import numpy as np
import pytest

from bayesic_b200.backend import library as L
from oracle import closed_forms as O

pytestmark = pytest.mark.gpu


def _ptrs(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _make_comms(world, capacity, spin_ms=2000.0):
    import torch
    lib = L.load()
    dev = torch.device('cuda')
    ins = [torch.zeros(capacity, dtype=torch.float64, device=dev) for _ in range(world)]
    outs = [torch.zeros(capacity, dtype=torch.float64, device=dev) for _ in range(world)]
    flags = [torch.zeros(lib.bb_comm_flag_bytes(world) // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    comms = []
    for r in range(world):
        h = ctypes.c_void_p()
        L.check(lib.bb_comm_create(r, world, _ptrs(ins), _ptrs(outs), _ptrs(flags), capacity, spin_ms, ctypes.byref(h)))
        comms.append(h)
    return lib, ins, outs, flags, comms


@pytest.mark.parametrize('world,count', [(2, 1), (2, 33), (2, 4161), (3, 4161), (2, 300001), (4, 1060000)])
def test_comm_allreduce_ranks_on_one_gpu(world, count):
    import torch
    lib, ins, outs, flags, comms = _make_comms(world, count + 5)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.RandomState(count)
    try:
        for epoch in range(3):
            parts = [rng.randn(count) for _ in range(world)]
            for r in range(world):
                ins[r][:count].copy_(torch.from_numpy(parts[r]))
                outs[r].fill_(-7.0)
            torch.cuda.synchronize()
            for r in range(world):
                L.check(lib.bb_comm_allreduce_sum(comms[r], count, ctypes.c_void_p(streams[r].cuda_stream)))
            torch.cuda.synchronize()
            want = np.zeros(count)
            for r in range(world):
                want = want + parts[r]                      # rank order, like the kernel
            for r in range(world):
                status = ctypes.c_int32(-1)
                L.check(lib.bb_comm_status(comms[r], ctypes.byref(status), None))
                assert status.value == 0
                np.testing.assert_array_equal(outs[r][:count].cpu().numpy(), want)     # bit-identical on every rank
                assert float(outs[r][count]) == -7.0                                   # nothing past count is touched
    finally:
        for h in comms:
            lib.bb_comm_destroy(h)


def test_comm_lost_peer_is_reported_and_poisons_the_output():
    import torch
    lib, ins, outs, flags, comms = _make_comms(2, 5000, spin_ms=5.0)
    try:
        ins[0].fill_(1.0)
        L.check(lib.bb_comm_allreduce_sum(comms[0], 5000, None))        # rank 1 never calls
        status = ctypes.c_int32(0)
        L.check(lib.bb_comm_status(comms[0], ctypes.byref(status), None))
        assert status.value == 2                                        # 1 + rank of the missing peer
        assert bool(torch.isnan(outs[0]).all())                         # never a partial sum
        with pytest.raises(ValueError):
            L.check(lib.bb_comm_allreduce_sum(comms[0], 5001, None))    # beyond capacity
    finally:
        for h in comms:
            lib.bb_comm_destroy(h)


def _make_passes(world, d, spin_ms=2000.0):
    import torch
    lib = L.load()
    dev = torch.device('cuda')
    recv_bytes, flag_bytes = ctypes.c_int64(0), ctypes.c_int64(0)
    L.check(lib.bb_gaussian_pass_peer_bytes(d, world, ctypes.byref(recv_bytes), ctypes.byref(flag_bytes)))
    recv = [torch.zeros(recv_bytes.value // 8, dtype=torch.float64, device=dev) for _ in range(world)]
    flags = [torch.zeros(flag_bytes.value // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    passes = []
    for r in range(world):
        h = ctypes.c_void_p()
        L.check(lib.bb_gaussian_pass_create(d, ctypes.byref(h)))
        if world > 1:
            L.check(lib.bb_gaussian_pass_attach_peers(h, r, world, _ptrs(recv), _ptrs(flags), spin_ms))
        passes.append(h)
    return lib, recv, flags, passes


@pytest.mark.parametrize('world,d,rows', [(1, 64, [20000]), (2, 64, [1500, 2100]), (2, 16, [640, 0]),
                                          (3, 32, [1000, 130, 2500]), (2, 64, [128 * 20, 128 * 20])])
def test_gaussian_pass_one_launch_statistics_exchange_and_loglik(world, d, rows):
    """Every rank's outputs equal the float64 oracle on the concatenated rows and are bit-identical
    across ranks; several steps in a row exercise the epoch-parity double buffering."""
    import torch
    import bayesic_b200.stats as S
    lib, recv, flags, passes = _make_passes(world, d)
    dev = torch.device('cuda')
    rng = np.random.RandomState(d + sum(rows))
    a = rng.randn(d, d)
    e_lambda = torch.from_numpy(a @ a.T / d + np.eye(d)).to(dev)
    e_lambda_mu = torch.from_numpy(rng.randn(d)).to(dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    s1 = [torch.zeros(d, dtype=torch.float64, device=dev) for _ in range(world)]
    s2 = [torch.zeros(d, d, dtype=torch.float64, device=dev) for _ in range(world)]
    cnt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    ll = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    try:
        for step in range(4):
            Xs = [(rng.randn(n, d) * 1.2 + 0.3).astype(np.float32) for n in rows]
            Xd = [torch.from_numpy(x).to(dev) for x in Xs]
            torch.cuda.synchronize()
            before = S.launch_count()
            for r in range(world):
                L.check(lib.bb_gaussian_pass_run(passes[r], Xd[r].data_ptr() if rows[r] else None, rows[r],
                                                 e_lambda.data_ptr(), e_lambda_mu.data_ptr(), 0.7, -1.3,
                                                 float(sum(rows)), s1[r].data_ptr(), s2[r].data_ptr(),
                                                 cnt[r].data_ptr(), ll[r].data_ptr(),
                                                 ctypes.c_void_p(streams[r].cuda_stream)))
            assert S.launch_count() - before == world                   # ONE launch per rank per step
            torch.cuda.synchronize()
            n, w1, w2 = O.gaussian_suffstats(np.concatenate(Xs, axis=0))
            want_ll = O.gaussian_expected_loglik(n, w1, w2, e_lambda.cpu().numpy(), e_lambda_mu.cpu().numpy(), 0.7, -1.3)
            for r in range(world):
                status = ctypes.c_int32(-1)
                L.check(lib.bb_gaussian_pass_status(passes[r], ctypes.byref(status), None))
                assert status.value == 0
                np.testing.assert_allclose(s2[r].cpu().numpy(), w2, rtol=1e-4, atol=1e-6 * np.abs(w2).max())
                np.testing.assert_allclose(s1[r].cpu().numpy(), w1, rtol=1e-4, atol=1e-6 * np.abs(w1).max())
                if world > 1:
                    assert float(cnt[r]) == float(n)
                np.testing.assert_allclose(float(ll[r]), want_ll, rtol=1e-5)
                assert torch.equal(s2[r], s2[0]) and torch.equal(s1[r], s1[0]) and torch.equal(ll[r], ll[0])
    finally:
        for h in passes:
            lib.bb_gaussian_pass_destroy(h)


def test_gaussian_pass_lost_peer_gives_nan_and_a_status():
    import torch
    d = 32
    lib, recv, flags, passes = _make_passes(2, d, spin_ms=5.0)
    dev = torch.device('cuda')
    X = torch.randn(1000, d, device=dev)
    s1 = torch.zeros(d, dtype=torch.float64, device=dev)
    s2 = torch.zeros(d, d, dtype=torch.float64, device=dev)
    try:
        L.check(lib.bb_gaussian_pass_run(passes[0], X.data_ptr(), 1000, None, None, 0.0, 0.0, 1000.0, s1.data_ptr(),
                                         s2.data_ptr(), None, None, None))            # rank 1 never runs
        status = ctypes.c_int32(0)
        L.check(lib.bb_gaussian_pass_status(passes[0], ctypes.byref(status), None))
        assert status.value == 2
        assert bool(torch.isnan(s2).all()) and bool(torch.isnan(s1).all())
    finally:
        for h in passes:
            lib.bb_gaussian_pass_destroy(h)


def test_gaussian_pass_matches_the_workspace_entry_point():
    """bb_suffstats_gaussian_loglik (caller workspace) and the handle-based pass are the same kernel (its CTAs'
    float64 partial sums meet in arrival order, so two launches agree to ~1e-16 relative, not bit for bit)."""
    import torch
    import bayesic_b200.stats as S
    from bayesic_b200.parallel import GaussianPass
    d, n = 64, 70001
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(5)
    X = torch.randn(n, d, device=dev, generator=g) * 1.3 + 0.4
    e_lambda = torch.eye(d, dtype=torch.float64, device=dev) * 1.5
    e_lambda_mu = torch.linspace(-1, 1, d, dtype=torch.float64, device=dev)
    _, a1, a2, all_ = S.gaussian_suffstats_loglik(X, e_lambda, e_lambda_mu, 0.3, -0.2)
    p = GaussianPass(d, dev)
    cnt, b1, b2, bll = p.run(X, e_lambda, e_lambda_mu, 0.3, -0.2)
    p.check()
    close = lambda u, v: torch.testing.assert_close(u, v, rtol=1e-12, atol=1e-12 * float(v.abs().max()))
    close(a1, b1), close(a2, b2), close(all_, bll)
    _, c1, c2 = S.gaussian_suffstats(X)
    close(a2, c2)
    assert torch.equal(b2, b2.T)                      # symmetric by construction


def _reduce_on_one_gpu(layout, per_rank_parts, rows):
    """What ``parallel._reduce_into`` does on every rank -- pack the local parts and the row count into the
    communicator's input, one ``bb_comm_allreduce_sum`` -- with all the "ranks" on this device."""
    import torch
    world = len(per_rank_parts)
    lib, ins, outs, flags, comms = _make_comms(world, layout.numel)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        for r in range(world):
            views = layout.views(ins[r][:layout.numel])
            for name, value in per_rank_parts[r].items():
                views[name].copy_(value.reshape(views[name].shape))
            views['count'].fill_(float(rows[r]))
        torch.cuda.synchronize()
        for r in range(world):
            L.check(lib.bb_comm_allreduce_sum(comms[r], layout.numel, ctypes.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        for r in range(world):
            status = ctypes.c_int32(-1)
            L.check(lib.bb_comm_status(comms[r], ctypes.byref(status), None))
            assert status.value == 0
            assert torch.equal(outs[r][:layout.numel], outs[0][:layout.numel])       # replicated bit for bit
        return {k: v.cpu().numpy() for k, v in layout.views(outs[0][:layout.numel].clone()).items()}
    finally:
        for h in comms:
            lib.bb_comm_destroy(h)


def _scaled_close(got, want, scale, tol):
    np.testing.assert_array_less(np.abs(got - want), tol * scale + 1e-300)


def test_sharded_regression_statistics_equal_the_unsharded_oracle():
    """cfg4 over 3 shards (ragged: 2304 / 1 / 1791 rows): local tcgen05 Gram passes, one peer all-reduce of
    D^2 + D + 2 float64; against the float64 oracle on the concatenated rows."""
    import torch
    import bayesic_b200.stats as S
    from bayesic_b200.parallel import PackedStats
    d, rows = 256, [2304, 1, 1791]
    rng = np.random.RandomState(11)
    X = rng.randn(sum(rows), d).astype(np.float32)
    y = (X @ (rng.randn(d) / np.sqrt(d)) + 0.1 * rng.randn(sum(rows))).astype(np.float32)
    parts, lo = [], 0
    for n in rows:
        xtx, xty, yty = S.regression_suffstats(torch.from_numpy(X[lo:lo + n]).cuda(), torch.from_numpy(y[lo:lo + n]).cuda())
        parts.append({'xtx': xtx, 'xty': xty, 'yty': yty})
        lo += n
    got = _reduce_on_one_gpu(PackedStats.regression(d), parts, rows)
    wxtx, wxty, wyty = O.regression_suffstats(X, y)
    dg = np.sqrt(np.diag(wxtx))
    _scaled_close(got['xtx'], wxtx, np.outer(dg, dg), 2e-5)
    _scaled_close(got['xty'], wxty, dg * np.sqrt(wyty), 2e-5)
    np.testing.assert_allclose(got['yty'], wyty, rtol=1e-6)
    assert float(got['count']) == float(sum(rows))


def test_sharded_mixture_statistics_equal_the_unsharded_oracle():
    """cfg3 over 2 shards: responsibilities stay sharded; {N_k, sum r x, sum r x x^T, sum lse} are summed."""
    import torch
    import bayesic_b200.stats as S
    from bayesic_b200.parallel import PackedStats
    d, k, rows = 64, 256, [3000, 2120]
    rng = np.random.RandomState(12)
    X = rng.randn(sum(rows), d).astype(np.float32)
    logits = (rng.randn(sum(rows), k) * 2.0).astype(np.float32)
    lr64 = logits.astype('f8') - np.log(np.exp(logits.astype('f8')).sum(1, keepdims=True))
    R64 = np.exp(lr64)
    parts, lo = [], 0
    for n in rows:
        lg = torch.from_numpy(logits[lo:lo + n]).cuda()
        log_resp, lse, sum_lse = S.log_responsibilities(lg)
        nk, rx, rxx = S.weighted_suffstats(torch.from_numpy(X[lo:lo + n]).cuda(), torch.exp(log_resp))
        parts.append({'nk': nk, 'rx': rx, 'rxx': rxx, 'sum_lse': sum_lse})
        lo += n
    got = _reduce_on_one_gpu(PackedStats.mixture(k, d), parts, rows)
    X64 = X.astype('f8')
    wnk, wrx = R64.sum(0), R64.T @ X64
    wrxx = np.einsum('nk,nd,ne->kde', R64, X64, X64)
    x2 = R64.T @ (X64 ** 2)
    np.testing.assert_allclose(got['nk'], wnk, rtol=2e-5)
    _scaled_close(got['rx'], wrx, np.sqrt(wnk[:, None] * x2), 2e-5)
    _scaled_close(got['rxx'], wrxx, np.sqrt(np.einsum('kd,ke->kde', x2, x2)), 2e-5)
    want_lse = np.log(np.exp(logits.astype('f8')).sum(1)).sum()
    np.testing.assert_allclose(float(got['sum_lse']), want_lse, rtol=1e-6)
    assert float(got['count']) == float(sum(rows))


def test_sharded_logistic_gradient_equals_the_unsharded_oracle():
    """cfg5 over 2 shards: replicated draws W, {loglik[S], G[D, S]} summed over the ranks' rows."""
    import torch
    import bayesic_b200.stats as S
    from bayesic_b200.parallel import PackedStats
    d, s, rows = 512, 64, [1400, 2700]
    rng = np.random.RandomState(13)
    X = rng.randn(sum(rows), d).astype(np.float32)
    y = (rng.rand(sum(rows)) < 0.5).astype(np.float32)
    W = (rng.randn(s, d) / np.sqrt(d)).astype(np.float32)
    Wd = torch.from_numpy(W).cuda()
    parts, lo = [], 0
    for n in rows:
        loglik, G = S.logistic_reparam_stats(torch.from_numpy(X[lo:lo + n]).cuda(), torch.from_numpy(y[lo:lo + n]).cuda(), Wd)
        parts.append({'loglik': loglik, 'G': G})
        lo += n
    got = _reduce_on_one_gpu(PackedStats.logistic(d, s), parts, rows)
    Z = X.astype('f8') @ W.astype('f8').T
    want_ll = (y[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)
    resid = y[:, None] - 1.0 / (1.0 + np.exp(-Z))
    want_G = X.astype('f8').T @ resid
    np.testing.assert_allclose(got['loglik'], want_ll, rtol=2e-5)
    scale = np.linalg.norm(X.astype('f8'), axis=0)[:, None] * np.linalg.norm(resid, axis=0)[None, :]
    _scaled_close(got['G'], want_G, scale, 2e-5)
    assert float(got['count']) == float(sum(rows))


def test_gaussian_pass_back_to_back_launches_do_not_interfere():
    """Consecutive passes of one handle overlap (programmatic dependent launch: the next launch streams while the
    previous launch's last CTA gathers, re-zeroes the accumulator block and writes the outputs) and share the
    accumulator block, the ticket and two alternating tile counters.  400 launches back to back over three
    different inputs, no synchronisation in between: the outputs after the last launch must be those of its input
    alone, and the shared state must be clean for the launch after it."""
    import torch
    import bayesic_b200.stats as S
    from bayesic_b200.parallel import GaussianPass
    d = 64
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(11)
    Xs = [torch.randn(n, d, device=dev, generator=g) * s + m for n, s, m in ((300000, 1.3, 0.4), (40001, 0.7, -1.0), (128 * 148 * 4, 2.0, 0.1))]
    e_lambda = torch.eye(d, dtype=torch.float64, device=dev) * 1.5
    e_lambda_mu = torch.linspace(-1, 1, d, dtype=torch.float64, device=dev)
    want = []
    for X in Xs:
        _, w1, w2, wll = S.gaussian_suffstats_loglik(X, e_lambda, e_lambda_mu, 0.3, -0.2)
        want.append((w1.clone(), w2.clone(), wll.clone()))
    p = GaussianPass(d, dev)
    close = lambda u, v: torch.testing.assert_close(u, v, rtol=1e-12, atol=1e-12 * float(v.abs().max()))
    for last in (0, 1, 2):
        for i in range(400):
            p.run(Xs[(i + last + 1) % 3], e_lambda, e_lambda_mu, 0.3, -0.2)
        cnt, s1, s2, ll = p.run(Xs[last], e_lambda, e_lambda_mu, 0.3, -0.2)
        p.check()
        close(s1, want[last][0]), close(s2, want[last][1]), close(ll, want[last][2])
        assert float(cnt) == float(Xs[last].shape[0])
