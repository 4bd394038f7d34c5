"""GPU parity of the fused statistic / responsibility kernels against the float64 oracle
(``oracle/closed_forms.py``), through the C-ABI entry points."""
import numpy as np
import pytest

import bayesic_b200.stats as S
from oracle import closed_forms as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _close(got, want, rtol=RTOL, scale_atol=1e-6):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=scale_atol * max(1.0, np.abs(want).max()))


@pytest.mark.parametrize('n,d', [(1, 64), (127, 64), (128, 64), (129, 64), (1000, 16), (10000, 16),
                                 (65536, 64), (100003, 32), (5000, 60), (777, 48), (2048, 4),
                                 (777, 20), (513, 96), (40, 3), (0, 64)])
def test_gaussian_suffstats_device(n, d):
    import torch
    rng = np.random.RandomState(n * 131 + d)
    X = (rng.randn(n, d) * 1.5 + 0.7).astype(np.float32)
    before = S.launch_count()
    cnt, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    rn, r1, r2 = O.gaussian_suffstats(X)
    assert cnt == rn
    _close(s1.cpu().numpy(), r1)
    _close(s2.cpu().numpy(), r2)
    if n > 0:
        assert S.launch_count() > before


def test_compensated_tf32_is_exact_to_float32_level():
    # low mantissa bits set so that plain TF32 (or RN-on-B) would be off by ~2e-4
    import torch
    rng = np.random.RandomState(3)
    X = (1.0 + 2.0 ** -11 + 2.0 ** -12 + rng.rand(4096, 64) * 2.0 ** -14).astype(np.float32)
    _, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    _, r1, r2 = O.gaussian_suffstats(X)
    np.testing.assert_allclose(s2.cpu().numpy(), r2, rtol=5e-6)
    np.testing.assert_allclose(s1.cpu().numpy(), r1, rtol=1e-6)


def test_cfg1_reference_expression_through_stats():
    # BASELINE cfg1: tr(Lambda . sum x x^T), N = 10 000, D = 16
    import torch
    rng = np.random.RandomState(1234)
    X = rng.randn(10000, 16).astype(np.float32)
    a = rng.randn(16, 16)
    Lam = (a @ a.T / 16 + np.eye(16))
    _, _, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    got = float(np.sum(Lam * s2.cpu().numpy()))
    want = float(np.einsum('de,nd,ne->', Lam, X.astype('f8'), X.astype('f8')))
    assert abs(got - want) <= RTOL * abs(want)


def test_host_streamed_matches_device_and_oracle():
    import torch
    rng = np.random.RandomState(5)
    X = (rng.randn(300001, 64) + 0.25).astype(np.float32)
    _, h1, h2 = S.gaussian_suffstats(X, chunk_rows=65536)
    _, r1, r2 = O.gaussian_suffstats(X)
    _close(h1, r1)
    _close(h2, r2)
    _, d1, d2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    _close(d2.cpu().numpy(), h2, rtol=1e-9)
    pinned = torch.from_numpy(X).pin_memory()
    _, p1, p2 = S.gaussian_suffstats(pinned, chunk_rows=65536)
    np.testing.assert_allclose(p2, h2, rtol=1e-12, atol=1e-12 * np.abs(h2).max())   # same chunking; CTA partials meet in arrival order (float64)
    _, q1, q2 = S.gaussian_suffstats(pinned)       # default chunking
    _close(q2, h2, rtol=1e-5)


@pytest.mark.parametrize('n,d', [(1, 4), (1000, 16), (50000, 64), (4097, 10), (300, 100)])
def test_suffstats_loglik_single_entry_point(n, d):
    """bb_suffstats_gaussian_loglik (ELBO term in the finalize kernel's last block on the tcgen05
    path, generic kernels otherwise) = bb_suffstats_gaussian + bb_gaussian_expected_loglik, also when
    called repeatedly (the block counter must reset) and with recycled output buffers."""
    import torch
    rng = np.random.RandomState(n + d)
    X = torch.from_numpy((rng.randn(n, d) * 1.1 + 0.2).astype(np.float32)).cuda()
    a = rng.randn(d, d)
    e_lambda = torch.from_numpy(a @ a.T / d + np.eye(d)).cuda()
    e_lambda_mu = torch.from_numpy(rng.randn(d)).cuda()
    cnt, s1, s2 = S.gaussian_suffstats(X)
    want = S.gaussian_expected_loglik(cnt, s1, s2, e_lambda, e_lambda_mu, 0.37, -1.9)
    out = None
    for _ in range(3):
        got_n, g1, g2, ell = S.gaussian_suffstats_loglik(X, e_lambda, e_lambda_mu, 0.37, -1.9, out=out)
        out = (g1, g2, ell)
        assert got_n == n
        torch.testing.assert_close(g1, s1, rtol=1e-12, atol=1e-12 * float(s1.abs().max()))    # float64 sums over CTAs in arrival order
        torch.testing.assert_close(g2, s2, rtol=1e-12, atol=1e-12 * float(s2.abs().max()))
        np.testing.assert_allclose(float(ell), float(want), rtol=1e-12)
    ref = O.gaussian_expected_loglik(n, *O.gaussian_suffstats(X.cpu().numpy())[1:], e_lambda.cpu().numpy(),
                                     e_lambda_mu.cpu().numpy(), 0.37, -1.9)
    np.testing.assert_allclose(float(ell), ref, rtol=1e-4)
    with pytest.raises(TypeError):
        S.gaussian_suffstats_loglik(X, e_lambda.float(), e_lambda_mu, 0.37, -1.9)


def test_full_size_cfg2_properties():
    # N = 16 Mi, D = 64 (BASELINE cfg2): size-independent properties instead of a CPU oracle
    import torch
    n, d = 1 << 24, 64
    g = torch.Generator(device='cuda').manual_seed(1234)
    X = torch.randn(n, d, device='cuda', dtype=torch.float32, generator=g) * 1.3 + 0.4
    _, s1, s2 = S.gaussian_suffstats(X)
    # (1) additivity over a split of the data axis (linearity of the statistics)
    _, a1, a2 = S.gaussian_suffstats(X[: n // 3])
    _, b1, b2 = S.gaussian_suffstats(X[n // 3:])
    # (fp32 TMEM accumulation runs over 512-row chunks whose alignment differs between the
    # split and the whole pass, so agreement is to float32-chunk level, not bitwise)
    np.testing.assert_allclose((a2 + b2).cpu().numpy(), s2.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose((a1 + b1).cpu().numpy(), s1.cpu().numpy(), rtol=1e-6)
    # (2) symmetry, (3) checksums: trace = sum of squares, S1 = column sums (float64 on device)
    np.testing.assert_array_equal(s2.cpu().numpy(), s2.cpu().numpy().T)
    tr = float((X.double() ** 2).sum())
    assert abs(float(s2.diagonal().sum()) - tr) <= 5e-6 * tr      # float32-level, far inside rtol 1e-4
    col = X.double().sum(0)
    np.testing.assert_allclose(s1.cpu().numpy(), col.cpu().numpy(), rtol=1e-6)
    # (4) a random projection: u^T S2 v = sum_n (x_n.u)(x_n.v)
    u = torch.randn(d, device='cuda', dtype=torch.float64, generator=None)
    v = torch.randn(d, device='cuda', dtype=torch.float64)
    want = float(((X.double() @ u) * (X.double() @ v)).sum())
    got = float(u @ s2 @ v)
    assert abs(got - want) <= 1e-5 * max(abs(want), float(s2.abs().max()))


def test_gaussian_expected_loglik():
    import torch
    rng = np.random.RandomState(11)
    n, d = 5000, 16
    X = rng.randn(n, d).astype(np.float32)
    m = rng.randn(d) * 0.1
    a = rng.randn(d, d)
    W = np.linalg.inv(a @ a.T / d + np.eye(d)) / (d + 4.0)
    e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet = O.gaussian_wishart_expectations(m, 2.0, W, d + 4.0)
    cnt, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    got = float(S.gaussian_expected_loglik(cnt, s1, s2, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet))
    rn, r1, r2 = O.gaussian_suffstats(X)
    want = O.gaussian_expected_loglik(rn, r1, r2, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet)
    assert abs(got - want) <= RTOL * abs(want)


@pytest.mark.parametrize('n,k', [(4096, 256), (1000, 128), (33, 7), (5, 1000), (64, 1024), (7, 33),
                                 (1, 1), (3, 384), (0, 256)])
def test_log_responsibilities(n, k):
    import torch
    rng = np.random.RandomState(n + 7 * k)
    Lg = (rng.randn(n, k) * 3).astype(np.float32)
    lr, lse, tot = S.log_responsibilities(torch.from_numpy(Lg).cuda())
    ref_lr, ref_lse = O.log_responsibilities(Lg)
    np.testing.assert_allclose(lr.cpu().numpy(), ref_lr, rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse, rtol=RTOL, atol=2e-6)
    assert abs(float(tot) - ref_lse.sum()) <= RTOL * max(1.0, abs(ref_lse.sum()))


@pytest.mark.parametrize('n,k', [(4096, 256), (1000, 128), (33, 7), (5, 1000), (7, 33), (1, 1), (0, 256)])
def test_responsibilities_single_pass_and_in_place(n, k):
    """bb_softmax_rows: r = exp(Lg - lse) in the same single pass as the log-softmax, also in place."""
    import torch
    rng = np.random.RandomState(n + 11 * k)
    Lg = (rng.randn(n, k) * 3).astype(np.float32)
    ref_lr, ref_lse = O.log_responsibilities(Lg)
    dev = torch.from_numpy(Lg).cuda()
    r, lse, tot = S.responsibilities(dev)
    np.testing.assert_allclose(r.cpu().numpy(), np.exp(ref_lr), rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse, rtol=RTOL, atol=2e-6)
    assert abs(float(tot) - ref_lse.sum()) <= RTOL * max(1.0, abs(ref_lse.sum()))
    r2, _, _ = S.responsibilities(dev, want_lse=False, want_sum=False, out=dev)
    assert r2.data_ptr() == dev.data_ptr() and torch.equal(r2, r)


def test_log_responsibilities_extremes():
    import torch
    Lg = np.zeros((4, 256), dtype=np.float32)
    Lg[0, 3] = 100.0                 # would overflow the reference's unstabilised spelling
    Lg[1] = -300.0
    Lg[2, :] = np.linspace(-50, 50, 256)
    Lg[3, 7] = 30.0                  # dominant component: log r must keep relative precision
    lr, lse, _ = S.log_responsibilities(torch.from_numpy(Lg).cuda())
    ref_lr, ref_lse = O.log_responsibilities(Lg)
    got = lr.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref_lr, rtol=RTOL, atol=1e-30)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse, rtol=1e-6)


@pytest.mark.parametrize('n,d,k', [(1000, 64, 8), (4099, 16, 5), (300, 6, 3), (50000, 64, 16), (1, 4, 1),
                                   (2048, 64, 4), (5000, 32, 8), (40000, 64, 256), (1025, 16, 12),
                                   (300000, 64, 8), (6000, 64, 516), (3000, 32, 260), (2500, 16, 1024)])
def test_weighted_suffstats(n, d, k):
    import torch
    rng = np.random.RandomState(n + d + k)
    X = rng.randn(n, d).astype(np.float32)
    R = rng.dirichlet(np.ones(k), size=n).astype(np.float32)
    nk, rx, rxx = S.weighted_suffstats(torch.from_numpy(X).cuda(), torch.from_numpy(R).cuda())
    rnk, rrx, rrxx = O.weighted_suffstats(X, R)
    _close(nk.cpu().numpy(), rnk)
    _close(rx.cpu().numpy(), rrx, scale_atol=1e-5)
    _close(rxx.cpu().numpy(), rrxx, scale_atol=1e-5)
    got = rxx.cpu().numpy()
    np.testing.assert_allclose(got, np.swapaxes(got, 1, 2), rtol=1e-5,
                               atol=1e-6 * np.abs(got).max())               # symmetric


@pytest.mark.parametrize('n,d,k', [(4096, 64, 256), (1, 64, 256), (1031, 32, 256), (5000, 16, 512), (2000, 64, 1024),
                                   (40000, 64, 256)])
def test_weighted_suffstats_from_pre_split_responsibilities(n, d, k):
    """bb_softmax_rows_split writes r = softmax(logits) as the statistics kernel's BF16 operand tiles and
    bb_suffstats_weighted_split consumes them: same statistics as the float64 oracle on exp(log-softmax)."""
    import torch
    rng = np.random.RandomState(n + d + k)
    X = rng.randn(n, d).astype(np.float32)
    Lg = (rng.randn(n, k) * 2.5).astype(np.float32)
    rsplit, lse, tot = S.responsibilities_split(torch.from_numpy(Lg).cuda())
    ref_lr, ref_lse = O.log_responsibilities(Lg)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse, rtol=RTOL, atol=2e-6)
    assert abs(float(tot) - ref_lse.sum()) <= RTOL * max(1.0, abs(ref_lse.sum()))
    nk, rx, rxx = S.weighted_suffstats_split(torch.from_numpy(X).cuda(), rsplit, k)
    rnk, rrx, rrxx = O.weighted_suffstats(X, np.exp(ref_lr))
    _close(nk.cpu().numpy(), rnk)
    _close(rx.cpu().numpy(), rrx, scale_atol=1e-5)
    _close(rxx.cpu().numpy(), rrxx, scale_atol=1e-5)
    # and the float32-R route gives the same numbers to the level of the BF16 split
    r32, _, _ = S.responsibilities(torch.from_numpy(Lg).cuda())
    nk2, rx2, rxx2 = S.weighted_suffstats(torch.from_numpy(X).cuda(), r32)
    np.testing.assert_allclose(rxx.cpu().numpy(), rxx2.cpu().numpy(), rtol=1e-4, atol=1e-5 * float(rxx2.abs().max()))


def test_pre_split_responsibilities_exact_on_one_hot_rows():
    """One-hot responsibilities (a logit 200 above the rest) and small-integer data make every product and sum
    exact: a wrong byte in the operand-tile layout shows up as a whole-number error."""
    import torch
    n, d, k = 3000, 64, 512
    rng = np.random.RandomState(4)
    z = rng.randint(0, k, size=n)
    Lg = np.full((n, k), -100.0, dtype=np.float32)
    Lg[np.arange(n), z] = 100.0
    X = rng.randint(-4, 5, size=(n, d)).astype(np.float32)
    rsplit, _, _ = S.responsibilities_split(torch.from_numpy(Lg).cuda(), want_lse=False, want_sum=False)
    nk, rx, rxx = S.weighted_suffstats_split(torch.from_numpy(X).cuda(), rsplit, k)
    R = np.zeros((n, k)); R[np.arange(n), z] = 1.0
    rnk, rrx, rrxx = O.weighted_suffstats(X, R)
    assert np.array_equal(nk.cpu().numpy(), rnk)
    assert np.array_equal(rx.cpu().numpy(), rrx)
    assert np.array_equal(rxx.cpu().numpy(), rrxx)


def test_weighted_suffstats_sum_over_components_is_the_plain_statistic():
    # responsibilities sum to one per row => sum_k Srxx[k] = S2, sum_k Srx[k] = S1, sum_k Nk = n
    import torch
    rng = np.random.RandomState(9)
    n, d, k = 200000, 64, 16
    X = torch.from_numpy((rng.randn(n, d) + 0.3).astype(np.float32)).cuda()
    R = torch.from_numpy(rng.dirichlet(np.ones(k) * 0.3, size=n).astype(np.float32)).cuda()
    nk, rx, rxx = S.weighted_suffstats(X, R)
    _, s1, s2 = S.gaussian_suffstats(X)
    rowsum = R.double().sum(1, keepdim=True)           # float32 rows do not sum to exactly 1
    Xw = X.double() * rowsum
    np.testing.assert_allclose(rxx.sum(0).cpu().numpy(), (Xw.T @ X.double()).cpu().numpy(), rtol=2e-5,
                               atol=1e-6 * float(s2.abs().max()))
    np.testing.assert_allclose(rx.sum(0).cpu().numpy(), Xw.sum(0).cpu().numpy(), rtol=1e-5)
    # Nk comes out of the same tensor-core contraction (a constant-1 column): same 1e-5 class
    assert abs(float(nk.sum()) - float(rowsum.sum())) <= 2e-5 * n


# ---- regression / Gram statistics (cfg4): X^T X, X^T y, y^T y in one pass -------------------

def _close_gram(got, want, tol=3e-5):
    """Elementwise |err| <= rtol |want| + tol sqrt(S_aa S_bb): the Cauchy-Schwarz scale of entry
    (a, b) is the natural unit for an off-diagonal sum that cancels (float32 BLAS has the same
    shape of error).  Anchors (profiles/r02_parity_report.txt, column `cs`, N = 32 Ki x 1024): this kernel
    1.16e-5, a float32 numpy/BLAS evaluation of the same plan (the reference's own arithmetic) 4.1e-7 -- the
    BF16x3 kernel is ~28x a float32 BLAS on this scale (two-part BF16 operands carry 16 mantissa bits,
    DESIGN.md 4.3) and the coefficient is 2.5x its own measured figure, NOT 2x the float32 reference's."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + tol * scale + 1e-30)


@pytest.mark.parametrize('n,d', [(1, 256), (31, 256), (32, 256), (33, 256), (1000, 256), (4097, 512),
                                 (20000, 1024), (70001, 256), (0, 256), (500, 96), (300, 130), (64, 768),
                                 (3000, 128), (2000, 320), (1500, 68), (900, 1000), (40, 4096)])
def test_regression_suffstats(n, d):
    import torch
    rng = np.random.RandomState(n * 7 + d)
    X = (rng.randn(n, d) * 1.3 + 0.2).astype(np.float32)
    y = (X[:, : min(d, 8)].sum(1) + rng.randn(n)).astype(np.float32)
    before = S.launch_count()
    xtx, xty, yty = S.regression_suffstats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    w_xtx, w_xty, w_yty = O.regression_suffstats(X, y)
    if n > 0:
        assert S.launch_count() > before
        _close_gram(xtx.cpu().numpy(), w_xtx)
    else:
        assert float(xtx.abs().max()) == 0.0
    _close(xty.cpu().numpy(), w_xty, scale_atol=2e-6)
    assert abs(float(yty) - w_yty) <= 1e-6 * max(1.0, w_yty)
    # Gram matrix only (no y) is the same matrix, and exactly symmetric
    only, none1, none2 = S.regression_suffstats(torch.from_numpy(X).cuda())
    assert none1 is None and none2 is None
    assert torch.equal(only, xtx)
    assert torch.equal(xtx, xtx.T)


@pytest.mark.parametrize('d', [512, 200, 68, 1020])
def test_regression_suffstats_exact_on_integers(d):
    """Small integers are exact in bf16, so every product and every fp32 partial sum is exact:
    any layout / quadrant / pipeline mistake in the CTA-pair kernel -- or in the zero padding of the feature
    axis to a multiple of 256 -- shows up as a whole-number error."""
    import torch
    n = 777
    i, j = np.meshgrid(np.arange(n), np.arange(d), indexing='ij')
    X = (((i * 3 + j * 5) % 11) - 5).astype(np.float32)
    y = ((np.arange(n) % 5) - 2).astype(np.float32)
    xtx, xty, yty = S.regression_suffstats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    w_xtx, w_xty, w_yty = O.regression_suffstats(X, y)
    assert np.array_equal(xtx.cpu().numpy(), w_xtx)
    assert np.array_equal(xty.cpu().numpy(), w_xty)
    assert float(yty) == w_yty


def test_regression_suffstats_is_additive_over_row_shards():
    """The statistic is a sum over the data axis (the axis sharded across GPUs): two shards add up."""
    import torch
    rng = np.random.RandomState(5)
    n, d = 30000, 256
    X = torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.randn(n).astype(np.float32)).cuda()
    full = S.regression_suffstats(X, y)
    a = S.regression_suffstats(X[:11111], y[:11111])
    b = S.regression_suffstats(X[11111:], y[11111:])
    for f, pa, pb in zip(full, a, b):
        s = (pa + pb).cpu().numpy()
        np.testing.assert_allclose(s, f.cpu().numpy(), rtol=1e-4, atol=3e-5 * float(full[0].diagonal().max()))


def test_full_size_cfg4_properties():
    """BASELINE cfg4 (minibatch 1 Mi, D = 1024) is too large for the CPU oracle, so check
    size-independent properties against float64 device arithmetic on the same data: exact
    symmetry, the trace (= sum of squares), and the action on a random vector."""
    import torch
    n, d = 1 << 20, 1024
    g = torch.Generator(device='cuda').manual_seed(1234)
    X = torch.randn(n, d, device='cuda', generator=g)
    w = torch.randn(d, device='cuda', generator=g) / d ** 0.5
    y = X @ w + 0.1 * torch.randn(n, device='cuda', generator=g)
    xtx, xty, yty = S.regression_suffstats(X, y)
    assert torch.equal(xtx, xtx.T)
    sq = torch.zeros((), dtype=torch.float64, device='cuda')
    v = torch.randn(d, dtype=torch.float64, device='cuda', generator=g)
    xtxv = torch.zeros(d, dtype=torch.float64, device='cuda')
    ref_xty = torch.zeros(d, dtype=torch.float64, device='cuda')
    for lo in range(0, n, 1 << 17):                     # float64 in slabs of 128 Ki rows
        Xd = X[lo:lo + (1 << 17)].double()
        sq += (Xd * Xd).sum()
        xtxv += Xd.T @ (Xd @ v)
        ref_xty += Xd.T @ y[lo:lo + (1 << 17)].double()
    assert abs(float(xtx.diagonal().sum() - sq)) <= 2e-5 * float(sq)
    err = (xtx @ v - xtxv).abs().max() / xtxv.abs().max()
    assert float(err) <= 2e-5
    assert float((xty - ref_xty).abs().max() / ref_xty.abs().max()) <= 1e-5
    assert abs(float(yty) - float((y.double() ** 2).sum())) <= 1e-6 * float(yty)


# ---- tcgen05 projections (cfg5): Z = X W^T per row, G = X^T R over the data axis ---------------

@pytest.mark.parametrize('n,d,q', [(1, 64, 16), (127, 64, 64), (128, 128, 64), (129, 512, 64), (1000, 256, 128),
                                   (5000, 128, 48), (20000, 512, 64), (257, 64, 256)])
def test_row_projection(n, d, q):
    """dot(X, W.T) on the tcgen05 row-projection kernel vs float64; error measured against
    |x_n| |w_q| (the Cauchy-Schwarz scale of entry (n, q))."""
    import torch
    rng = np.random.RandomState(n + d + q)
    X = (rng.randn(n, d) * 1.3 + 0.2).astype(np.float32)
    W = (rng.randn(q, d) / np.sqrt(d)).astype(np.float32)
    before = S.launch_count()
    Z = S.row_projection(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()).cpu().numpy().astype(np.float64)
    assert S.launch_count() > before
    want = X.astype(np.float64) @ W.astype(np.float64).T
    scale = np.linalg.norm(X.astype(np.float64), axis=1)[:, None] * np.linalg.norm(W.astype(np.float64), axis=1)[None, :]
    assert np.all(np.abs(Z - want) <= RTOL * np.abs(want) + 1e-5 * scale)


@pytest.mark.parametrize('n,d,q', [(1, 128, 64), (15, 128, 64), (16, 128, 64), (17, 256, 64), (1000, 512, 64),
                                   (5000, 256, 128), (333, 128, 256), (4097, 384, 64), (70001, 128, 64)])
def test_column_projection(n, d, q):
    """dot(X.T, R) over the data axis on the tcgen05 column-projection kernel vs float64."""
    import torch
    rng = np.random.RandomState(n + 3 * d + q)
    X = (rng.randn(n, d) * 1.3 + 0.2).astype(np.float32)
    R = rng.randn(n, q).astype(np.float32)
    G = S.column_projection(torch.from_numpy(X).cuda(), torch.from_numpy(R).cuda()).cpu().numpy()
    want = X.astype(np.float64).T @ R.astype(np.float64)
    scale = np.linalg.norm(X.astype(np.float64), axis=0)[:, None] * np.linalg.norm(R.astype(np.float64), axis=0)[None, :]
    # Cauchy-Schwarz-scale coefficient: measured 5.9e-7 at 64 Ki rows (profiles/r02_parity_report.txt, cfg5 G, `cs`;
    # a float32 BLAS: 4.6e-8); 1e-5 leaves room for the short row counts above
    assert np.all(np.abs(G - want) <= RTOL * np.abs(want) + 1e-5 * scale)


def test_projections_exact_on_integers():
    import torch
    rng = np.random.RandomState(11)
    X = rng.randint(-4, 5, size=(1000, 512)).astype(np.float32)
    W = rng.randint(-3, 4, size=(64, 512)).astype(np.float32)
    R = rng.randint(-3, 4, size=(1000, 64)).astype(np.float32)
    Z = S.row_projection(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()).cpu().numpy()
    G = S.column_projection(torch.from_numpy(X).cuda(), torch.from_numpy(R).cuda()).cpu().numpy()
    assert np.array_equal(Z.astype(np.float64), X.astype(np.float64) @ W.astype(np.float64).T)
    assert np.array_equal(G, X.astype(np.float64).T @ R.astype(np.float64))


def test_projection_unsupported_shapes_fail_loudly():
    import torch
    from bayesic_b200.backend.library import BackendError
    X = torch.zeros(64, 96, device='cuda')
    with pytest.raises(BackendError):
        S.row_projection(X, torch.zeros(64, 96, device='cuda'))        # d % 64 != 0
    with pytest.raises(BackendError):
        S.column_projection(X, torch.zeros(64, 64, device='cuda'))     # d % 128 != 0


# ---- mixture logits (cfg3): c_k - 1/2 |U_k x - t_k|^2 on tcgen05 --------------------------------

@pytest.mark.parametrize('triangular', [False, True], ids=['general', 'triangular'])
@pytest.mark.parametrize('n,d,k', [(1, 64, 4), (127, 64, 8), (128, 64, 4), (129, 16, 12), (1000, 32, 8),
                                   (5000, 48, 20), (20000, 64, 256), (4097, 64, 64), (3000, 64, 36),
                                   (2000, 8, 8), (3000, 24, 12), (1500, 40, 16), (2500, 56, 8)])
def test_mixture_logits(n, d, k, triangular):
    """Whitened Gaussian-mixture logits + row log-sum-exp vs float64.  The logits are O(1e2) in
    magnitude and come out of a float32 epilogue, hence the absolute term."""
    import torch
    from scipy.special import logsumexp
    rng = np.random.RandomState(n + d + k)
    centers = rng.randn(k, d) * 2
    X = (centers[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    A = np.stack([_spd_np(rng, d) for _ in range(k)])
    U = np.stack([np.linalg.cholesky(a).T for a in A])                  # A = U^T U, upper triangular
    if not triangular:                                                  # any factor with A = U^T U will do
        Q = np.linalg.qr(rng.randn(d, d))[0]
        U = np.einsum('ij,kjl->kil', Q, U)
    t = np.einsum('kji,ki->kj', U, centers)
    c = rng.randn(k)
    before = S.launch_count()
    logits, lse, total = S.mixture_logits(torch.from_numpy(X).cuda(), torch.from_numpy(U.astype(np.float32)).cuda(),
                                          torch.from_numpy(t.astype(np.float32)).cuda(),
                                          torch.from_numpy(c.astype(np.float32)).cuda(),
                                          upper_triangular=triangular)
    assert S.launch_count() > before
    U32, t32, c32 = U.astype(np.float32).astype(np.float64), t.astype(np.float32).astype(np.float64), c.astype(np.float32).astype(np.float64)
    z = np.einsum('kji,ni->nkj', U32, X.astype(np.float64)) - t32[None]
    want = c32[None, :] - 0.5 * (z * z).sum(-1)
    np.testing.assert_allclose(logits.cpu().numpy(), want, rtol=1e-4, atol=2e-3)
    want_lse = logsumexp(want, axis=1)
    np.testing.assert_allclose(lse.cpu().numpy(), want_lse, rtol=1e-4, atol=2e-3)
    assert abs(float(total) - want_lse.sum()) <= 1e-4 * abs(want_lse.sum()) + 1e-3


def _spd_np(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


# ---- full-size property checks for the tensor-pipe configs (too large for the CPU oracle) ------

def test_full_size_cfg5_properties():
    """BASELINE cfg5 extents (minibatch 4 Mi, D = 512, S = 64): the fused pass against float64
    device arithmetic on the same data, in slabs."""
    import torch
    n, d, s = 1 << 22, 512, 64
    g = torch.Generator(device='cuda').manual_seed(99)
    X = torch.randn(n, d, device='cuda', generator=g)
    W = torch.randn(s, d, device='cuda', generator=g) / d ** 0.5
    y = (torch.rand(n, device='cuda', generator=g) < torch.sigmoid(X @ W[0])).float()
    loglik, G = S.logistic_reparam_stats(X, y, W)
    ll_ref = torch.zeros(s, dtype=torch.float64, device='cuda')
    G_ref = torch.zeros(d, s, dtype=torch.float64, device='cuda')
    Wd = W.double()
    for lo in range(0, n, 1 << 18):
        Xd, yd = X[lo:lo + (1 << 18)].double(), y[lo:lo + (1 << 18)].double()[:, None]
        Z = Xd @ Wd.T
        ll_ref += (yd * Z - torch.nn.functional.softplus(Z)).sum(0)
        G_ref += Xd.T @ (yd - torch.sigmoid(Z))
    assert float(((loglik - ll_ref).abs() / ll_ref.abs()).max()) <= 1e-5
    # G entries are sums of 4 Mi zero-mean-ish terms: scale by |x_d| |resid_s| ~ sqrt(n) * sqrt(n)/2
    assert float((G - G_ref).abs().max()) <= 3e-5 * n ** 0.5 * (n ** 0.5) * 0.5
    assert float((G - G_ref).abs().max() / G_ref.abs().max()) <= 1e-4


def test_large_cfg3_properties():
    """cfg3 extents (K = 256, D = 64) at 2 Mi rows: responsibilities from the tcgen05 logits
    kernel sum to one, the weighted statistics summed over components reproduce the plain
    Gaussian statistics of the same data, and a random subset of logit rows matches float64."""
    import torch
    n, d, k = 1 << 21, 64, 256
    g = torch.Generator(device='cuda').manual_seed(5)
    centers = torch.randn(k, d, device='cuda', generator=g) * 2
    X = centers[torch.randint(k, (n,), device='cuda', generator=g)] + torch.randn(n, d, device='cuda', generator=g)
    U = (torch.eye(d, device='cuda') + 0.05 * torch.randn(k, d, d, device='cuda', generator=g).triu()).contiguous()
    t = torch.einsum('kji,ki->kj', U, centers).contiguous()
    c = torch.randn(k, device='cuda', generator=g)
    logits, lse, total = S.mixture_logits(X, U, t, c, upper_triangular=True)
    rows = torch.randint(n, (4096,), device='cuda', generator=g)
    z = torch.einsum('kji,ni->nkj', U.double(), X[rows].double()) - t.double()[None]
    want = c.double()[None] - 0.5 * (z * z).sum(-1)
    # logits are O(-300) here: rtol 1e-4 (north-star) plus a small absolute term
    assert bool(((logits[rows].double() - want).abs() <= 1e-4 * want.abs() + 1e-3).all())
    want_lse = torch.logsumexp(want, 1)
    assert bool(((lse[rows].double() - want_lse).abs() <= 1e-4 * want_lse.abs() + 1e-3).all())
    assert abs(float(total) - float(lse.double().sum())) <= 1e-6 * abs(float(total))
    R = torch.exp(logits - lse[:, None])
    assert float((R.sum(1) - 1).abs().max()) <= 1e-4
    nk, rx, rxx = S.weighted_suffstats(X, R)
    _, s1, s2 = S.gaussian_suffstats(X)
    rowsum = R.double().sum(1)
    assert abs(float(nk.sum()) - float(rowsum.sum())) <= 2e-5 * n
    Xw = X.double() * rowsum[:, None]
    ref2 = torch.zeros(d, d, dtype=torch.float64, device='cuda')
    for lo in range(0, n, 1 << 19):
        ref2 += Xw[lo:lo + (1 << 19)].T @ X[lo:lo + (1 << 19)].double()
    np.testing.assert_allclose(rxx.sum(0).cpu().numpy(), ref2.cpu().numpy(), rtol=5e-5,
                               atol=3e-5 * float(s2.abs().max()))
    np.testing.assert_allclose(rx.sum(0).cpu().numpy(), Xw.sum(0).cpu().numpy(), rtol=5e-5,
                               atol=3e-5 * float(s1.abs().max()))
    # the default route of the local step at this size: responsibilities as operand tiles, CTA-pair kernel --
    # same statistics as the float32-R kernel to the level of the BF16 split, component by component
    rsplit, lse2, total2 = S.responsibilities_split(logits)
    np.testing.assert_allclose(lse2.cpu().numpy(), lse.cpu().numpy(), rtol=1e-5, atol=1e-3)
    assert abs(float(total2) - float(total)) <= 1e-6 * abs(float(total))
    nk2, rx2, rxx2 = S.weighted_suffstats_split(X, rsplit, k)
    np.testing.assert_allclose(nk2.cpu().numpy(), nk.cpu().numpy(), rtol=2e-5)
    scale = float(rxx.abs().max())
    np.testing.assert_allclose(rxx2.cpu().numpy(), rxx.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(rx2.cpu().numpy(), rx.cpu().numpy(), rtol=1e-4, atol=2e-5 * float(rx.abs().max()))


def test_full_size_cfg3_properties():
    """BASELINE cfg3 at its full extent (N = 64 Mi, D = 64, K = 256: 16 GiB of data, 64 GiB of logits,
    N * K = 2^34 elements -- past 32-bit indexing): the local step without materialising R, checked
    through size-independent properties.  Responsibilities sum to one per row, so the weighted
    statistics summed over components must reproduce the plain Gaussian statistics {N, sum x,
    sum x x^T} of the same data; sum_n lse must equal the float64 sum of the lse vector; logit rows
    sampled from the whole range (including the last rows) must match float64; and the in-place
    log-softmax of the full logit matrix (cfg3b: 64 GiB in place) must agree with the kernel's lse."""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    n, d, k = 1 << 26, 64, 256
    if free < 100 * 2 ** 30:
        pytest.skip('needs ~90 GiB of free device memory')
    g = torch.Generator(device='cuda').manual_seed(11)
    centers = torch.randn(k, d, device='cuda', generator=g) * 2
    X = torch.randn(n, d, device='cuda', generator=g)
    for lo in range(0, n, 1 << 22):                      # mixture structure, built in slabs
        idx = torch.randint(k, (min(1 << 22, n - lo),), device='cuda', generator=g)
        X[lo:lo + idx.shape[0]] += centers[idx]
    U = (torch.eye(d, device='cuda') + 0.05 * torch.randn(k, d, d, device='cuda', generator=g).triu()).contiguous()
    t = torch.einsum('kji,ki->kj', U, centers).contiguous()
    c = torch.randn(k, device='cuda', generator=g)
    logits, lse, total = S.mixture_logits(X, U, t, c, upper_triangular=True)
    rows = torch.cat([torch.randint(n, (2048,), device='cuda', generator=g),
                      torch.arange(n - 64, n, device='cuda'), torch.arange(0, 64, device='cuda')])
    z = torch.einsum('kji,ni->nkj', U.double(), X[rows].double()) - t.double()[None]
    want = c.double()[None] - 0.5 * (z * z).sum(-1)
    assert bool(((logits[rows].double() - want).abs() <= 1e-4 * want.abs() + 1e-3).all())
    want_lse = torch.logsumexp(want, 1)
    assert bool(((lse[rows].double() - want_lse).abs() <= 1e-4 * want_lse.abs() + 1e-3).all())
    assert abs(float(total) - float(lse.double().sum())) <= 1e-6 * abs(float(total))
    nk, rx, rxx = S.weighted_suffstats_from_logits(X, logits, lse)
    cnt, s1, s2 = S.gaussian_suffstats(X)
    assert abs(float(nk.sum()) - n) <= 2e-5 * n
    np.testing.assert_allclose(rx.sum(0).cpu().numpy(), s1.cpu().numpy(), rtol=5e-5, atol=3e-5 * float(s2.abs().max()) ** 0.5 * n ** 0.5)
    np.testing.assert_allclose(rxx.sum(0).cpu().numpy(), s2.cpu().numpy(), rtol=5e-5, atol=3e-5 * float(s2.abs().max()))
    # cfg3b at full size: in-place log-softmax of the 64 GiB logit matrix
    log_resp, lse2, total2 = S.log_responsibilities(logits, out=logits)
    assert bool(((lse2[rows].double() - want_lse).abs() <= 1e-4 * want_lse.abs() + 1e-3).all())
    assert abs(float(total2) - float(total)) <= 1e-6 * abs(float(total))
    resp_rows = torch.exp(log_resp[rows].double()).sum(1)
    assert float((resp_rows - 1).abs().max()) <= 1e-4
    del logits, log_resp, X
    torch.cuda.empty_cache()


@pytest.mark.parametrize('n,d,s', [(1, 128, 64), (100, 128, 64), (129, 256, 64), (2049, 384, 64), (6000, 512, 64),
                                   (40000, 512, 64), (3000, 128, 128)])
def test_logistic_reparam_stats(n, d, s):
    """Fused logistic pass (single-kernel path when S = 64, D % 128 == 0, D <= 512; the two-kernel
    path otherwise) vs float64: loglik[s] and G[d, s]."""
    import torch
    rng = np.random.RandomState(n + d + s)
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(s, d) / np.sqrt(d)).astype(np.float32)
    y = (rng.rand(n) < 1 / (1 + np.exp(-X @ W[0]))).astype(np.float32)
    ll, G = S.logistic_reparam_stats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(W).cuda())
    Z = X.astype(np.float64) @ W.astype(np.float64).T
    want_ll = (y[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)
    resid = y[:, None] - 1.0 / (1.0 + np.exp(-Z))
    want_G = X.astype(np.float64).T @ resid
    np.testing.assert_allclose(ll.cpu().numpy(), want_ll, rtol=1e-4, atol=1e-5)
    scale = np.linalg.norm(X.astype(np.float64), axis=0)[:, None] * np.linalg.norm(resid, axis=0)[None, :]
    assert np.all(np.abs(G.cpu().numpy() - want_G) <= RTOL * np.abs(want_G) + 1e-5 * scale + 1e-12)    # anchors: see test_column_projection


def test_logistic_reparam_stats_saturated_logits():
    """Logits far outside [-20, 20] (sigmoid saturates, exp(-|z|) underflows): softplus must stay
    max(z, 0) + log1p(exp(-|z|)) and the residual exactly y - {0, 1}, as in float64."""
    import torch
    rng = np.random.RandomState(21)
    n, d, s = 5000, 256, 64
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(s, d) * 8.0 / np.sqrt(d)).astype(np.float32)           # |z| up to ~ 40
    W[3] *= 10.0                                                          # one draw with |z| in the hundreds
    y = (rng.rand(n) < 0.5).astype(np.float32)
    ll, G = S.logistic_reparam_stats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(W).cuda())
    Z = X.astype(np.float64) @ W.astype(np.float64).T
    want_ll = (y[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)
    resid = y[:, None] - 1.0 / (1.0 + np.exp(-Z))
    want_G = X.astype(np.float64).T @ resid
    assert np.isfinite(ll.cpu().numpy()).all() and np.isfinite(G.cpu().numpy()).all()
    np.testing.assert_allclose(ll.cpu().numpy(), want_ll, rtol=1e-4)
    scale = np.linalg.norm(X.astype(np.float64), axis=0)[:, None] * np.linalg.norm(resid, axis=0)[None, :]
    assert np.all(np.abs(G.cpu().numpy() - want_G) <= RTOL * np.abs(want_G) + 3e-5 * scale + 1e-12)


@pytest.mark.parametrize('env', [{'BB_LOGISTIC_UNFUSED': '1'}])
def test_logistic_reparam_alternative_kernels(env):
    """The path behind the same entry point that is not the default -- the two-kernel row / column projection
    path (BB_LOGISTIC_UNFUSED=1) -- stays parity-green.  The switch is read once per process, hence the
    subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, torch\n"
        "from bayesic_b200 import stats as S\n"
        "for n, d in [(129, 128), (5000, 256), (2049, 384), (20000, 512)]:\n"
        "    rng = np.random.RandomState(n + d)\n"
        "    X = rng.randn(n, d).astype(np.float32); W = (rng.randn(64, d) / np.sqrt(d)).astype(np.float32)\n"
        "    y = (rng.rand(n) < 0.5).astype(np.float32)\n"
        "    ll, G = S.logistic_reparam_stats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(W).cuda())\n"
        "    Z = X.astype(np.float64) @ W.astype(np.float64).T\n"
        "    want_ll = (y[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)\n"
        "    R = y[:, None] - 1.0 / (1.0 + np.exp(-Z)); want_G = X.astype(np.float64).T @ R\n"
        "    scale = np.linalg.norm(X.astype(np.float64), axis=0)[:, None] * np.linalg.norm(R, axis=0)[None, :]\n"
        "    np.testing.assert_allclose(ll.cpu().numpy(), want_ll, rtol=1e-4, atol=1e-5)\n"
        "    assert np.all(np.abs(G.cpu().numpy() - want_G) <= 1e-4 * np.abs(want_G) + 1e-5 * scale + 1e-12)\n"
        "print('ok')\n")
    run = subprocess.run([sys.executable, '-c', code], cwd=root, env=dict(os.environ, **env), capture_output=True,
                         text=True, timeout=600)
    assert run.returncode == 0 and run.stdout.strip().endswith('ok'), run.stdout + run.stderr


def test_empty_minibatch_through_the_fused_entry_points():
    """n = 0 (an empty shard / minibatch) gives zero statistics and empty per-row outputs, not an error."""
    import torch
    d, k, s = 64, 16, 64
    X = torch.empty((0, d), device='cuda')
    ll, G = S.logistic_reparam_stats(torch.empty((0, 128), device='cuda'), torch.empty(0, device='cuda'),
                                     torch.randn(s, 128, device='cuda'))
    assert float(ll.abs().sum()) == 0.0 and float(G.abs().sum()) == 0.0 and tuple(G.shape) == (128, s)
    U = torch.eye(d, device='cuda').repeat(k, 1, 1).contiguous()
    logits, lse, total = S.mixture_logits(X, U, torch.zeros(k, d, device='cuda'), torch.zeros(k, device='cuda'))
    assert tuple(logits.shape) == (0, k) and tuple(lse.shape) == (0,) and float(total) == 0.0
    nk, rx, rxx = S.weighted_suffstats_from_logits(X, logits, lse)
    assert float(nk.abs().sum()) == 0.0 and float(rx.abs().sum()) == 0.0 and float(rxx.abs().sum()) == 0.0
    log_resp, lse2, total2 = S.log_responsibilities(torch.empty((0, k), device='cuda'))
    assert tuple(log_resp.shape) == (0, k) and float(total2) == 0.0


# ---- error behaviour of the C-ABI entry points (status codes -> exceptions) ----

def test_entry_points_reject_bad_arguments_and_small_workspaces():
    import ctypes
    import torch
    from bayesic_b200.backend import library as L
    lib = L.load()
    X = torch.randn(2048, 256, device='cuda')
    y = torch.randn(2048, device='cuda')
    out = torch.empty(256 * 256 + 256 + 1, dtype=torch.float64, device='cuda')
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # y without xty / yty
    st = lib.bb_suffstats_regression(X.data_ptr(), y.data_ptr(), 2048, 256, out.data_ptr(), None, None,
                                     ws.data_ptr(), ws.numel(), stream)
    assert st == 1 and b'go together' in lib.bb_last_error()               # BB_ERR_INVALID
    with pytest.raises(ValueError):
        L.check(st, 'bb_suffstats_regression')
    # workspace too small
    st = lib.bb_suffstats_regression(X.data_ptr(), None, 2048, 256, out.data_ptr(), None, None, ws.data_ptr(), 16, stream)
    assert st == 5                                                         # BB_ERR_WORKSPACE
    with pytest.raises(L.BackendError):
        L.check(st, 'bb_suffstats_regression')
    # shapes the tensor-core kernels do not serve are refused loudly, never computed elsewhere
    W = torch.randn(64, 250, device='cuda')
    Z = torch.empty(2048, 64, device='cuda')
    st = lib.bb_rowproj(X.data_ptr(), W.data_ptr(), 2048, 250, 64, Z.data_ptr(), ws.data_ptr(), ws.numel(), stream)
    assert st == 3                                                         # BB_ERR_UNSUPPORTED
    st = lib.bb_mixture_logits(X.data_ptr(), X.data_ptr(), X.data_ptr(), X.data_ptr(), 2048, 36, 8, 0, Z.data_ptr(),
                               None, None, ws.data_ptr(), ws.numel(), stream)
    assert st == 3
    st = lib.bb_logistic_reparam_pass(X.data_ptr(), y.data_ptr(), W.data_ptr(), 2048, 250, 64, out.data_ptr(),
                                      out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
    assert st == 3
    torch.cuda.synchronize()


def test_mismatched_shapes_raise_before_any_launch():
    import torch
    X = torch.randn(100, 128, device='cuda')
    with pytest.raises(ValueError):
        S.regression_suffstats(X, torch.randn(99, device='cuda'))
    with pytest.raises(ValueError):
        S.row_projection(X, torch.randn(64, 64, device='cuda'))
    with pytest.raises(ValueError):
        S.column_projection(X, torch.randn(99, 64, device='cuda'))
    with pytest.raises(ValueError):
        S.mixture_logits(torch.randn(10, 64, device='cuda'), torch.randn(4, 64, 64, device='cuda'),
                         torch.randn(4, 32, device='cuda'), torch.randn(4, device='cuda'))
    with pytest.raises(TypeError):
        S.regression_suffstats(np.zeros((4, 4), dtype=np.float32))         # host array: device passes only
