"""GPU parity of the parameter-space update kernels (bayesic_b200/updates.py -> C-ABI) against
the float64 restatements in oracle/closed_forms.py: VMP global step of a Gaussian mixture, SVI
natural-parameter blend, reparameterised draws / gradient assembly, Adam; and the two loops
built from them (GmmVmp, LogisticReparamSgd) against oracle loops.  These kernels compute in
float64, so the tolerances are float64-ish except where a float32 output is specified."""
import numpy as np
import pytest

import bayesic_b200.updates as Up
from bayesic_b200 import stats
from bayesic_b200.backend import library as L
from oracle import closed_forms as O

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _np(t):
    return t.detach().cpu().numpy()


def _spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def _mixture_stats(rng, n, d, k):
    X = rng.randn(n, d) * 0.8 + rng.randn(k, d)[rng.randint(k, size=n)] * 2.0
    R = rng.dirichlet(np.ones(k) * 0.3, size=n)
    return X, O.weighted_suffstats(X, R)


@pytest.mark.parametrize('k,d', [(3, 5), (8, 16), (256, 64), (5, 96)])
def test_gmm_global_update_matches_oracle(k, d):
    rng = np.random.RandomState(k + d)
    X, (nk, rx, rxx) = _mixture_stats(rng, 40 * k + 200, d, k)
    nk[0], rx[0], rxx[0] = 0.0, 0.0, 0.0                    # an empty component stays at the prior
    m0, W0_inv, alpha0, beta0, nu0 = rng.randn(d) * 0.1, _spd(rng, d), 0.7, 1.3, d + 1.5
    want = O.gmm_global_update(nk, rx, rxx, alpha0, beta0, nu0, m0, W0_inv)
    got = Up.gmm_global_update(_t(nk), _t(rx), _t(rxx), alpha0, beta0, nu0, _t(m0), _t(W0_inv))
    assert int(_np(got['status'])[0]) == 0
    for key in ('alpha', 'beta', 'nu', 'm', 'W_inv'):
        np.testing.assert_allclose(_np(got[key]), want[key], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(_np(got['kl']), want['kl'], rtol=1e-9, atol=1e-8)
    assert abs(_np(got['kl'])[0]) < 1e-8                     # KL(prior || prior)
    U = _np(got['U']).astype(np.float64)
    assert np.abs(np.tril(U, -1)).max() == 0.0               # upper triangular: the kernel's skip is exact
    for j in range(k):
        a = want['nu'][j] * want['W'][j]
        np.testing.assert_allclose(U[j].T @ U[j], a, rtol=2e-5, atol=2e-6 * np.abs(a).max())
        np.testing.assert_allclose(_np(got['t'])[j], U[j] @ want['m'][j], rtol=2e-5, atol=2e-5 * np.abs(U[j]).max())
    logdet_w = np.array([np.linalg.slogdet(w)[1] for w in want['W']])
    from scipy.special import digamma
    e_logdet = digamma(0.5 * (want['nu'][:, None] - np.arange(d)[None, :])).sum(1) + d * np.log(2.0) + logdet_w
    c = want['e_log_pi'] + 0.5 * e_logdet - 0.5 * d * O.LOG_2PI - 0.5 * d / want['beta']
    np.testing.assert_allclose(_np(got['c']), c, rtol=1e-6, atol=1e-5)


def test_gmm_global_update_feeds_the_logits_kernel():
    """c_k - |U_k x - t_k|^2 / 2 from the update kernel's outputs is the expected log joint of
    Bishop 10.46 under the updated posterior."""
    import torch
    rng = np.random.RandomState(3)
    n, d, k = 4096, 32, 16
    X, (nk, rx, rxx) = _mixture_stats(rng, n, d, k)
    m0, W0_inv = np.zeros(d), np.eye(d)
    want = O.gmm_global_update(nk, rx, rxx, 1.0, 1.0, d + 2.0, m0, W0_inv)
    got = Up.gmm_global_update(_t(nk), _t(rx), _t(rxx), 1.0, 1.0, d + 2.0, _t(m0), _t(W0_inv))
    X32 = X.astype(np.float32)
    logits, lse, _ = stats.mixture_logits(torch.from_numpy(X32).cuda(), got['U'], got['t'], got['c'], upper_triangular=True)
    ref = O.gmm_expected_logits(X32, want['e_log_pi'], want['m'], want['beta'], want['W'], want['nu'])
    np.testing.assert_allclose(_np(logits), ref, rtol=1e-4, atol=1e-3)


def test_gmm_global_update_flags_a_non_spd_component():
    rng = np.random.RandomState(4)
    d, k = 8, 4
    _, (nk, rx, rxx) = _mixture_stats(rng, 500, d, k)
    rxx[2] = -50.0 * np.eye(d)                                 # impossible statistics
    got = Up.gmm_global_update(_t(nk), _t(rx), _t(rxx), 1.0, 1.0, d + 1.0, _t(np.zeros(d)), _t(np.eye(d)))
    assert int(_np(got['status'])[0]) == 3                     # 1 + component index
    assert np.isnan(_np(got['U'])[2]).all() and np.isnan(_np(got['c'])[2])
    assert np.isfinite(_np(got['U'])[[0, 1, 3]]).all()


def test_gmm_global_update_argument_errors():
    d, k = 4, 2
    z = lambda *s: _t(np.zeros(s))
    with pytest.raises(L.BackendError):                         # d > 96: BB_ERR_UNSUPPORTED
        Up.gmm_global_update(z(k), z(k, 100), z(k, 100, 100), 1.0, 1.0, 101.0, z(100), _t(np.eye(100)))
    with pytest.raises(ValueError):                             # nu0 <= d - 1
        Up.gmm_global_update(z(k), z(k, d), z(k, d, d), 1.0, 1.0, d - 1.0, z(d), _t(np.eye(d)))
    with pytest.raises(ValueError):
        Up.gmm_global_update(z(k), z(k, d), z(k, d + 1, d), 1.0, 1.0, d + 1.0, z(d), _t(np.eye(d)))
    with pytest.raises(TypeError):
        Up.gmm_global_update(np.zeros(k), z(k, d), z(k, d, d), 1.0, 1.0, d + 1.0, z(d), _t(np.eye(d)))


def test_svi_blend_draws_gradient_adam_match_oracle():
    rng = np.random.RandomState(5)
    d, s = 300, 24
    eta, prior, stat = rng.randn(d, d), rng.randn(d, d), rng.randn(d, d)
    got = Up.svi_natural_blend(_t(eta), _t(prior), _t(stat), 3.5, 0.3)
    np.testing.assert_allclose(_np(got), 0.7 * eta + 0.3 * (prior + 3.5 * stat), rtol=1e-14, atol=1e-14)
    with pytest.raises(ValueError):
        Up.svi_natural_blend(_t(eta), _t(prior), _t(stat), 1.0, 1.5)
    mu, ls, eps = rng.randn(d) * 0.1, rng.randn(d) * 0.1 - 1.0, rng.randn(s, d)
    W = Up.reparam_draws(_t(mu), _t(ls), _t(eps))
    np.testing.assert_array_equal(_np(W), (mu[None, :] + np.exp(ls)[None, :] * eps).astype(np.float32))
    G, ll = rng.randn(d, s) * 10, rng.randn(s) * 100
    elbo, gm, gs = Up.reparam_gradient(_t(G), _t(ll), _t(eps), _t(mu), _t(ls))
    sg = np.exp(ls)
    np.testing.assert_allclose(_np(gm), G.mean(1) - mu, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(_np(gs), (G * eps.T).mean(1) * sg - sg ** 2 + 1.0, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(_np(elbo)[0], ll.mean() - 0.5 * np.sum(sg ** 2 + mu ** 2 - 1.0 - 2.0 * ls), rtol=1e-12)
    p, g, m, v = rng.randn(d), rng.randn(d), rng.rand(d) * 0.1, rng.rand(d) * 0.01
    for maximize in (False, True):
        tp, tm, tv = _t(p), _t(m), _t(v)
        Up.adam_step(tp, _t(g), tm, tv, 4, lr=0.05, maximize=maximize)
        wp, wm, wv = O.adam_step(p, g, m, v, 0.05, 0.9, 0.999, 1e-8, 4, maximize=maximize)
        np.testing.assert_allclose(_np(tp), wp, rtol=1e-13)
        np.testing.assert_allclose(_np(tm), wm, rtol=1e-13)
        np.testing.assert_allclose(_np(tv), wv, rtol=1e-13)


def test_gmm_vmp_loop_tracks_the_float64_oracle_and_raises_the_elbo():
    import torch
    rng = np.random.RandomState(6)
    n, d, k = 6000, 16, 8
    centres = rng.randn(k, d) * 3.0
    X = (centres[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    m0, W0_inv, alpha0, beta0, nu0 = np.zeros(d), np.eye(d), 1.0, 1.0, d + 2.0
    R0 = rng.dirichlet(np.ones(k), size=n)
    nk, rx, rxx = O.weighted_suffstats(X, R0)
    vmp = Up.GmmVmp(k, d, alpha0, beta0, nu0, _t(m0), _t(W0_inv))
    vmp.initialise(_t(nk), _t(rx), _t(rxx))
    ref = O.gmm_global_update(nk, rx, rxx, alpha0, beta0, nu0, m0, W0_inv)
    Xd = torch.from_numpy(X).cuda()
    elbos = []
    for it in range(4):
        state = vmp.step(Xd)
        local = O.gmm_vmp_step(X, ref['e_log_pi'], ref['m'], ref['beta'], ref['W'], ref['nu'])
        ref_elbo = local['sum_lse'] - ref['kl'].sum()
        ref = O.gmm_global_update(local['nk'], local['rx'], local['rxx'], alpha0, beta0, nu0, m0, W0_inv)
        elbos.append(float(_np(state['elbo'])[0]))
        np.testing.assert_allclose(elbos[-1], ref_elbo, rtol=2e-5)
        np.testing.assert_allclose(_np(state['alpha']), ref['alpha'], rtol=1e-3, atol=1e-2)
        np.testing.assert_allclose(_np(state['m']), ref['m'], rtol=1e-3, atol=2e-3)
        assert int(_np(state['status'])[0]) == 0
    assert all(b >= a - 1e-6 * abs(a) for a, b in zip(elbos, elbos[1:]))       # VMP never lowers the ELBO


def test_logistic_reparam_sgd_loop_matches_the_oracle_loop():
    import torch
    rng = np.random.RandomState(8)
    n, d, s = 8192, 128, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 1 / (1 + np.exp(-X @ (rng.randn(d) / np.sqrt(d))))).astype(np.float32)
    mu, ls, eps = np.zeros(d), np.full(d, -2.0), rng.randn(s, d)
    loop = Up.LogisticReparamSgd(_t(mu), _t(ls), _t(eps), lr=0.02)
    moments = [np.zeros(d) for _ in range(4)]
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    for step in range(1, 4):
        got = loop.step(Xd, yd)
        want = O.logistic_reparam_gradient(X, y, mu, ls, eps)
        np.testing.assert_allclose(float(_np(got['elbo'])[0]), want['elbo'], rtol=1e-4)
        scale = np.abs(want['grad_mu']).max()
        np.testing.assert_allclose(_np(got['grad_mu']), want['grad_mu'], rtol=1e-4, atol=1e-4 * scale)
        mu, moments[0], moments[1] = O.adam_step(mu, want['grad_mu'], moments[0], moments[1], 0.02, 0.9, 0.999,
                                                 1e-8, step, maximize=True)
        ls, moments[2], moments[3] = O.adam_step(ls, want['grad_log_sigma'], moments[2], moments[3], 0.02, 0.9,
                                                 0.999, 1e-8, step, maximize=True)
        np.testing.assert_allclose(_np(loop.mu), mu, rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(_np(loop.log_sigma), ls, rtol=1e-3, atol=1e-4)
