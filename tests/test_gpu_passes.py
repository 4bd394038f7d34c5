"""GPU parity of the five BASELINE-config passes (bayesic_b200/passes.py) against their float64
restatements in oracle/closed_forms.py, at sizes the oracle finishes in seconds.  Tolerance:
rtol 1e-4 (north-star), atol scaled to the magnitude of each output."""
import numpy as np
import pytest

import bayesic_b200.passes as P
from oracle import closed_forms as O

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-4, scale_atol=2e-6):
    got = np.asarray(got.detach().cpu().numpy() if hasattr(got, 'detach') else got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=rtol, atol=scale_atol * max(1.0, np.abs(want).max()))


def _spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def test_cfg2_gaussian_pass():
    import torch
    rng = np.random.RandomState(0)
    n, d = 20000, 64
    X = (rng.randn(n, d) * 1.2 + 0.3).astype(np.float32)
    ex = O.gaussian_wishart_expectations(rng.randn(d) * 0.1, 2.0, np.linalg.inv(_spd(rng, d)) / (d + 4.0), d + 4.0)
    cnt, s1, s2, ell = P.gaussian_pass(torch.from_numpy(X).cuda(), *ex)
    rn, r1, r2 = O.gaussian_suffstats(X)
    _close(s1, r1)
    _close(s2, r2)
    _close(ell, [O.gaussian_expected_loglik(rn, r1, r2, *ex)])


def test_cfg3_gmm_vmp_step():
    import torch
    rng = np.random.RandomState(1)
    n, d, k = 4096, 16, 8
    centers = rng.randn(k, d) * 3
    X = (centers[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    m = centers + rng.randn(k, d) * 0.1
    beta = rng.rand(k) * 5 + 1
    nu = d + 2 + rng.rand(k) * 5
    W = np.stack([np.linalg.inv(_spd(rng, d)) / nu[j] for j in range(k)])
    want = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    step = P.GmmStep()
    Ak, bk, ck = step.expectations(log_pi, m, beta, W, nu)
    got = step(torch.from_numpy(X).cuda(), torch.from_numpy(Ak).cuda(), torch.from_numpy(bk).cuda(),
               torch.from_numpy(ck).cuda())
    # log-responsibilities: float32 logits of magnitude ~1e2 limit the absolute accuracy
    np.testing.assert_allclose(got['log_resp'].cpu().numpy(), want['log_resp'], rtol=1e-4, atol=2e-3)
    _close(got['nk'], want['nk'], rtol=1e-4, scale_atol=1e-5)
    _close(got['rx'], want['rx'], rtol=1e-4, scale_atol=1e-4)
    _close(got['rxx'], want['rxx'], rtol=1e-4, scale_atol=1e-4)
    assert abs(float(got['sum_lse']) - want['sum_lse']) <= 1e-4 * abs(want['sum_lse'])


def test_cfg4_linreg_svi_step():
    import torch
    rng = np.random.RandomState(2)
    b, d = 8192, 128
    X = rng.randn(b, d).astype(np.float32)
    w_true = rng.randn(d) / np.sqrt(d)
    y = (X @ w_true + 0.1 * rng.randn(b)).astype(np.float32)
    tau, n_total, rho = 100.0, 10 * b, 0.3
    eta1_prior, eta2_prior = np.zeros(d), -0.5 * np.eye(d)
    eta1, eta2 = rng.randn(d) * 0.1, -0.5 * _spd(rng, d) * 10
    want = O.linreg_svi_step(X, y, eta1, eta2, tau, n_total, rho, eta1_prior, eta2_prior)
    dev = torch.device('cuda')
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    got = P.LinRegSviStep()(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), t(eta1), t(eta2), tau,
                            n_total, rho, t(eta1_prior), t(eta2_prior))
    # D = 128 runs on the tcgen05 CTA-pair kernel too (feature axis zero-padded to 256): BF16x3 class of error
    _close(got['xtx'], want['xtx'], scale_atol=3e-5)
    _close(got['xty'], want['xty'], scale_atol=1e-5)
    _close(got['yty'], want['yty'])
    _close(got['eta1'], want['eta1'], scale_atol=1e-5)
    _close(got['eta2'], want['eta2'], scale_atol=3e-5)
    assert abs(float(got['ell']) - want['ell']) <= 1e-4 * 0.5 * tau * want['yty']


def test_cfg5_logistic_reparam_gradient():
    import torch
    rng = np.random.RandomState(3)
    b, d, s = 4096, 64, 16
    X = rng.randn(b, d).astype(np.float32)
    w_true = rng.randn(d) / np.sqrt(d)
    y = (rng.rand(b) < 1 / (1 + np.exp(-X @ w_true))).astype(np.float32)
    mu, log_sigma = rng.randn(d) * 0.1, np.log(0.1 + 0.05 * rng.rand(d))
    eps = rng.randn(s, d)
    want = O.logistic_reparam_gradient(X, y, mu, log_sigma, eps)
    dev = torch.device('cuda')
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    got = P.LogisticReparamGrad()(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), t(mu), t(log_sigma), t(eps))
    _close(got['G'], want['G'], scale_atol=2e-5)
    _close(got['grad_mu'], want['grad_mu'], scale_atol=2e-5)
    _close(got['grad_log_sigma'], want['grad_log_sigma'], scale_atol=2e-5)
    assert abs(float(got['elbo']) - want['elbo']) <= 1e-4 * abs(want['elbo'])


def test_cfg4_linreg_svi_step_on_the_cta_pair_kernel():
    """Same step at D = 256 (the tcgen05 CTA-pair Gram kernel), fused pass and compiled-plan route
    (dot(X.T, X) lowers to the SYRK node, which dispatches to the same kernel)."""
    import torch
    rng = np.random.RandomState(4)
    b, d = 6000, 256
    X = rng.randn(b, d).astype(np.float32)
    w_true = rng.randn(d) / np.sqrt(d)
    y = (X @ w_true + 0.1 * rng.randn(b)).astype(np.float32)
    tau, n_total, rho = 100.0, 10 * b, 0.3
    eta1_prior, eta2_prior = np.zeros(d), -0.5 * np.eye(d)
    eta1, eta2 = rng.randn(d) * 0.1, -0.5 * _spd(rng, d) * 10
    want = O.linreg_svi_step(X, y, eta1, eta2, tau, n_total, rho, eta1_prior, eta2_prior)
    dev = torch.device('cuda')
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    for fused in (True, False):
        got = P.LinRegSviStep(fused=fused)(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), t(eta1),
                                          t(eta2), tau, n_total, rho, t(eta1_prior), t(eta2_prior))
        _close(got['xtx'], want['xtx'], scale_atol=3e-5)
        _close(got['xty'], want['xty'], scale_atol=1e-5)
        _close(got['yty'], want['yty'])
        _close(got['eta1'], want['eta1'], scale_atol=1e-5)
        _close(got['eta2'], want['eta2'], scale_atol=3e-5)
        # ell = const - tau/2 (yty - 2 m.xty + <E[ww^T], xtx>) cancels ~100-fold here (tau = 100,
        # residual variance 0.01), so 1e-4 is taken relative to the terms that cancel
        assert abs(float(got['ell']) - want['ell']) <= 1e-4 * 0.5 * tau * want['yty']


def test_cfg5_logistic_reparam_gradient_on_the_projection_kernels():
    """Same pass at D = 512, S = 64 (BASELINE cfg5's feature/sample extents) through the fused
    tcgen05 projection kernels, and through the compiled plans (fused=False) as a cross-check."""
    import torch
    rng = np.random.RandomState(6)
    b, d, s = 6000, 512, 64
    X = rng.randn(b, d).astype(np.float32)
    w_true = rng.randn(d) / np.sqrt(d)
    y = (rng.rand(b) < 1 / (1 + np.exp(-X @ w_true))).astype(np.float32)
    mu, log_sigma = rng.randn(d) * 0.05, np.log(0.05 + 0.02 * rng.rand(d))
    eps = rng.randn(s, d)
    want = O.logistic_reparam_gradient(X, y, mu, log_sigma, eps)
    dev = torch.device('cuda')
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    step = P.LogisticReparamGrad()
    for fused in (True, False):
        got = step(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), t(mu), t(log_sigma), t(eps), fused=fused)
        _close(got['G'], want['G'], scale_atol=2e-5)
        _close(got['grad_mu'], want['grad_mu'], scale_atol=2e-5)
        _close(got['grad_log_sigma'], want['grad_log_sigma'], scale_atol=2e-5)
        assert abs(float(got['elbo']) - want['elbo']) <= 1e-4 * abs(want['elbo'])


def test_cfg3_gmm_vmp_step_on_the_tensor_core_kernels():
    """Same step at D = 64 with enough rows for the tcgen05 kernels (whitened logits, regrouped
    weighted statistics), and the plan route (fused=False) at the same inputs as a cross-check."""
    import torch
    rng = np.random.RandomState(7)
    n, d, k = 6000, 64, 8
    centers = rng.randn(k, d) * 1.5
    X = (centers[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    m = centers + rng.randn(k, d) * 0.05
    beta = rng.rand(k) * 5 + 1
    nu = d + 2 + rng.rand(k) * 5
    W = np.stack([np.linalg.inv(_spd(rng, d)) / nu[j] for j in range(k)])
    want = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    step = P.GmmStep()
    Ak, bk, ck = step.expectations(log_pi, m, beta, W, nu)
    for fused in (True, False):
        got = step(torch.from_numpy(X).cuda(), torch.from_numpy(Ak).cuda(), torch.from_numpy(bk).cuda(),
                   torch.from_numpy(ck).cuda(), fused=fused)
        np.testing.assert_allclose(got['log_resp'].cpu().numpy(), want['log_resp'], rtol=1e-4, atol=3e-3)
        _close(got['nk'], want['nk'], rtol=1e-4, scale_atol=2e-5)
        _close(got['rx'], want['rx'], rtol=1e-4, scale_atol=1e-4)
        _close(got['rxx'], want['rxx'], rtol=1e-4, scale_atol=1e-4)
        assert abs(float(got['sum_lse']) - want['sum_lse']) <= 1e-4 * abs(want['sum_lse'])


def test_cfg3_local_step_at_k256_uses_pre_split_responsibilities():
    """K = 256 (BASELINE cfg3's extent): the default local step writes the responsibilities as operand tiles."""
    import torch
    rng = np.random.RandomState(18)
    n, d, k = 7000, 64, 256
    centers = rng.randn(k, d) * 1.5
    X = (centers[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    m = centers + rng.randn(k, d) * 0.05
    beta, nu = rng.rand(k) * 5 + 1, d + 2 + rng.rand(k) * 5
    W = np.stack([np.linalg.inv(_spd(rng, d)) / nu[j] for j in range(k)])
    want = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    step = P.GmmStep()
    Ak, bk, ck = (torch.from_numpy(a).cuda() for a in step.expectations(log_pi, m, beta, W, nu))
    Xd = torch.from_numpy(X).cuda()
    outs = [step(Xd, Ak, bk, ck, want_log_resp=False), step.local_step(Xd, *step.whiten(Ak, bk, ck), materialise=True)]
    assert 'resp' not in outs[0] and 'logits' in outs[0]          # pre-split route: no float32 R
    for out in outs:
        _close(out['nk'], want['nk'], rtol=1e-4, scale_atol=2e-5)
        _close(out['rx'], want['rx'], rtol=1e-4, scale_atol=1e-4)
        _close(out['rxx'], want['rxx'], rtol=1e-4, scale_atol=1e-4)
        assert abs(float(out['sum_lse']) - want['sum_lse']) <= 1e-4 * abs(want['sum_lse'])
    log_resp = (outs[0]['logits'] - outs[0]['lse'][:, None]).cpu().numpy()
    np.testing.assert_allclose(log_resp, want['log_resp'], rtol=1e-4, atol=3e-3)


def test_cfg3_local_step_both_routes():
    """``local_step`` from whitened parameters at a K the pre-split route does not serve: ``materialise=True``
    (logits -> responsibilities in place -> statistics) and ``materialise=False`` (logits + lse, r formed
    inside the statistics kernel, R never written); same statistics as the oracle's full step either way."""
    import torch
    rng = np.random.RandomState(8)
    n, d, k = 5000, 64, 12
    centers = rng.randn(k, d) * 1.5
    X = (centers[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    m = centers + rng.randn(k, d) * 0.05
    beta = rng.rand(k) * 5 + 1
    nu = d + 2 + rng.rand(k) * 5
    W = np.stack([np.linalg.inv(_spd(rng, d)) / nu[j] for j in range(k)])
    want = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    step = P.GmmStep()
    Ak, bk, ck = (torch.from_numpy(a).cuda() for a in step.expectations(log_pi, m, beta, W, nu))
    Xd = torch.from_numpy(X).cuda()
    got = step.local_step(Xd, *step.whiten(Ak, bk, ck), materialise=True)
    assert 'log_resp' not in got and 'logits' not in got
    np.testing.assert_allclose(got['resp'].cpu().numpy(), np.exp(want['log_resp']), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(got['resp'].sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    other = step.local_step(Xd, *step.whiten(Ak, bk, ck), materialise=False)
    log_resp = (other['logits'] - other['lse'][:, None]).cpu().numpy()
    np.testing.assert_allclose(log_resp, want['log_resp'], rtol=1e-4, atol=3e-3)
    for out in (got, other):
        _close(out['nk'], want['nk'], rtol=1e-4, scale_atol=2e-5)
        _close(out['rx'], want['rx'], rtol=1e-4, scale_atol=1e-4)
        _close(out['rxx'], want['rxx'], rtol=1e-4, scale_atol=1e-4)
        assert abs(float(out['sum_lse']) - want['sum_lse']) <= 1e-4 * abs(want['sum_lse'])
        np.testing.assert_allclose(out['lse'].cpu().numpy(), got['lse'].cpu().numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('n,d,l', [(20000, 256, 16), (5000, 96, 5), (8192, 1024, 32)])
def test_cfg4_factor_analysis_local_step(n, d, l):
    """Second cfg4 variant: all statistics from one Gram pass, latent means from the row projection."""
    import torch
    rng = np.random.RandomState(11)
    Lam, psi, mu = rng.randn(d, l) / np.sqrt(l), rng.rand(d) + 0.5, rng.randn(d) * 0.3
    X = (rng.randn(n, l) @ Lam.T + mu + rng.randn(n, d) * np.sqrt(psi)).astype(np.float32)
    want = O.factor_analysis_local_step(X, Lam, psi, mu)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).cuda()
    got = P.FactorAnalysisStep()(torch.from_numpy(X).cuda(), t(Lam), t(psi), t(mu), want_latent_means=True)
    for key in ('sum_x', 'diag_xx', 'sum_xz', 'sum_zz', 'sigma_z'):
        _close(got[key], want[key], scale_atol=2e-5)
    # sum E[z] is a difference of O(n) terms that cancel to O(sqrt n): tolerance relative to the terms
    np.testing.assert_allclose(got['sum_z'].cpu().numpy(), want['sum_z'], atol=2e-5 * n)
    _close(got['Ez'], want['Ez'], rtol=1e-4, scale_atol=5e-5)
    np.testing.assert_allclose(float(got['ell']), want['ell'], rtol=1e-5)
