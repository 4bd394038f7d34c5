"""Pins the float64 closed-form oracle (oracle/closed_forms.py) against scipy.stats / scipy.special,
since no reference output can exist for bayesic/distribution/ (it does not parse)."""
import numpy as np
from scipy.special import logsumexp
from scipy.stats import multivariate_normal, wishart

from oracle import closed_forms as O


def _spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def test_mvn_log_likelihood_matches_scipy():
    rng = np.random.RandomState(0)
    X, mean, prec = rng.randn(200, 5), rng.randn(5), _spd(rng, 5)
    want = multivariate_normal(mean, np.linalg.inv(prec)).logpdf(X).sum()
    np.testing.assert_allclose(O.mvn_log_likelihood(X, mean, prec), want, rtol=1e-12)


def test_suffstats_are_the_iid_sums():
    rng = np.random.RandomState(1)
    X = rng.randn(37, 4).astype(np.float32)
    n, s1, s2 = O.gaussian_suffstats(X)
    assert n == 37
    np.testing.assert_allclose(s1, sum(x.astype('f8') for x in X), rtol=1e-13)
    np.testing.assert_allclose(s2, sum(np.outer(x, x).astype('f8') for x in X.astype('f8')), rtol=1e-13)


def test_gaussian_wishart_expectations_by_monte_carlo_and_identities():
    rng = np.random.RandomState(2)
    d, beta, nu = 3, 2.5, 9.0
    m, W = rng.randn(d), np.linalg.inv(_spd(rng, d)) / nu
    e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet = O.gaussian_wishart_expectations(m, beta, W, nu)
    np.testing.assert_allclose(e_lambda, nu * W, rtol=1e-13)
    np.testing.assert_allclose(e_lambda_mu, nu * W @ m, rtol=1e-13)
    samples = wishart(df=nu, scale=W).rvs(size=40000, random_state=3)
    logdets = np.linalg.slogdet(samples)[1]
    assert abs(logdets.mean() - e_logdet) < 4 * logdets.std() / np.sqrt(len(logdets))
    # E[mu^T Lambda mu] with mu | Lambda ~ N(m, (beta Lambda)^-1): d/beta + m^T E[Lambda] m
    np.testing.assert_allclose(e_mu_l_mu, d / beta + m @ e_lambda @ m, rtol=1e-13)


def test_expected_loglik_reduces_to_loglik_for_a_point_posterior():
    # with E[Lambda]=Lambda, E[Lambda mu]=Lambda mu, E[mu' Lambda mu]=mu' Lambda mu, E[log|L|]=log|L|
    rng = np.random.RandomState(4)
    X, mean, prec = rng.randn(100, 4), rng.randn(4), _spd(rng, 4)
    n, s1, s2 = O.gaussian_suffstats(X)
    got = O.gaussian_expected_loglik(n, s1, s2, prec, prec @ mean, mean @ prec @ mean,
                                     np.linalg.slogdet(prec)[1])
    np.testing.assert_allclose(got, O.mvn_log_likelihood(X, mean, prec), rtol=1e-12)


def test_log_responsibilities_and_weighted_stats():
    rng = np.random.RandomState(5)
    Lg = rng.randn(50, 7) * 4
    lr, lse = O.log_responsibilities(Lg)
    np.testing.assert_allclose(lse, logsumexp(Lg, axis=1), rtol=1e-13)
    np.testing.assert_allclose(np.exp(lr).sum(1), 1.0, rtol=1e-12)
    X, R = rng.randn(50, 3), np.exp(lr)
    nk, rx, rxx = O.weighted_suffstats(X, R)
    np.testing.assert_allclose(nk.sum(), 50.0, rtol=1e-12)
    np.testing.assert_allclose(rx.sum(0), X.sum(0), rtol=1e-12)
    np.testing.assert_allclose(rxx.sum(0), X.T @ X, rtol=1e-12)


def test_gmm_global_update_matches_bishop_and_scipy():
    """The VMP global step: W_k^-1 in Bishop's (10.62) grouping, KL = 0 at the prior, the
    Gaussian-Wishart / Dirichlet KL terms against scipy densities by Monte Carlo, and the
    logit constants against gmm_expected_logits."""
    from scipy.stats import dirichlet
    rng = np.random.RandomState(7)
    d, k, n = 3, 2, 60
    m0, W0_inv, alpha0, beta0, nu0 = rng.randn(d), _spd(rng, d) * d, 1.5, 2.0, d + 2.0
    zero = O.gmm_global_update(np.zeros(k), np.zeros((k, d)), np.zeros((k, d, d)), alpha0, beta0, nu0, m0, W0_inv)
    np.testing.assert_allclose(zero['kl'], 0.0, atol=1e-10)
    X, R = rng.randn(n, d) + 1.0, rng.dirichlet(np.ones(k), size=n)
    nk, rx, rxx = O.weighted_suffstats(X, R)
    out = O.gmm_global_update(nk, rx, rxx, alpha0, beta0, nu0, m0, W0_inv)
    xbar = rx / nk[:, None]
    for j in range(k):
        S = rxx[j] / nk[j] - np.outer(xbar[j], xbar[j])
        want = W0_inv + nk[j] * S + beta0 * nk[j] / (beta0 + nk[j]) * np.outer(xbar[j] - m0, xbar[j] - m0)
        np.testing.assert_allclose(out['W_inv'][j], want, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out['m'], (beta0 * m0 + nk[:, None] * xbar) / (beta0 + nk)[:, None], rtol=1e-12)
    # Monte Carlo KL of component 0 (Wishart part by densities, Gaussian part in closed form given Lambda)
    W, nu, beta, m = out['W'][0], out['nu'][0], out['beta'][0], out['m'][0]
    lam = wishart(df=nu, scale=W).rvs(size=60000, random_state=1)
    lam_t = np.moveaxis(lam, 0, -1)
    log_ratio = wishart(df=nu, scale=W).logpdf(lam_t) - wishart(df=nu0, scale=np.linalg.inv(W0_inv)).logpdf(lam_t)
    dm = m - m0
    gauss = 0.5 * (d * beta0 / beta + beta0 * np.einsum('i,nij,j->n', dm, lam, dm) - d + d * np.log(beta / beta0))
    mc = log_ratio + gauss
    assert abs(mc.mean() - out['kl'][0]) < 5 * mc.std() / np.sqrt(len(mc))
    draws = dirichlet(out['alpha']).rvs(60000, random_state=2)
    mc = dirichlet(out['alpha']).logpdf(draws.T) - dirichlet(np.full(k, alpha0)).logpdf(draws.T)
    assert abs(mc.mean() - out['kl'][k]) < 5 * mc.std() / np.sqrt(len(mc))
    # E[log pi] sums consistently and feeds the logits
    np.testing.assert_allclose(np.exp(out['e_log_pi']).sum() < 1.0, True)
    logits = O.gmm_expected_logits(X, out['e_log_pi'], out['m'], out['beta'], out['W'], out['nu'])
    assert np.isfinite(logits).all()


def test_adam_step_first_step_moves_by_lr():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -3.0])
    new, m, v = O.adam_step(p, g, np.zeros(2), np.zeros(2), 0.1, 0.9, 0.999, 1e-12, 1)
    np.testing.assert_allclose(new, p - 0.1 * np.sign(g), rtol=1e-9)
    new, _, _ = O.adam_step(p, g, np.zeros(2), np.zeros(2), 0.1, 0.9, 0.999, 1e-12, 1, maximize=True)
    np.testing.assert_allclose(new, p + 0.1 * np.sign(g), rtol=1e-9)


def test_factor_analysis_local_step_by_brute_force():
    rng = np.random.RandomState(9)
    n, d, l = 120, 6, 3
    Lam, psi, mu = rng.randn(d, l), rng.rand(d) + 0.5, rng.randn(d)
    X = rng.randn(n, l) @ Lam.T + mu + rng.randn(n, d) * np.sqrt(psi)
    out = O.factor_analysis_local_step(X, Lam, psi, mu)
    # posterior mean = conditional mean of the joint Gaussian (Woodbury-free form)
    np.testing.assert_allclose(out['Ez'], (X - mu) @ np.linalg.solve(Lam @ Lam.T + np.diag(psi), Lam), rtol=1e-10)
    ell = 0.0
    for i in range(n):
        e = X[i] - mu - Lam @ out['Ez'][i]
        ell += (-0.5 * (d * np.log(2 * np.pi) + np.log(psi).sum()) - 0.5 * (e * e / psi).sum()
                - 0.5 * np.trace((Lam / psi[:, None]).T @ Lam @ out['sigma_z']))
    np.testing.assert_allclose(out['ell'], ell, rtol=1e-12)
    np.testing.assert_allclose(out['sum_zz'], n * out['sigma_z'] + out['Ez'].T @ out['Ez'], rtol=1e-12)
