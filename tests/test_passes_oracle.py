"""CPU checks of the float64 restatements for configs 3-5 (oracle/closed_forms.py) against
independent formulations (scipy / finite differences), since no reference output can exist."""
import numpy as np
from scipy.special import expit

from oracle import closed_forms as O


def _spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def test_gmm_logits_match_the_gaussian_wishart_expectation():
    rng = np.random.RandomState(0)
    n, d, k = 50, 3, 4
    X = rng.randn(n, d)
    m, beta, nu = rng.randn(k, d), rng.rand(k) + 1, d + 1 + rng.rand(k) * 3
    W = np.stack([np.linalg.inv(_spd(rng, d)) / nu[j] for j in range(k)])
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    logits = O.gmm_expected_logits(X, log_pi, m, beta, W, nu)
    for j in range(k):
        e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet = O.gaussian_wishart_expectations(m[j], beta[j], W[j], nu[j])
        for i in (0, n - 1):
            x = X[i]
            want = log_pi[j] + O.gaussian_expected_loglik(1, x, np.outer(x, x), e_lambda, e_lambda_mu,
                                                          e_mu_l_mu, e_logdet)
            np.testing.assert_allclose(logits[i, j], want, rtol=1e-10)
    step = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    np.testing.assert_allclose(step['nk'].sum(), n, rtol=1e-12)


def test_linreg_svi_step_fixed_point_is_the_exact_posterior():
    # rho = 1, full batch: natural parameters of the exact conjugate posterior
    rng = np.random.RandomState(1)
    b, d, tau = 200, 5, 4.0
    X, y = rng.randn(b, d), rng.randn(b)
    out = O.linreg_svi_step(X, y, np.zeros(d), -0.5 * np.eye(d), tau, b, 1.0, np.zeros(d), -0.5 * np.eye(d))
    prec = np.eye(d) + tau * X.T @ X
    np.testing.assert_allclose(-2 * out['eta2'], prec, rtol=1e-12)
    np.testing.assert_allclose(out['eta1'], tau * X.T @ y, rtol=1e-12)


def test_logistic_gradient_matches_finite_differences():
    rng = np.random.RandomState(2)
    b, d, s = 60, 4, 5
    X, y = rng.randn(b, d), (rng.rand(b) < 0.5).astype(float)
    mu, ls, eps = rng.randn(d) * 0.2, np.log(0.3 + 0.1 * rng.rand(d)), rng.randn(s, d)
    out = O.logistic_reparam_gradient(X, y, mu, ls, eps)
    h = 1e-6
    for i in range(d):
        dm = np.zeros(d); dm[i] = h
        up = O.logistic_reparam_gradient(X, y, mu + dm, ls, eps)['elbo']
        dn = O.logistic_reparam_gradient(X, y, mu - dm, ls, eps)['elbo']
        np.testing.assert_allclose(out['grad_mu'][i], (up - dn) / (2 * h), rtol=1e-5, atol=1e-6)
        up = O.logistic_reparam_gradient(X, y, mu, ls + dm, eps)['elbo']
        dn = O.logistic_reparam_gradient(X, y, mu, ls - dm, eps)['elbo']
        np.testing.assert_allclose(out['grad_log_sigma'][i], (up - dn) / (2 * h), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(1 / (1 + np.exp(-out['Z'])), expit(out['Z']), rtol=1e-12)
