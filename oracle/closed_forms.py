"""float64 closed-form restatements of the distribution-level quantities on the hot path.

TEST INFRASTRUCTURE.  ``bayesic/distribution/`` in the reference is a non-importable sketch
(syntax errors at base.py:202-205, :217, :223; core.py imports missing modules), so these are
restatements of what those lines clearly intend, each citing the lines it follows.  PARITY
UNPINNED against reference outputs (none can be produced); pinned instead against
scipy.stats / scipy.special in ``tests/test_closed_forms.py``.
"""
import numpy as np
from scipy.special import logsumexp, log_softmax, digamma, gammaln

LOG_2PI = float(np.log(2 * np.pi))


def gaussian_suffstats(X):
    """Iid-summed statistics (x, x x^T) of MultivariateNormal (distribution/core.py:41-44 --
    the code there says ``mean`` where ``data`` is meant) summed over the iid axis
    (distribution/base.py:328-332).  Returns (n, S1[d], S2[d, d]) in float64."""
    X = np.asarray(X, dtype=np.float64)
    return X.shape[0], X.sum(axis=0), X.T @ X


def mvn_natural_parameters(mean, precision):
    """eta = (Lambda mu, -1/2 Lambda)  (distribution/core.py:46-47)."""
    return precision @ mean, -0.5 * precision


def mvn_log_normalizer(mean, precision):
    """Per-draw log-normaliser written as the reference does, i.e. the NEGATIVE of the usual
    A(eta) so that log p = <s, eta> - log_normalizer reads (distribution/base.py:25-100):
        log_normalizer = 1/2 D log 2pi - 1/2 log|Lambda| + 1/2 mu^T Lambda mu.
    (core.py:49-52 has the sign of the 2 pi term flipped)"""
    d = mean.shape[0]
    _, logdet = np.linalg.slogdet(precision)
    return 0.5 * d * LOG_2PI - 0.5 * logdet + 0.5 * mean @ precision @ mean


def mvn_log_likelihood(X, mean, precision):
    """sum_n log N(x_n | mean, precision^-1) = interaction - n * log_normalizer, with the
    interaction term sum_i <flatten s_i, flatten eta_i> (distribution/base.py:279-291) and the
    normaliser multiplied by the number of draws (base.py:235-242)."""
    n, s1, s2 = gaussian_suffstats(X)
    eta1, eta2 = mvn_natural_parameters(mean, precision)
    return s1 @ eta1 + np.sum(s2 * eta2) - n * mvn_log_normalizer(mean, precision)


def gaussian_wishart_expectations(m, beta, W, nu):
    """Moments of q(mu, Lambda) = N(mu | m, (beta Lambda)^-1) Wishart(Lambda | W, nu):
    E[Lambda], E[Lambda mu], E[mu^T Lambda mu], E[log|Lambda|] (Bishop PRML 10.64-10.65)."""
    d = m.shape[0]
    e_lambda = nu * W
    e_lambda_mu = e_lambda @ m
    e_mu_l_mu = d / beta + m @ e_lambda @ m
    _, logdet_w = np.linalg.slogdet(W)
    e_logdet = digamma(0.5 * (nu - np.arange(d))).sum() + d * np.log(2.0) + logdet_w
    return e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet


def gaussian_expected_loglik(n, s1, s2, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet):
    """E_q[ sum_n log N(x_n | mu, Lambda^-1) ] from the statistics: the same
    data/interaction/normaliser split as distribution/base.py:25-100 with expectations of the
    natural parameters substituted (VMP, README.md:30-37)."""
    d = s1.shape[0]
    return (-0.5 * n * d * LOG_2PI + 0.5 * n * e_logdet - 0.5 * np.sum(e_lambda * s2)
            + s1 @ e_lambda_mu - 0.5 * n * e_mu_l_mu)


def log_responsibilities(logits):
    """log r[n, k] = logits[n, k] - logsumexp_k logits[n, :], and the per-row lse."""
    logits = np.asarray(logits, dtype=np.float64)
    return log_softmax(logits, axis=1), logsumexp(logits, axis=1)


def weighted_suffstats(X, R):
    """N_k = sum_n r_nk;  sum_n r_nk x_n;  sum_n r_nk x_n x_n^T  (float64)."""
    X = np.asarray(X, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    return R.sum(axis=0), R.T @ X, np.einsum('nk,nd,ne->kde', R, X, X)
