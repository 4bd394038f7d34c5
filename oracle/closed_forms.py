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


# ---- BASELINE configs 3-5: float64 restatements of the passes (north-star text; README.md:30-37,
# 47-51, 69-80 describe the algorithms, the reference has no code for them) -----------------------

def gmm_expected_logits(X, log_pi, m, beta, W, nu):
    """E_q[log pi_k] + E_q[log N(x_n | mu_k, Lambda_k^-1)] for a Gaussian mixture with
    Gaussian-Wishart factors (Bishop PRML 10.46, 10.64-10.66):
      logits[n,k] = log_pi[k] + 1/2 E[log|L_k|] - D/2 log 2pi - D/(2 beta_k)
                    - nu_k/2 (x_n - m_k)^T W_k (x_n - m_k)"""
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    k = m.shape[0]
    out = np.empty((n, k))
    for j in range(k):
        _, logdet_w = np.linalg.slogdet(W[j])
        e_logdet = digamma(0.5 * (nu[j] - np.arange(d))).sum() + d * np.log(2.0) + logdet_w
        diff = X - m[j]
        quad = np.einsum('nd,de,ne->n', diff, W[j], diff)
        out[:, j] = log_pi[j] + 0.5 * e_logdet - 0.5 * d * LOG_2PI - 0.5 * d / beta[j] - 0.5 * nu[j] * quad
    return out


def gmm_vmp_step(X, log_pi, m, beta, W, nu):
    """One local step of mean-field VMP for a GMM: logits -> log-softmax responsibilities ->
    {N_k, sum r x, sum r x x^T} and sum_n logsumexp (the data-dependent part of the ELBO)."""
    logits = gmm_expected_logits(X, log_pi, m, beta, W, nu)
    log_r, lse = log_responsibilities(logits)
    nk, rx, rxx = weighted_suffstats(X, np.exp(log_r))
    return {'logits': logits, 'log_resp': log_r, 'nk': nk, 'rx': rx, 'rxx': rxx, 'sum_lse': lse.sum()}


def regression_suffstats(X, y):
    """float64 X^T X, X^T y, y^T y: the declared value of the plans of dot(X.T, X), dot(X.T, y),
    dot(y, y) (bayesic/algebra.py:1151-1158 -> 527-551)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return X.T @ X, X.T @ y, float(y @ y)


def linreg_svi_step(X, y, eta1, eta2, tau, n_total, rho, eta1_prior, eta2_prior):
    """Conjugate natural-gradient SVI step for Bayesian linear regression with known noise
    precision tau (Hoffman et al. 2013, README.md:69-80): minibatch statistics
    {X^T X, X^T y, y^T y, B}, then
      eta <- (1 - rho) eta + rho (eta_prior + (N_total / B) [tau X^T y, -1/2 tau X^T X]),
    and the expected log-likelihood of the minibatch under q(w) = N(mean, cov) (from the
    updated natural parameters)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    b, d = X.shape
    xtx, xty, yty = X.T @ X, X.T @ y, y @ y
    scale = n_total / float(b)
    new1 = (1 - rho) * eta1 + rho * (eta1_prior + scale * tau * xty)
    new2 = (1 - rho) * eta2 + rho * (eta2_prior - 0.5 * scale * tau * xtx)
    cov = np.linalg.inv(-2.0 * new2)
    mean = cov @ new1
    e_wwT = cov + np.outer(mean, mean)
    ell = 0.5 * b * (np.log(tau) - LOG_2PI) - 0.5 * tau * (yty - 2 * mean @ xty + np.sum(e_wwT * xtx))
    return {'xtx': xtx, 'xty': xty, 'yty': yty, 'eta1': new1, 'eta2': new2, 'ell': ell}


def logistic_reparam_gradient(X, y, mu, log_sigma, eps):
    """Reparameterised ELBO gradient for Bayesian logistic regression with q(w) = N(mu,
    diag sigma^2), prior N(0, I) and S fixed standard-normal draws eps[S, D] (README.md:47-51):
      W_s = mu + sigma * eps_s;  Z = X W^T;  l = y z - softplus(z)
      elbo = mean_s sum_n l - KL;  G = X^T (y - sigmoid(Z))  [D, S]
      grad_mu = mean_s G_s - mu;  grad_log_sigma = mean_s G_s * eps_s * sigma - sigma^2 + 1"""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    sigma = np.exp(log_sigma)
    Wm = mu[None, :] + sigma[None, :] * eps                  # [S, D]
    Z = X @ Wm.T                                             # [B, S]
    softplus = np.logaddexp(0.0, Z)
    ll = (y[:, None] * Z - softplus).sum(axis=0)             # [S]
    G = X.T @ (y[:, None] - 1.0 / (1.0 + np.exp(-Z)))        # [D, S]
    kl = 0.5 * np.sum(sigma ** 2 + mu ** 2 - 1.0 - 2.0 * log_sigma)
    grad_mu = G.mean(axis=1) - mu
    grad_ls = (G * eps.T).mean(axis=1) * sigma - sigma ** 2 + 1.0
    return {'Z': Z, 'elbo': ll.mean() - kl, 'G': G, 'grad_mu': grad_mu, 'grad_log_sigma': grad_ls}


def _log_wishart_b(W, nu):
    """log B(W, nu), the Wishart normaliser (Bishop PRML B.79)."""
    d = W.shape[0]
    _, logdet = np.linalg.slogdet(W)
    return (-0.5 * nu * logdet - (0.5 * nu * d * np.log(2.0) + 0.25 * d * (d - 1) * np.log(np.pi)
                                  + gammaln(0.5 * (nu - np.arange(d))).sum()))


def gmm_global_update(nk, rx, rxx, alpha0, beta0, nu0, m0, W0_inv):
    """VMP global step of a Gaussian mixture with Dirichlet(alpha0) weights and Gaussian-Wishart
    (m0, beta0, W0, nu0) components, from the local step's statistics (Bishop PRML 10.58-10.63;
    the reference names the algorithm at README.md:30-37 and has no code for it):
      alpha_k = alpha0 + N_k; beta_k = beta0 + N_k; nu_k = nu0 + N_k; m_k = (beta0 m0 + sum r x) / beta_k
      W_k^-1 = W0^-1 + N_k S_k + beta0 N_k / (beta0 + N_k) (xbar_k - m0)(xbar_k - m0)^T
    written with sums so that N_k = 0 is fine, plus the quantities the next local step needs
    (E[log pi_k], W_k) and the KL terms of the ELBO (10.74-10.77 regrouped as KL(q || p))."""
    nk = np.asarray(nk, dtype=np.float64)
    k, d = rx.shape
    alpha, beta, nu = alpha0 + nk, beta0 + nk, nu0 + nk
    m = (beta0 * m0[None, :] + rx) / beta[:, None]
    W0 = np.linalg.inv(W0_inv)
    W_inv = np.empty((k, d, d))
    W = np.empty((k, d, d))
    kl = np.empty(k + 1)
    for j in range(k):
        W_inv[j] = (W0_inv + rxx[j] + beta0 * np.outer(m0, m0) - beta[j] * np.outer(m[j], m[j]))
        W_inv[j] = 0.5 * (W_inv[j] + W_inv[j].T)
        W[j] = np.linalg.inv(W_inv[j])
        _, logdet_w = np.linalg.slogdet(W[j])
        e_logdet = digamma(0.5 * (nu[j] - np.arange(d))).sum() + d * np.log(2.0) + logdet_w
        dm = m[j] - m0
        kl_wishart = (_log_wishart_b(W[j], nu[j]) - _log_wishart_b(W0, nu0) + 0.5 * (nu[j] - nu0) * e_logdet
                      - 0.5 * nu[j] * d + 0.5 * nu[j] * np.trace(W0_inv @ W[j]))
        kl_gauss = 0.5 * (d * beta0 / beta[j] + beta0 * nu[j] * dm @ W[j] @ dm - d + d * np.log(beta[j] / beta0))
        kl[j] = kl_wishart + kl_gauss
    e_log_pi = digamma(alpha) - digamma(alpha.sum())
    kl[k] = (gammaln(alpha.sum()) - gammaln(alpha).sum() - gammaln(k * alpha0) + k * gammaln(alpha0)
             + ((alpha - alpha0) * e_log_pi).sum())
    return {'alpha': alpha, 'beta': beta, 'nu': nu, 'm': m, 'W_inv': W_inv, 'W': W, 'e_log_pi': e_log_pi, 'kl': kl}


def adam_step(param, grad, m, v, lr, b1, b2, eps, step, maximize=False):
    """One Adam step (Kingma & Ba 2015), float64; returns (param, m, v)."""
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    update = lr * (m / (1 - b1 ** step)) / (np.sqrt(v / (1 - b2 ** step)) + eps)
    return (param + update if maximize else param - update), m, v


def factor_analysis_local_step(X, Lam, psi, mu):
    """Local (E) step of factor analysis x = Lam z + mu + eps, z ~ N(0, I_L), eps ~ N(0, diag psi)
    (the second minibatch-SVI variant BASELINE.json cfg4 names; README.md:69-80), written per row:
      Sigma_z = (I + Lam^T Psi^-1 Lam)^-1;  E[z_n] = Sigma_z Lam^T Psi^-1 (x_n - mu)
      sum_n E[z_n], sum_n x_n E[z_n]^T, sum_n E[z_n z_n^T] = N Sigma_z + sum_n E[z_n] E[z_n]^T,
      sum_n x_n, diag sum_n x_n x_n^T
    and the expected complete-data log-likelihood of the minibatch given those expectations."""
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    l = Lam.shape[1]
    sigma_z = np.linalg.inv(np.eye(l) + Lam.T @ (Lam / psi[:, None]))
    G = sigma_z @ (Lam / psi[:, None]).T                     # [L, D]
    Ez = (X - mu) @ G.T                                      # [N, L]
    sum_z = Ez.sum(0)
    sum_xz = X.T @ Ez
    sum_zz = n * sigma_z + Ez.T @ Ez
    sum_x = X.sum(0)
    diag_xx = (X * X).sum(0)
    # E log p(x | z) summed over rows, with residual e = x - mu - Lam z
    xc_sq = diag_xx - 2 * mu * sum_x + n * mu * mu
    cross = np.einsum('dl,dl->d', Lam, sum_xz - np.outer(mu, sum_z))
    quad = np.einsum('dl,lm,dm->d', Lam, sum_zz, Lam)
    ell = -0.5 * n * (d * LOG_2PI + np.log(psi).sum()) - 0.5 * ((xc_sq - 2 * cross + quad) / psi).sum()
    return {'Ez': Ez, 'sum_z': sum_z, 'sum_xz': sum_xz, 'sum_zz': sum_zz, 'sum_x': sum_x, 'diag_xx': diag_xx,
            'sigma_z': sigma_z, 'ell': ell}
