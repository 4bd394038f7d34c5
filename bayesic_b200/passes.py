"""The per-minibatch passes of the five BASELINE configurations, assembled from compiled einsum
plans (``bayesic_b200.algebra`` -> C-ABI executor) and the fused kernels in ``stats``.

The reference names these algorithms only in prose (VMP ``README.md:30-37``, reparameterised
gradients ``:47-51``, minibatch SVI ``:69-80``); what it has as code is the expression layer
they would be written in.  Each pass here is therefore written the way a user of the reference
would write it -- tensor expressions compiled once, called per minibatch -- and every
data-axis contraction lands on a device kernel:

  cfg1/2  Gaussian(-Wishart) statistics + expected log-likelihood   ``gaussian_pass``
  cfg3    GMM VMP local step: logits -> responsibilities -> weighted stats ``GmmStep``
  cfg4    conjugate natural-gradient SVI step for linear regression  ``LinRegSviStep``
  cfg4b   factor-analysis local step from the same Gram pass           ``FactorAnalysisStep``
  cfg5    reparameterised ELBO gradient, logistic regression          ``LogisticReparamGrad``

Global parameters are tiny and replicated; data stays sharded/resident (CUDA tensors in, CUDA
tensors out).  Parameter-space linear algebra (D x D, K x D x D) uses torch float64 on the
device -- plumbing around the hot path, not part of it.
"""
import math

import numpy as np

from . import algebra as A
from . import stats
from . import updates
from .backend.compiled import compile_many

__all__ = ['gaussian_pass', 'GmmStep', 'LinRegSviStep', 'FactorAnalysisStep', 'LogisticReparamGrad',
           'gram_issued_flops_per_row', 'gmm_issued_flops_per_row']

_LOG_2PI = math.log(2.0 * math.pi)


def gram_issued_flops_per_row(d):
    """Tensor-core flops ``gram_pair_kernel`` really issues per data row (csrc/gram_sm100.cu): the
    upper-triangle 256 x 256 feature blocks, BF16x3 = three bf16 products per off-diagonal block and
    two per diagonal block (b1^T b2 + its transpose gives the third)."""
    nb = -(-d // 256)
    off, diag = nb * (nb - 1) // 2, nb
    return 2.0 * 256 * 256 * (3 * off + GRAM_DIAGONAL_PRODUCTS * diag)


GRAM_DIAGONAL_PRODUCTS = 2     # products issued on a diagonal block: b1^T b1 + b1^T (2 b2), symmetrised in the finalize (gram_sm100.cu)


def gmm_issued_flops_per_row(d, k, upper_triangular=True):
    """Tensor-core flops the two kernels of the cfg3 local step issue per data row: the whitened
    projection (BF16x3; 10 of 16 K-steps when the factors are upper triangular at d = 64) and the
    weighted statistics R^T (X (x) X) over 36 symmetric 8 x 8 pair blocks + 1 linear block of 64 columns."""
    steps = d // 16
    kept = (sum(1 for j in range(steps) for i in range(steps) if i >= j) / float(steps * steps)
            if upper_triangular else 1.0)
    nb = d // 8
    blocks = nb * (nb + 1) // 2 + 1
    return 3 * kept * 2.0 * k * d * d + 3 * 2.0 * k * 64 * blocks


def gaussian_pass(X, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet):
    """cfg1/cfg2: ``{n, sum x, sum x x^T}`` in one pass and the expected log-likelihood under
    q(mu, Lambda).  Returns ``(n, s1, s2, elbo_term)`` (float64 CUDA tensors)."""
    torch = stats._torch()
    if isinstance(X, torch.Tensor) and X.is_cuda:
        # resident data: statistics + ELBO term through one entry point (two launches)
        def device_f64(a):
            if isinstance(a, torch.Tensor):
                return a.to(device=X.device, dtype=torch.float64)
            return torch.as_tensor(np.asarray(a, dtype=np.float64), device=X.device)
        return stats.gaussian_suffstats_loglik(X, device_f64(e_lambda), device_f64(e_lambda_mu), e_mu_l_mu, e_logdet)
    n, s1, s2 = stats.gaussian_suffstats(X)
    ell = stats.gaussian_expected_loglik(n, s1, s2, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet)
    return n, s1, s2, ell


class GmmStep(object):
    """cfg3: local VMP step of a Gaussian mixture with Gaussian-Wishart factors.

    logits[n,k] = c_k - nu_k/2 (x_n - m_k)^T W_k (x_n - m_k) is written as one einsum plan
    ``c_k + x.b_k - 1/2 x^T A_k x`` with A_k = nu_k W_k, b_k = A_k m_k; then the fused
    log-softmax kernel gives log r and sum_n lse, and the tcgen05 weighted-statistics kernel
    gives {N_k, sum r x, sum r x x^T} without materialising K x D x N."""

    def __init__(self):
        X, Ak, bk, ck = A.var('X', 2), A.var('Ak', 3), A.var('bk', 2), A.var('ck', 1)
        quad = A.einsum([(X, [('out', 0), ('sum', 0)]), (Ak, [('out', 1), ('sum', 0), ('sum', 1)]),
                         (X, [('out', 0), ('sum', 1)])], 2)
        lin = A.dot(X, bk.T)
        self.logits_fn = (lin + (-0.5) * quad + ck.dimshuffle('x', 0)).compile()
        self.exp_fn = A.exp(A.var('LR', 2)).compile()
        self._whiten_key, self._whiten_value = None, None

    @staticmethod
    def expectations(log_pi, m, beta, W, nu):
        """Per-component (A_k, b_k, c_k) from the variational parameters (float64 numpy in)."""
        from scipy.special import digamma
        k, d = m.shape
        Ak = nu[:, None, None] * W
        bk = np.einsum('kde,ke->kd', Ak, m)
        logdet_w = np.linalg.slogdet(W)[1]
        e_logdet = digamma(0.5 * (nu[:, None] - np.arange(d)[None, :])).sum(1) + d * math.log(2.0) + logdet_w
        ck = (log_pi + 0.5 * e_logdet - 0.5 * d * _LOG_2PI - 0.5 * d / beta
              - 0.5 * np.einsum('kd,kd->k', bk, m))
        return Ak.astype(np.float32), bk.astype(np.float32), ck.astype(np.float32)

    @staticmethod
    def whiten(Ak, bk, ck):
        """(U_k, t_k, c'_k) with A_k = U_k^T U_k, t_k = U_k m_k = U_k^-T b_k, c'_k = c_k + |t_k|^2 / 2,
        so that  c_k + x.b_k - x^T A_k x / 2  =  c'_k - |U_k x - t_k|^2 / 2   (float64 on the device)."""
        import torch
        A64 = Ak.double()
        U = torch.linalg.cholesky(A64, upper=True)
        t = torch.linalg.solve_triangular(U.transpose(1, 2), bk.double().unsqueeze(-1), upper=False).squeeze(-1)
        c = ck.double() + 0.5 * (t * t).sum(1)
        return U.float().contiguous(), t.float().contiguous(), c.float().contiguous()

    @staticmethod
    def local_step(X, U, t, c, materialise=None):
        """The whole local step from WHITENED parameters (``whiten`` once per global update, or the
        ``U, t, c`` that ``updates.gmm_global_update`` emits), three device kernels either way:

        default where the shapes allow (K in {256, 512, 768, 1024}, D % 8 == 0): logits -> responsibilities
        written directly as the statistics kernel's BF16 operand tiles (``bb_softmax_rows_split``: the row pass
        that finds the log-sum-exp also normalises and splits, same bytes as float32 R) -> statistics
        (``bb_suffstats_weighted_split``, which bulk-copies R's tiles and converts only X (x) X);
        ``materialise=True``: logits -> float32 responsibilities in place (``bb_softmax_rows``) -> statistics
        (every column-tile CTA converts R again: the converters bound that kernel);
        ``materialise=False``: logits + row log-sum-exp, then the statistics with r = exp(logit - lse) formed
        inside the operand conversion (two kernels, R never written -- and the exponentials re-evaluated by each of
        the ten column-tile CTAs of a row range: the slowest of the three)."""
        d, k = X.shape[1], U.shape[0]
        if materialise is None and stats.split_responsibilities_supported(d, k):
            logits, _, _ = stats.mixture_logits(X, U, t, c, want_lse=False, want_sum=False, upper_triangular=True)
            rsplit, lse, sum_lse = stats.responsibilities_split(logits)
            nk, rx, rxx = stats.weighted_suffstats_split(X, rsplit, k)
            return {'logits': logits, 'lse': lse, 'sum_lse': sum_lse, 'nk': nk, 'rx': rx, 'rxx': rxx}
        if materialise is None or materialise:
            logits, _, _ = stats.mixture_logits(X, U, t, c, want_lse=False, want_sum=False, upper_triangular=True)
            resp, lse, sum_lse = stats.responsibilities(logits, out=logits)          # in place
            nk, rx, rxx = stats.weighted_suffstats(X, resp)
            return {'resp': resp, 'lse': lse, 'sum_lse': sum_lse, 'nk': nk, 'rx': rx, 'rxx': rxx}
        logits, lse, sum_lse = stats.mixture_logits(X, U, t, c, upper_triangular=True)
        nk, rx, rxx = stats.weighted_suffstats_from_logits(X, logits, lse)
        return {'logits': logits, 'lse': lse, 'sum_lse': sum_lse, 'nk': nk, 'rx': rx, 'rxx': rxx}

    def _whitened(self, Ak, bk, ck):
        """``whiten`` once per parameter update, not per minibatch: cached on the identity and version
        counters of the parameter tensors."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (Ak, bk, ck))
        if self._whiten_key != key:
            self._whiten_value, self._whiten_key = self.whiten(Ak, bk, ck), key
        return self._whiten_value

    def __call__(self, X, Ak, bk, ck, fused=True, want_log_resp=True):
        """Local step from (A_k, b_k, c_k).  Every per-minibatch operation is a device kernel of this
        library: the shapes the tcgen05 kernels serve take ``local_step`` (plus the in-place row normalisation
        of the logits when the log-responsibilities themselves are asked for); other
        shapes take the compiled einsum plan, the log-softmax kernel, a compiled ``exp`` and the generic
        weighted-statistics kernel."""
        d, k = X.shape[1], Ak.shape[0]
        if fused and stats.mixture_logits_supported(d, k) and d % 8 == 0 and k <= 4096 and k % 4 == 0:
            U, t, c = self._whitened(Ak, bk, ck)
            if not want_log_resp:
                return self.local_step(X, U, t, c)
            # a route that keeps the logits: the pre-split one where the shapes allow, else r formed in the kernel
            out = self.local_step(X, U, t, c, None if stats.split_responsibilities_supported(d, k) else False)
            logits = out.pop('logits')
            out['log_resp'], _, _ = stats.log_responsibilities(logits, want_lse=False, want_sum=False, out=logits)
            return out
        logits = self.logits_fn(X=X, Ak=Ak, bk=bk, ck=ck)
        log_resp, lse, sum_lse = stats.log_responsibilities(logits, out=logits)     # in place
        nk, rx, rxx = stats.weighted_suffstats(X, self.exp_fn(LR=log_resp))
        return {'log_resp': log_resp, 'lse': lse, 'sum_lse': sum_lse, 'nk': nk, 'rx': rx, 'rxx': rxx}


class LinRegSviStep(object):
    """cfg4: conjugate natural-gradient SVI for Bayesian linear regression (noise precision tau).
    The minibatch statistics {X^T X, X^T y, y^T y} -- the plans of ``dot(X.T, X)``,
    ``dot(X.T, y)``, ``dot(y, y)`` -- come from ONE fused pass over X
    (``stats.regression_suffstats``; tcgen05 CTA pairs when D % 4 == 0, D > 64), or with
    ``fused=False`` from one compiled multi-output plan; the natural-parameter blend and the
    expected log-likelihood are float64 parameter-space arithmetic."""

    def __init__(self, fused=True):
        self.fused = fused
        X, y = A.var('X', 2), A.var('y', 1)
        self.stats_fn = compile_many([A.dot(X.T, X), A.dot(X.T, y), A.dot(y, y)])

    def __call__(self, X, y, eta1, eta2, tau, n_total, rho, eta1_prior, eta2_prior):
        import torch
        if self.fused:
            xtx, xty, yty = stats.regression_suffstats(X, y)
            yty = yty.reshape(())
        else:
            xtx, xty, yty = self.stats_fn(X=X, y=y)
        b = X.shape[0]
        xtx, xty, yty = xtx.double(), xty.double(), yty.double()
        scale = float(n_total) / float(b)
        new1 = updates.svi_natural_blend(eta1.clone(), eta1_prior, xty, scale * tau, rho)
        new2 = updates.svi_natural_blend(eta2.clone(), eta2_prior, xtx, -0.5 * scale * tau, rho)
        chol = torch.linalg.cholesky(-2.0 * new2)                 # precision of q(w) is SPD
        cov = torch.cholesky_inverse(chol)
        mean = cov @ new1
        e_wwT = cov + torch.outer(mean, mean)
        ell = 0.5 * b * (math.log(tau) - _LOG_2PI) - 0.5 * tau * (yty - 2 * mean @ xty + (e_wwT * xtx).sum())
        return {'xtx': xtx, 'xty': xty, 'yty': yty, 'eta1': new1, 'eta2': new2, 'ell': ell}


class FactorAnalysisStep(object):
    """cfg4, second variant (BASELINE.json: "factor-analysis local step"): the
    local step of x = Lam z + mu + eps, z ~ N(0, I_L), eps ~ N(0, diag psi).  Because
    E[z_n] = G (x_n - mu) is linear in x_n with a shared G = Sigma_z Lam^T Psi^-1, every statistic
    the global step needs -- sum E[z], sum x E[z]^T, sum E[z z^T] -- is a small product with
    {sum x, X^T X}, so the minibatch is read ONCE by the Gram kernel of ``LinRegSviStep``
    (``dot(X.T, X)`` and ``dot(X.T, ones)``); the per-row latent means, when asked for, are the
    plan of ``dot(X, G.T)`` (tcgen05 row projection when the shape fits)."""

    def __init__(self):
        X, G = A.var('X', 2), A.var('G', 2)
        self.latent_fn = A.dot(X, G.T).compile()
        self._ones = None

    def __call__(self, X, Lam, psi, mu, want_latent_means=False):
        import torch
        n, d = X.shape
        if self._ones is None or self._ones.shape[0] != n or self._ones.device != X.device:
            self._ones = torch.ones(n, dtype=torch.float32, device=X.device)
        xtx, sum_x, _ = stats.regression_suffstats(X, self._ones)
        xtx, sum_x = xtx.double(), sum_x.double()
        l = Lam.shape[1]
        lam_p = Lam / psi[:, None]
        sigma_z = torch.cholesky_inverse(torch.linalg.cholesky(
            torch.eye(l, dtype=torch.float64, device=X.device) + Lam.T @ lam_p))
        G = sigma_z @ lam_p.T                                          # [L, D]
        sum_z = G @ (sum_x - n * mu)
        sum_xz = (xtx - torch.outer(sum_x, mu)) @ G.T
        centred = xtx - torch.outer(sum_x, mu) - torch.outer(mu, sum_x) + n * torch.outer(mu, mu)
        sum_zz = n * sigma_z + G @ centred @ G.T
        diag_xx = torch.diagonal(xtx)
        xc_sq = diag_xx - 2 * mu * sum_x + n * mu * mu
        cross = (Lam * (sum_xz - torch.outer(mu, sum_z))).sum(1)
        quad = ((Lam @ sum_zz) * Lam).sum(1)
        ell = -0.5 * n * (d * _LOG_2PI + torch.log(psi).sum()) - 0.5 * ((xc_sq - 2 * cross + quad) / psi).sum()
        out = {'sum_z': sum_z, 'sum_xz': sum_xz, 'sum_zz': sum_zz, 'sum_x': sum_x, 'diag_xx': diag_xx,
               'sigma_z': sigma_z, 'ell': ell}
        if want_latent_means:
            out['Ez'] = self.latent_fn(X=X, G=G.float().contiguous()) - (G @ mu).float()[None, :]
        return out


class LogisticReparamGrad(object):
    """cfg5: reparameterised ELBO gradient for Bayesian logistic regression, S fixed draws.
    Written in the reference's vocabulary: softplus(z) = log(1 + exp(z)),
    sigmoid(z) = (1 + exp(-z))**-1 (``algebra.py:1435-1448`` has log/exp/pow only)."""

    def __init__(self):
        X, y, Wm = A.var('X', 2), A.var('y', 1), A.var('Wm', 2)        # Wm[S, D]
        Z = A.dot(X, Wm.T)                                            # [B, S]
        ycol = y.dimshuffle(0, 'x')
        loglik = A.sum(ycol * Z - A.log(1 + A.exp(Z)), axis=0)        # [S]
        resid = ycol - (1 + A.exp(-1 * Z)) ** -1                      # [B, S]
        G = A.dot(X.T, resid)                                         # [D, S]
        self.fn = compile_many([loglik, G])

    def __call__(self, X, y, mu, log_sigma, eps, fused=True):
        import torch
        if fused and stats.logistic_reparam_supported(X.shape[1], eps.shape[0]):
            # the same two plans on the tcgen05 projection kernels, elementwise chain fused; draws
            # and gradient assembly are device kernels too (nothing syncs with the host)
            loglik, G = stats.logistic_reparam_stats(X, y, updates.reparam_draws(mu, log_sigma, eps))
            elbo, grad_mu, grad_ls = updates.reparam_gradient(G, loglik, eps, mu, log_sigma)
            return {'elbo': elbo.reshape(()), 'G': G, 'grad_mu': grad_mu, 'grad_log_sigma': grad_ls}
        sigma = torch.exp(log_sigma)
        Wm = (mu[None, :] + sigma[None, :] * eps).to(torch.float32)
        loglik, G = self.fn(X=X, y=y, Wm=Wm)
        G = G.double()
        kl = 0.5 * torch.sum(sigma ** 2 + mu ** 2 - 1.0 - 2.0 * log_sigma)
        grad_mu = G.mean(dim=1) - mu
        grad_ls = (G * eps.T).mean(dim=1) * sigma - sigma ** 2 + 1.0
        return {'elbo': loglik.double().mean() - kl, 'G': G, 'grad_mu': grad_mu,
                'grad_log_sigma': grad_ls}
