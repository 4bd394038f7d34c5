"""Exponential families beyond the reference's two: the likelihoods and conjugate
partners the BASELINE configurations name -- Bernoulli in logit form (cfg5), Categorical (mixture
assignments, cfg3), Gamma / Exponential (precisions, rates) and Dirichlet (mixture weights) --
written on the same contract as ``core.py``: sufficient statistics and natural parameters are
algebra expressions, so iid sums come out as data-axis contractions and ``match`` can pair
statistics with parameters.

The reference has no code for these (``bayesic/distribution/core.py`` stops at the two Gaussians;
``README.md:30-37`` names the conjugate-exponential setting).  Log-normalisers that need
``log Gamma`` use the ``lgamma`` extension of the algebra vocabulary; checked against
``scipy.stats`` in ``tests/test_distribution.py``.

  BernoulliLogit(logit):      s = (x,),        eta = (logit,),     A = log(1 + exp(logit))
  Exponential(rate):          s = (x,),        eta = (-rate,),     A = -log(rate)
  Gamma(shape, rate):         s = (log x, x),  eta = (a - 1, -b),  A = lgamma(a) - a log(b)
  Categorical(logits[K]):     s = (x[K],)      one-hot,  eta = (logits,),  A = log sum_k exp(logits_k)
  Dirichlet(concentration[K]): s = (log x,),   eta = (alpha - 1,), A = sum_k lgamma(alpha_k) - lgamma(sum_k alpha_k)
  Wishart(df, scale_inverse V = W^-1) over a precision matrix L[D, D]:
                              s = (log|L|, L), eta = ((nu - D - 1)/2, -V/2),
                              A = nu D/2 log 2 - nu/2 log|V| + log Gamma_D(nu/2)
  GaussianWishart(mean m, beta, df, scale_inverse V) over (mu[D], L[D, D]),
      p = N(mu | m, (beta L)^-1) W(L | V^-1, nu):
                              s = (L mu, mu^T L mu, L, log|L|),
                              eta = (beta m, -beta/2, -(V + beta m m^T)/2, (nu - D)/2),
                              A = D/2 log 2pi - D/2 log beta + nu D/2 log 2 - nu/2 log|V| + log Gamma_D(nu/2)
``log|.|`` is the ``logdet`` node (``BB_NODE_LOGDET``; the reference's MVN normaliser calls a
``T.logdet`` Theano never had, ``distribution/core.py:49-52``).  Both are parameterised by the INVERSE
scale so that every natural parameter is multilinear in the parameters (no matrix inverse in the
vocabulary); ``updates.gmm_global_update`` keeps W_k^-1 for the same reason.
"""
import math

import numpy as np


from .. import algebra as A
from .base import ExponentialFamily

floatX = 'float32'

__all__ = ['BernoulliLogit', 'Exponential', 'Gamma', 'Categorical', 'Dirichlet', 'Wishart', 'GaussianWishart']


class _ScalarDatum(ExponentialFamily):
    """Scalar observations: <s, eta> is a plain (broadcasting) product per observation."""

    data_type = (floatX, 0)

    def log_likelihood_data_term(self, data):
        return A.constant(0)

    def _pair(self, stat, natural):
        stat, natural = A.wrap_if_literal(stat), A.wrap_if_literal(natural)
        if natural.ndim == 0 or natural.ndim == stat.ndim:
            return stat * natural
        return ExponentialFamily._pair(self, stat, natural)


class BernoulliLogit(_ScalarDatum):
    """x in {0, 1} with log-odds ``logit``: log p = x logit - log(1 + exp(logit))."""

    parameter_types = dict(logit=(floatX, 0))

    def is_discrete(self):
        return True

    def sufficient_statistics(self, data):
        return (A.wrap_if_literal(data),)

    def natural_parameters(self, logit):
        return (A.wrap_if_literal(logit),)

    def log_normalizer(self, logit, data_shape=None):
        return A.log(1 + A.exp(A.wrap_if_literal(logit)))


class Exponential(_ScalarDatum):
    """x > 0 with density rate * exp(-rate x)."""

    parameter_types = dict(rate=(floatX, 0))

    def sufficient_statistics(self, data):
        return (A.wrap_if_literal(data),)

    def natural_parameters(self, rate):
        return (-1 * A.wrap_if_literal(rate),)

    def log_normalizer(self, rate, data_shape=None):
        return -1 * A.log(A.wrap_if_literal(rate))


class Gamma(_ScalarDatum):
    """x > 0 with density b^a / Gamma(a) x^(a-1) exp(-b x)  (shape a, rate b)."""

    parameter_types = dict(shape=(floatX, 0), rate=(floatX, 0))

    def sufficient_statistics(self, data):
        data = A.wrap_if_literal(data)
        return A.log(data), data

    def natural_parameters(self, shape, rate):
        return A.wrap_if_literal(shape) + (-1), -1 * A.wrap_if_literal(rate)

    def log_normalizer(self, shape, rate, data_shape=None):
        shape, rate = A.wrap_if_literal(shape), A.wrap_if_literal(rate)
        return A.lgamma(shape) - shape * A.log(rate)


class _VectorDatum(ExponentialFamily):
    data_type = (floatX, 1)

    def log_likelihood_data_term(self, data):
        return A.constant(0)

    def _core_ndim(self, stat):
        return 1


class Categorical(_VectorDatum):
    """One-hot x[K] with unnormalised log-probabilities ``logits[K]``.  The log-normaliser is
    written in the reference's vocabulary, log(sum(exp(.))) -- fine for logits of moderate size; the
    mixture passes use the fused, stabilised log-softmax kernel instead."""

    parameter_types = dict(logits=(floatX, 1))

    def is_discrete(self):
        return True

    def sufficient_statistics(self, data):
        return (A.wrap_if_literal(data),)

    def natural_parameters(self, logits):
        return (A.wrap_if_literal(logits),)

    def log_normalizer(self, logits, data_shape=None):
        logits = A.wrap_if_literal(logits)
        return A.log(A.sum(A.exp(logits), axis=logits.ndim - 1))


class Dirichlet(_VectorDatum):
    """x[K] on the simplex with concentration ``alpha[K]``."""

    parameter_types = dict(concentration=(floatX, 1))

    def sufficient_statistics(self, data):
        return (A.log(A.wrap_if_literal(data)),)

    def natural_parameters(self, concentration):
        return (A.wrap_if_literal(concentration) + (-1),)

    def log_normalizer(self, concentration, data_shape=None):
        alpha = A.wrap_if_literal(concentration)
        last = alpha.ndim - 1
        return A.sum(A.lgamma(alpha), axis=last) - A.lgamma(A.sum(alpha, axis=last))


def _log_multigamma(a, dim):
    """log Gamma_D(a) = D (D - 1)/4 log pi + sum_{i < D} lgamma(a - i/2) for a scalar (or per-copy
    vector) expression ``a`` and a static dimension ``dim``."""
    a = A.wrap_if_literal(a)
    const = 0.25 * dim * (dim - 1) * math.log(math.pi)
    if a.ndim == 0:
        shifted = a + A.constant(-0.5 * np.arange(dim, dtype=np.float32))          # [D]
        return A.sum(A.lgamma(shifted)) + const
    lead = a.ndim
    offsets = A.constant(-0.5 * np.arange(dim, dtype=np.float32))
    shifted = A.dimshuffle(a, *(list(range(lead)) + ['x'])) + A.dimshuffle(offsets, *(['x'] * lead + [0]))
    return A.sum(A.lgamma(shifted), axis=lead) + const


def _frobenius(X, Y, lead):
    """sum over the last two axes of X * Y, keeping ``lead`` leading axes."""
    idx = [('out', i) for i in range(lead)] + [('sum', 0), ('sum', 1)]
    return A.einsum([(X, idx), (Y, idx)], lead)


class Wishart(ExponentialFamily):
    """Precision matrix L[D, D] ~ W(W, nu), parameterised by ``df`` = nu and ``scale_inverse`` = W^-1.
    ``dim`` (= D) is static: the multivariate log-Gamma sums over arange(D).  Any number of leading
    axes index independent copies (K mixture components) on data and parameters alike."""

    def __init__(self, dim):
        self.dim = int(dim)

    parameter_types = dict(df=(floatX, 0), scale_inverse=(floatX, 2))
    data_type = (floatX, 2)

    def sufficient_statistics(self, data):
        data = A.wrap_if_literal(data)
        return A.logdet(data), data

    def natural_parameters(self, df, scale_inverse):
        df, scale_inverse = A.wrap_if_literal(df), A.wrap_if_literal(scale_inverse)
        return 0.5 * df + (-0.5 * (self.dim + 1)), -0.5 * scale_inverse

    def log_likelihood_interaction_term(self, data, **params):
        logdet_l, lam = self.sufficient_statistics(data)
        eta1, eta2 = self.natural_parameters(**params)
        return logdet_l * eta1 + _frobenius(lam, eta2, lam.ndim - 2)

    def log_normalizer(self, df, scale_inverse, data_shape=None):
        df, scale_inverse = A.wrap_if_literal(df), A.wrap_if_literal(scale_inverse)
        return (0.5 * self.dim * math.log(2.0)) * df - 0.5 * df * A.logdet(scale_inverse) \
            + _log_multigamma(0.5 * df, self.dim)

    def log_likelihood_data_term(self, data):
        return A.constant(0)


class GaussianWishart(ExponentialFamily):
    """Joint conjugate prior of a Gaussian's (mean mu[D], precision L[D, D]):
    N(mu | mean, (beta L)^-1) W(L | scale_inverse^-1, df).  The datum is the PAIR ``(mu, L)``."""

    def __init__(self, dim):
        self.dim = int(dim)

    parameter_types = dict(mean=(floatX, 1), beta=(floatX, 0), df=(floatX, 0), scale_inverse=(floatX, 2))
    data_type = ((floatX, 1), (floatX, 2))

    def sufficient_statistics(self, data):
        mu, lam = (A.wrap_if_literal(v) for v in data)
        lead = mu.ndim - 1
        batch = [('out', i) for i in range(lead)]
        l_mu = A.einsum([(lam, batch + [('out', lead), ('sum', 0)]), (mu, batch + [('sum', 0)])], lead + 1)
        quad = A.einsum([(mu, batch + [('sum', 0)]), (lam, batch + [('sum', 0), ('sum', 1)]),
                         (mu, batch + [('sum', 1)])], lead)
        return l_mu, quad, lam, A.logdet(lam)

    def natural_parameters(self, mean, beta, df, scale_inverse):
        mean, beta, df, scale_inverse = (A.wrap_if_literal(v) for v in (mean, beta, df, scale_inverse))
        lead = mean.ndim - 1
        batch = [('out', i) for i in range(lead)]
        scalar = batch if beta.ndim else []
        beta_m = A.einsum([(beta, scalar), (mean, batch + [('out', lead)])], lead + 1)
        beta_mm = A.einsum([(beta, scalar), (mean, batch + [('out', lead)]), (mean, batch + [('out', lead + 1)])],
                           lead + 2)
        return beta_m, -0.5 * beta, -0.5 * (scale_inverse + beta_mm), 0.5 * df + (-0.5 * self.dim)

    def log_likelihood_interaction_term(self, data, **params):
        l_mu, quad, lam, logdet_l = self.sufficient_statistics(data)
        eta1, eta2, eta3, eta4 = self.natural_parameters(**params)
        lead = quad.ndim
        vec = [('out', i) for i in range(lead)] + [('sum', 0)]
        return A.einsum([(l_mu, vec), (eta1, vec)], lead) + quad * eta2 + _frobenius(lam, eta3, lead) \
            + logdet_l * eta4

    def log_normalizer(self, mean, beta, df, scale_inverse, data_shape=None):
        beta, df, scale_inverse = (A.wrap_if_literal(v) for v in (beta, df, scale_inverse))
        d = self.dim
        return 0.5 * d * math.log(2.0 * math.pi) - 0.5 * d * A.log(beta) + (0.5 * d * math.log(2.0)) * df \
            - 0.5 * df * A.logdet(scale_inverse) + _log_multigamma(0.5 * df, d)

    def log_likelihood_data_term(self, data):
        return A.constant(0)

    def log_likelihood(self, data, **params):
        return self.log_likelihood_data_term(data) + self.log_likelihood_interaction_term(data, **params) \
            - self.log_normalizer(**params)
