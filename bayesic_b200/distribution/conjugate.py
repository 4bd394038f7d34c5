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
"""
from .. import algebra as A
from .base import ExponentialFamily

floatX = 'float32'

__all__ = ['BernoulliLogit', 'Exponential', 'Gamma', 'Categorical', 'Dirichlet']


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
