"""Core distributions (``bayesic/distribution/core.py`` of the reference, which does not
import: it names modules that do not exist, :3-4, and uses undefined ``np``/``mean``/
``precision``, :23, :44-52).

Parametrisations kept from the reference:
  Normal(mean, variance):            s = (x, x^2),    eta = (mu/var, -1/2 / var)       core.py:16-20
  MultivariateNormal(mean, precision): s = (x, x x^T), eta = (Lambda mu, -1/2 Lambda)  core.py:41-47

Log-normalisers are the mathematically correct ones (the reference's ``Normal.log_normalizer``,
core.py:22-25, has the wrong sign on the 2 pi term and squares ``mean/variance``):
  Normal:  1/2 log 2pi + 1/2 log var + 1/2 mu^2 / var
  MVN:     1/2 D log 2pi - 1/2 log|Lambda| + 1/2 mu^T Lambda mu      (core.py:49-52 intent)
``log|Lambda|``: the reference calls a ``T.logdet`` that Theano never had (core.py:51); here it is the
``logdet`` node of the algebra (``BB_NODE_LOGDET``, float64 Cholesky on the device), so the normaliser
is an expression of (mean, precision) as core.py:49-52 intends.  ``log_det_precision`` may still be
passed explicitly -- in VMP it is E[log|Lambda|], which is not the log-determinant of E[Lambda].
"""
import numpy as np

from .. import algebra as A
from .base import ExponentialFamily

floatX = 'float32'
_LOG_2PI = float(np.log(2.0 * np.pi))

__all__ = ['Normal', 'MultivariateNormal']


class Normal(ExponentialFamily):
    """Univariate Gaussian in terms of mean and variance."""

    parameter_types = dict(mean=(floatX, 0), variance=(floatX, 0))
    data_type = (floatX, 0)

    def sufficient_statistics(self, data):
        return data, data ** 2

    def natural_parameters(self, mean, variance):
        mean, variance = A.wrap_if_literal(mean), A.wrap_if_literal(variance)
        return mean / variance, -0.5 / variance

    def log_normalizer(self, mean, variance, data_shape=None):
        mean, variance = A.wrap_if_literal(mean), A.wrap_if_literal(variance)
        return 0.5 * _LOG_2PI + 0.5 * A.log(variance) + 0.5 * (mean * mean) / variance

    def log_likelihood_data_term(self, data):
        return A.constant(0)

    def _pair(self, stat, natural):
        stat, natural = A.wrap_if_literal(stat), A.wrap_if_literal(natural)
        # scalar datum: the "dot product" is a plain (broadcasting) product per observation
        if natural.ndim == 0 or natural.ndim == stat.ndim:
            return stat * natural
        return ExponentialFamily._pair(self, stat, natural)


class MultivariateNormal(ExponentialFamily):
    """Multivariate Gaussian in terms of mean and precision matrix."""

    parameter_types = dict(mean=(floatX, 1), precision=(floatX, 2), log_det_precision=(floatX, 0))
    data_type = (floatX, 1)

    def sufficient_statistics(self, data):
        """(x, x x^T) with any number of leading observation axes: for data[n, d] the second
        statistic is ``einsum(out_uvw = X_uv X_uw)``; summing it over u gives the Gram
        contraction over the data axis (the reference's TODO at core.py:42-43)."""
        data = A.wrap_if_literal(data)
        lead = data.ndim - 1
        left = [('out', i) for i in range(lead)] + [('out', lead)]
        right = [('out', i) for i in range(lead)] + [('out', lead + 1)]
        return data, A.einsum([(data, left), (data, right)], lead + 2)

    def natural_parameters(self, mean, precision, log_det_precision=None):
        mean, precision = A.wrap_if_literal(mean), A.wrap_if_literal(precision)
        lead = mean.ndim - 1
        if lead == 0:
            return A.dot(precision, mean), -0.5 * precision
        # per-copy parameters: eta1[c, d] = sum_e precision[c, d, e] mean[c, e]
        batch = [('out', i) for i in range(lead)]
        eta1 = A.einsum([(precision, batch + [('out', lead), ('sum', 0)]),
                         (mean, batch + [('sum', 0)])], lead + 1)
        return eta1, -0.5 * precision

    def log_normalizer(self, mean, precision, log_det_precision=None, data_shape=None):
        mean, precision = A.wrap_if_literal(mean), A.wrap_if_literal(precision)
        if log_det_precision is None:
            log_det_precision = A.logdet(precision)
        lead = mean.ndim - 1
        batch = [('out', i) for i in range(lead)]
        quad = A.einsum([(mean, batch + [('sum', 0)]),
                         (precision, batch + [('sum', 0), ('sum', 1)]),
                         (mean, batch + [('sum', 1)])], lead)
        d = mean.shape[lead]
        return 0.5 * _LOG_2PI * d - 0.5 * A.wrap_if_literal(log_det_precision) + 0.5 * quad

    def log_likelihood_data_term(self, data):
        return A.constant(0)

    def _core_ndim(self, stat):
        return stat.ndim
