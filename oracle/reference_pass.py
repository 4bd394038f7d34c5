"""The cfg2 pass -- {N, sum x, sum x x^T} and the Gaussian expected log-likelihood -- evaluated by the
UNMODIFIED reference: expressions built with the reference's own ``bayesic.algebra`` vocabulary,
compiled with its ``Expression.compile()`` and called as ``f(**inputs)`` (``bayesic/algebra.py:50-58``).
Arithmetic below ``compile()`` is the numpy stand-in for Theano (``oracle/theano_shim``: tensordot ->
``numpy.tensordot`` -> BLAS sgemm, sum -> ``ndarray.sum``), i.e. what Theano's CPU backend does.

TEST INFRASTRUCTURE: the ``--impl reference`` arm and the ``cpu_baseline`` leg of ``bench.py``.

The reference compiles ONE output per ``compile()`` and cannot share a pass between outputs, so the
pass is three compiled functions: ``dot(X.T, X)`` (plan ``_tensordot(_dimshuffle(X,1,0), X, [1],[0])``),
``sum(X, axis=0)`` (plan ``_sum(X, 0)``), and the log-likelihood from the statistics,
``-0.5 trace(dot(L, S2)) + dot(S1, eta) + N c`` -- the cheapest way the reference's vocabulary offers
(X is read twice, not three times)."""
import numpy as np

from .reference_loader import load_reference_algebra


class ReferenceGaussianPass(object):
    def __init__(self, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet):
        A = load_reference_algebra()
        self.algebra = A
        d = int(np.asarray(e_lambda).shape[0])
        X = A.var('X', 2)
        self.f_s2 = A.dot(X.T, X).compile()
        self.f_s1 = A.sum(X, axis=0).compile()
        S2, S1, L, eta, n = A.var('S2', 2, 'float64'), A.var('S1', 1, 'float64'), A.var('L', 2, 'float64'), \
            A.var('eta', 1, 'float64'), A.var('n', 0, 'float64')
        const = -0.5 * d * np.log(2.0 * np.pi) + 0.5 * float(e_logdet) - 0.5 * float(e_mu_l_mu)
        self.f_ll = (-0.5 * A.trace(A.dot(L, S2)) + A.dot(S1, eta) + n * const).compile()
        self.e_lambda = np.asarray(e_lambda, dtype=np.float64)
        self.e_lambda_mu = np.asarray(e_lambda_mu, dtype=np.float64)

    def __call__(self, X):
        """numpy float32 [n, d] in -> (n, sum_x, sum_xxT, loglik); every array op goes through the
        reference's compiled functions."""
        s2 = self.f_s2(X=X)
        s1 = self.f_s1(X=X)
        n = float(X.shape[0])
        ll = self.f_ll(S2=s2.astype(np.float64), S1=s1.astype(np.float64), L=self.e_lambda, eta=self.e_lambda_mu, n=n)
        return n, s1, s2, float(ll)
