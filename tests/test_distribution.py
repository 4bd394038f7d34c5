"""The exponential-family layer (bayesic_b200/distribution) against scipy.stats, evaluated with
the float64 oracle on CPU and through the CUDA executor on the GPU."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal, norm

import bayesic_b200.algebra as A
from bayesic_b200.algebra import _sum, _tensordot, _dimshuffle
from bayesic_b200.distribution import (MultivariateNormal, Normal, ExponentialFamily,
                                       ExpFamIndependentObservations, IndependentObservations)
from oracle.semantics import evaluate

X, mu, Lam, ld = A.var('X', 2), A.var('mu', 1), A.var('Lam', 2), A.var('ld', 0)


def _inputs(seed=0, n=60, d=5):
    rng = np.random.RandomState(seed)
    a = rng.randn(d, d)
    prec = a @ a.T / d + np.eye(d)
    return {'X': (rng.randn(n, d) + 0.2).astype(np.float32), 'mu': rng.randn(d).astype(np.float32),
            'Lam': prec.astype(np.float32), 'ld': np.float32(np.linalg.slogdet(prec.astype(np.float32).astype('f8'))[1])}


def test_iid_wrapper_types():
    iid = MultivariateNormal().iid()
    assert isinstance(iid, ExpFamIndependentObservations) and isinstance(iid, IndependentObservations)
    assert iid.data_type == ('float32', 2)
    assert iid.parameter_types['precision'] == ('float32', 2)
    copies = MultivariateNormal().independent_observations(param_copy_ndim=1)
    assert copies.parameter_types['mean'] == ('float32', 2) and copies.data_type == ('float32', 2)
    assert not MultivariateNormal().is_discrete()
    assert isinstance(Normal(), ExponentialFamily)


def test_iid_statistics_are_the_data_axis_contractions():
    s1, s2 = MultivariateNormal().iid().sufficient_statistics(X)
    assert repr(s1) == 'einsum(out_u = sum_i X_iu)'
    assert repr(s2) == 'einsum(out_uv = sum_i X_iu X_iv)'
    assert s1._rewrite_as_special_case_ops() == _sum(X, 0)
    assert s2._rewrite_as_special_case_ops() == _tensordot(_dimshuffle(X, 1, 0), X, [1], [0])
    assert s2 == A.dot(X.T, X)                 # same canonical einsum as the hand-written form


def test_log_likelihood_takes_the_sufficient_statistic_form():
    ll = MultivariateNormal().iid().log_likelihood(X, mean=mu, precision=Lam, log_det_precision=ld)
    text = repr(ll.lower())
    assert '_tensordot(_sum(X, 0), _tensordot(Lam, mu, [1], [0]), [0], [0])' in text
    assert '_tensordot(_tensordot(_dimshuffle(X, 1, 0), X, [1], [0]), Lam, [0, 1], [0, 1])' in text


def test_mvn_iid_log_likelihood_value():
    inp = _inputs()
    ll = MultivariateNormal().iid().log_likelihood(X, mean=mu, precision=Lam, log_det_precision=ld)
    got = float(evaluate(ll, inp))
    cov = np.linalg.inv(inp['Lam'].astype('f8'))
    want = multivariate_normal(inp['mu'].astype('f8'), cov).logpdf(inp['X'].astype('f8')).sum()
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_per_point_log_likelihood_without_the_wrapper():
    inp = _inputs(1, n=1)
    x1 = A.var('x1', 1)
    ll = MultivariateNormal().log_likelihood(x1, mean=mu, precision=Lam, log_det_precision=ld)
    got = float(evaluate(ll, dict(inp, x1=inp['X'][0])))
    cov = np.linalg.inv(inp['Lam'].astype('f8'))
    want = multivariate_normal(inp['mu'].astype('f8'), cov).logpdf(inp['X'][0].astype('f8'))
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_normal_iid_log_likelihood_value_and_fixed_normaliser():
    rng = np.random.RandomState(3)
    xv = rng.randn(40).astype(np.float32)
    x, m, v = A.var('x', 1), A.var('m', 0), A.var('v', 0)
    ll = Normal().iid().log_likelihood(x, mean=m, variance=v)
    got = float(evaluate(ll, {'x': xv, 'm': 0.3, 'v': 1.7}))
    np.testing.assert_allclose(got, norm(0.3, np.sqrt(1.7)).logpdf(xv.astype('f8')).sum(), rtol=1e-9)
    # the per-draw normaliser integrates exp(interaction): 1/2 log(2 pi var) + mu^2 / (2 var)
    ln = float(evaluate(Normal().log_normalizer(m, v), {'m': 0.3, 'v': 1.7}))
    np.testing.assert_allclose(ln, 0.5 * np.log(2 * np.pi * 1.7) + 0.3 ** 2 / (2 * 1.7), rtol=1e-12)


def test_parameter_copies_broadcast_over_iid_draws():
    # 3 copies of the parameters x 7 iid draws each
    rng = np.random.RandomState(5)
    dist = Normal().independent_observations(param_copy_ndim=1, iid_draw_ndim=1)
    assert dist.data_type == ('float32', 2) and dist.parameter_types['mean'] == ('float32', 1)
    data, m, v = A.var('data', 2), A.var('m', 1), A.var('v', 1)
    ll = dist.log_likelihood(data, mean=m, variance=v)
    dv, mv, vv = rng.randn(3, 7), rng.randn(3), rng.rand(3) + 0.5
    got = float(evaluate(ll, {'data': dv, 'm': mv, 'v': vv}))
    want = sum(norm(mv[c], np.sqrt(vv[c])).logpdf(dv[c]).sum() for c in range(3))
    np.testing.assert_allclose(got, want, rtol=1e-9)


@pytest.mark.gpu
def test_gpu_log_likelihood_and_statistics():
    inp = _inputs(7, n=5000, d=16)
    iid = MultivariateNormal().iid()
    ll = iid.log_likelihood(X, mean=mu, precision=Lam, log_det_precision=ld)
    got = float(ll.compile()(**inp))
    cov = np.linalg.inv(inp['Lam'].astype('f8'))
    want = multivariate_normal(inp['mu'].astype('f8'), cov).logpdf(inp['X'].astype('f8')).sum()
    assert abs(got - want) <= 1e-4 * abs(want)
    fn = iid.compile_sufficient_statistics(X)
    s1, s2 = fn(X=inp['X'])
    Xd = inp['X'].astype('f8')
    np.testing.assert_allclose(s1, Xd.sum(0), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(s2, Xd.T @ Xd, rtol=1e-4, atol=1e-3)
    assert 21 in [n['kind'] for n in fn.plan.lowered.nodes]     # served by the SYRK kernel
