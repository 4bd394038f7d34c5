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


# ---- the families of distribution/conjugate.py against scipy.stats ----

def _conjugate_cases(seed=3):
    """(name, log-likelihood expression, inputs, scipy value) for every added family, iid over axis 0."""
    from scipy import stats as st
    from bayesic_b200.distribution import BernoulliLogit, Exponential, Gamma, Categorical, Dirichlet
    rng = np.random.RandomState(seed)
    x, M = A.var('x', 1), A.var('M', 2)
    k, n = 6, 400
    cases = []
    xv = (rng.rand(n) < 0.3).astype(np.float32)
    cases.append(('bernoulli', BernoulliLogit().iid().log_likelihood(x, logit=A.var('l', 0)),
                  {'x': xv, 'l': np.float32(-0.4)}, st.bernoulli(1 / (1 + np.exp(0.4))).logpmf(xv).sum()))
    xv = rng.exponential(0.5, n).astype(np.float32)
    cases.append(('exponential', Exponential().iid().log_likelihood(x, rate=A.var('r', 0)),
                  {'x': xv, 'r': np.float32(2.0)}, st.expon(scale=0.5).logpdf(xv.astype('f8')).sum()))
    xv = rng.gamma(3.0, 1 / 1.5, n).astype(np.float32)
    cases.append(('gamma', Gamma().iid().log_likelihood(x, shape=A.var('a', 0), rate=A.var('b', 0)),
                  {'x': xv, 'a': np.float32(3.0), 'b': np.float32(1.5)},
                  st.gamma(3.0, scale=1 / 1.5).logpdf(xv.astype('f8')).sum()))
    idx = rng.randint(k, size=n)
    lg = rng.randn(k).astype(np.float32)
    log_p = lg.astype('f8') - np.log(np.exp(lg.astype('f8')).sum())
    cases.append(('categorical', Categorical().iid().log_likelihood(M, logits=A.var('lg', 1)),
                  {'M': np.eye(k, dtype=np.float32)[idx], 'lg': lg}, log_p[idx].sum()))
    alpha = (rng.rand(k) * 3 + 0.5).astype(np.float32)
    Mv = rng.dirichlet(alpha, size=n).astype(np.float32)
    want = sum(st.dirichlet(alpha.astype('f8')).logpdf(r.astype('f8') / r.astype('f8').sum()) for r in Mv)
    cases.append(('dirichlet', Dirichlet().iid().log_likelihood(M, concentration=A.var('al', 1)),
                  {'M': Mv, 'al': alpha}, want))
    return cases


def test_conjugate_families_match_scipy():
    for name, ll, inputs, want in _conjugate_cases():
        got = float(evaluate(ll, inputs))
        np.testing.assert_allclose(got, want, rtol=1e-6, err_msg=name)


def test_conjugate_families_structure():
    from bayesic_b200.distribution import BernoulliLogit, Gamma, Categorical, Dirichlet
    x, M = A.var('x', 1), A.var('M', 2)
    # iid statistics are the data-axis sums the device kernels serve
    s_log, s_x = Gamma().iid().sufficient_statistics(x)
    assert s_x._rewrite_as_special_case_ops() == _sum(x, 0)
    assert 'log' in repr(s_log)
    (counts,) = Categorical().iid().sufficient_statistics(M)
    assert counts._rewrite_as_special_case_ops() == _sum(M, 0)          # class counts N_k = sum_n r_nk
    assert BernoulliLogit().is_discrete() and Categorical().is_discrete() and not Dirichlet().is_discrete()
    assert Dirichlet().iid().data_type == ('float32', 2) and Gamma().iid().data_type == ('float32', 1)
    # per-copy parameters: K Gamma copies, each with its own (shape, rate), iid draws along axis 1
    rng = np.random.RandomState(2)
    data = rng.gamma(2.0, 1.0, size=(3, 50)).astype(np.float32)
    a, b = np.array([1.5, 2.0, 3.0], np.float32), np.array([0.5, 1.0, 2.0], np.float32)
    copies = Gamma().independent_observations(param_copy_ndim=1, iid_draw_ndim=1)
    ll = copies.log_likelihood(A.var('data', 2), shape=A.var('a', 1), rate=A.var('b', 1))
    from scipy import stats as st
    want = sum(st.gamma(a[i], scale=1 / b[i]).logpdf(data[i].astype('f8')).sum() for i in range(3))
    np.testing.assert_allclose(float(evaluate(ll, {'data': data, 'a': a, 'b': b})), want, rtol=1e-6)


@pytest.mark.gpu
def test_gpu_conjugate_families():
    """The same log-likelihoods through compile() -> C-ABI executor (float32 device arithmetic,
    incl. the lgamma opcode), rtol 1e-4."""
    for name, ll, inputs, want in _conjugate_cases(seed=5):
        got = float(ll.compile()(**inputs))
        assert abs(got - want) <= 1e-4 * abs(want) + 1e-3, (name, got, want)


def test_match_recovers_the_statistic_paired_with_a_natural_parameter():
    """The statistic a natural parameter is paired with can be read off the interaction term with
    ``match`` (algebra.py:1037-1063) -- the reference's route from a log-likelihood expression to the
    contraction over the data axis that the device kernels serve."""
    from bayesic_b200.algebra import match
    from bayesic_b200.distribution import Categorical
    eta2, slot2 = A.var('eta2', 2), A.var('slot2', 2)
    s1, s2 = MultivariateNormal().iid().sufficient_statistics(X)
    pair = [('sum', 0), ('sum', 1)]
    term = A.einsum([(s2, pair), (eta2, pair)], 0)                       # <sum_n x x^T, eta2>
    found = match(term, A.einsum([(slot2, pair), (eta2, pair)], 0), slot2)
    assert found == s2 == A.dot(X.T, X)
    M, lg, slot1 = A.var('M', 2), A.var('lg', 1), A.var('slot1', 1)
    (counts,) = Categorical().iid().sufficient_statistics(M)
    term = A.einsum([(counts, [('sum', 0)]), (lg, [('sum', 0)])], 0)     # <class counts, logits>
    found = match(term, A.einsum([(slot1, [('sum', 0)]), (lg, [('sum', 0)])], 0), slot1)
    assert found == counts and found._rewrite_as_special_case_ops() == _sum(M, 0)
    assert match(term, A.einsum([(slot1, [('sum', 0)]), (A.var('other', 1), [('sum', 0)])], 0), slot1) is None


def test_conjugate_families_lower_to_a_faithful_descriptor():
    """CPU check of the lowering: the flat plan descriptor the C-ABI executor would be given (incl. the
    lgamma opcode), evaluated node by node in numpy, equals the expression-level value."""
    from bayesic_b200.backend.compiled import CompiledPlan
    from oracle.descriptor_eval import evaluate_descriptor
    for name, ll, inputs, want in _conjugate_cases(seed=9):
        low = CompiledPlan([ll]).lowered
        arrays = [inputs[n] if n else low.bound_constants[i] for i, n in enumerate(low.input_names)]
        (value,) = evaluate_descriptor(low.nodes, low.outputs, arrays)
        np.testing.assert_allclose(float(value), want, rtol=1e-6, err_msg=name)
        np.testing.assert_allclose(float(value), float(evaluate(ll, inputs)), rtol=1e-10, err_msg=name)
