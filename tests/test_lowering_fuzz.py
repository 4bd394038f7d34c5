"""Seeded random expressions through the PRODUCT lowering (``backend.lowering.lower_plans``: planner output ->
flat plan descriptor, fused-pattern recognition, sub-tree sharing), evaluated by the test-only numpy interpreter
of that descriptor and compared with the declared einsum semantics in float64 (``oracle.semantics``, which follows
bayesic/algebra.py:314-346).  What the C executor is handed must mean what the expression means -- for shapes of
expression nobody wrote down by hand, singly and as multi-output plans (which share sub-trees)."""
import numpy as np
import pytest

import bayesic_b200.algebra as A
from bayesic_b200.backend.lowering import lower_plans
from oracle.descriptor_eval import evaluate_descriptor
from oracle.semantics import evaluate
from tests.test_reference_crosscheck import _build, _recipe, _variables

DATA = np.random.RandomState(11)
INPUTS = {'X': DATA.randn(5, 5), 'Y': DATA.randn(5, 5), 'x': DATA.randn(5), 'y': DATA.randn(5)}


def _lowered_values(exprs):
    types = {}
    for e in exprs:
        types.update(e.input_types)
    low = lower_plans([e.lower() for e in exprs], types)
    return evaluate_descriptor(low.nodes, low.outputs, [INPUTS[n] for n in low.input_names])


def _expressions(seed, count, depth):
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < count:
        e = A.wrap_if_literal(_build(A, _recipe(rng, int(rng.randint(3)), depth), _variables(A)))
        if e.input_types:                       # a recipe can fold to a constant; nothing to lower then
            out.append(e)
    return out


@pytest.mark.parametrize('seed,depth', [(101, 2), (202, 3), (303, 4)])
def test_lowered_descriptor_means_what_the_expression_means(seed, depth):
    for i, e in enumerate(_expressions(seed, 80, depth)):
        (got,) = _lowered_values([e])
        want = evaluate(e, {k: INPUTS[k] for k in e.input_types})
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=1e-9,
                                   atol=1e-9 * max(1.0, float(np.abs(want).max())), err_msg='%d: %r' % (i, e))


def test_multi_output_plans_share_sub_trees_without_changing_values():
    exprs = _expressions(404, 60, 3)
    for i in range(0, len(exprs), 3):
        group = exprs[i:i + 3]
        got = _lowered_values(group)
        for e, g in zip(group, got):
            want = evaluate(e, {k: INPUTS[k] for k in e.input_types})
            np.testing.assert_allclose(np.asarray(g, dtype=np.float64), want, rtol=1e-9,
                                       atol=1e-9 * max(1.0, float(np.abs(want).max())), err_msg=repr(e))


def _decorated(seed, count):
    """the same recipes wrapped in the pointwise vocabulary (algebra.py:195-233, 1435-1448) and reduced again"""
    rng = np.random.RandomState(seed)
    out = []
    for e in _expressions(seed + 1, count, 2):
        kind = rng.choice(['exp', 'abs', 'square', 'log1pabs', 'mix'])
        if kind == 'exp':
            e = A.exp(0.1 * e)
        elif kind == 'abs':
            e = A.abs_(e)
        elif kind == 'square':
            e = e ** 2
        elif kind == 'log1pabs':
            e = A.log(A.abs_(e) + 1.0)
        else:
            e = A.exp(0.05 * e) * A.abs_(e) + e ** 2
        if e.ndim > 0 and rng.rand() < 0.5:
            e = A.sum(e, axis=int(rng.randint(e.ndim))) if rng.rand() < 0.5 else A.sum(e)
        out.append(e)
    return out


def test_pointwise_vocabulary_on_top_of_contractions():
    for i, e in enumerate(_decorated(505, 80)):
        (got,) = _lowered_values([e])
        want = evaluate(e, {k: INPUTS[k] for k in e.input_types})
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=1e-9,
                                   atol=1e-9 * max(1.0, float(np.abs(want).max())), err_msg='%d: %r' % (i, e))
