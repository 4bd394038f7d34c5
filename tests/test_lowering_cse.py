"""Sub-tree sharing in the lowering must be STRICT.  The expression classes' ``==`` follows the
reference (``bayesic/algebra.py:1297-1309``: ``_mul`` / ``add`` compare parents as frozensets;
``_tensordot`` ignores batch-axis order), which the reference never uses for evaluation
(``algebra.py:34-40`` applies every node).  Merging by that equality computed 2*X*X*Y for
X*X*Y + X*Y (round-1 advisor finding); merging is now by identity / value numbering."""
import numpy as np
import pytest

import bayesic_b200.algebra as A
from bayesic_b200.algebra.plan_ir import _tensordot
from bayesic_b200.backend.lowering import lower_plans
from oracle.descriptor_eval import evaluate_descriptor

RNG = np.random.RandomState(7)
XV, YV = RNG.randn(3, 4), RNG.randn(3, 4)


def _eval(exprs):
    types = {}
    for e in exprs:
        types.update(e.input_types)
    low = lower_plans([e.lower() for e in exprs], types)
    arrays = [{'X': XV, 'Y': YV}[n] for n in low.input_names]
    return low, evaluate_descriptor(low.nodes, low.outputs, arrays)


def test_repeated_factors_are_not_merged_with_single_ones():
    X, Y = A.var('X', 2), A.var('Y', 2)
    low, (val,) = _eval([X * X * Y + X * Y])
    np.testing.assert_allclose(val, XV * XV * YV + XV * YV, rtol=1e-12)
    muls = [n for n in low.nodes if n['kind'] == 5]
    assert sorted(len(n['parents']) for n in muls) == [2, 3]


def test_repeated_terms_across_outputs_stay_distinct():
    X, Y = A.var('X', 2), A.var('Y', 2)
    low, (a, b) = _eval([X + X + Y, X + Y])
    assert low.outputs[0] != low.outputs[1]
    np.testing.assert_allclose(a, 2 * XV + YV, rtol=1e-12)
    np.testing.assert_allclose(b, XV + YV, rtol=1e-12)


def test_polynomial():
    X = A.var('X', 2)
    low, (val,) = _eval([X ** 3 + X ** 2 + X * X * X])
    np.testing.assert_allclose(val, 2 * XV ** 3 + XV ** 2, rtol=1e-12)


def test_identical_subtrees_still_share_one_node():
    X = A.var('X', 2)
    low, (a, b) = _eval([A.exp(A.dot(X.T, X)), A.abs_(A.dot(X.T, X))])   # two separate objects
    assert sum(1 for n in low.nodes if n['kind'] in (7, 21)) == 1


def test_batch_axis_order_is_part_of_the_key():
    P, Q = A.var('P', 3), A.var('Q', 3)
    t1 = _tensordot(P, Q, [2], [2], [0, 1], [0, 1])
    t2 = _tensordot(P, Q, [2], [2], [1, 0], [1, 0])
    assert t1 == t2                      # reference equality: batch pairs as a set
    low = lower_plans([t1, t2], {'P': ('float32', 3), 'Q': ('float32', 3)})
    assert low.outputs[0] != low.outputs[1]
    p, q = RNG.randn(2, 3, 5), RNG.randn(2, 3, 5)
    arrays = [{'P': p, 'Q': q}[n] for n in low.input_names]
    a, b = evaluate_descriptor(low.nodes, low.outputs, arrays)
    np.testing.assert_allclose(a, np.einsum('abk,abk->ab', p, q), rtol=1e-12)
    np.testing.assert_allclose(b, np.einsum('abk,abk->ba', p, q), rtol=1e-12)


def test_host_call_signature_is_shapes_and_scalars():
    """``CompiledPlan._host_signature`` (the key of the persistent-buffer / CUDA-graph cache of numpy-in calls):
    array inputs contribute their shapes, rank-0 inputs their values; a wrong rank or integers beyond 2**24 give
    ``None`` so that the general path reports them."""
    import numpy as np
    import bayesic_b200.algebra as A
    from bayesic_b200.backend.compiled import CompiledPlan
    X, s = A.var('X', 2), A.var('s', 0)
    plan = CompiledPlan([A.sum(X * s, axis=0)])
    a = plan._host_signature({'X': np.zeros((5, 3), np.float32), 's': 2.0})
    b = plan._host_signature({'X': np.ones((5, 3), np.float32), 's': 2.0})
    assert a == b and dict(a)['X'] == (5, 3) and dict(a)['s'] == 2.0
    assert plan._host_signature({'X': np.zeros((6, 3), np.float32), 's': 2.0}) != a
    assert plan._host_signature({'X': np.zeros((5, 3), np.float32), 's': 3.0}) != a
    assert plan._host_signature({'X': np.zeros(5, np.float32), 's': 2.0}) is None
    counts = A.var('c', 1, dtype='int64')
    iplan = CompiledPlan([A.sum(counts, axis=0)])
    assert iplan._host_signature({'c': np.array([1, 2, 3])}) is not None
    assert iplan._host_signature({'c': np.array([1, (1 << 24) + 1])}) is None


def test_split_responsibility_shapes():
    import bayesic_b200.stats as S
    assert S.split_responsibilities_supported(64, 256) and S.split_responsibilities_supported(16, 1024)
    assert not S.split_responsibilities_supported(64, 128) and not S.split_responsibilities_supported(12, 256)
    assert not S.split_responsibilities_supported(64, 260)
