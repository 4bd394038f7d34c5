"""Structural behaviour of the host-side algebra: canonical forms, equality,
``match``, multiset injections, identity handling, error behaviour.  The expected
values are the ones the reference's own tests state
(``bayesic/tests/test_algebra.py:194-358``) plus cases for the defects we fix."""
from collections import Counter

import numpy as np
import pytest

from bayesic_b200.algebra import *  # noqa: F401,F403
from bayesic_b200.algebra import _sum, _mul, _dimshuffle, _tensordot, _diagonal

X, Y, Z, W = (var(n, 2) for n in 'XYZW')
x, y = var('x', 1), var('y', 1)
S = var('S', 3)


def test_equivalent_expressions_share_one_canonical_form():
    a, b = trace(dot(X.T, Y)), sum(X * Y)
    assert a.parents == (X, Y) and b.parents == (X, Y)
    assert a.factors_and_indices == b.factors_and_indices


def test_composition_collapses_to_one_einsum():
    expr = dot(diagonal(dot(X, outer(x, y))), Y)
    assert expr.parents == (X, x, y, Y)


@pytest.mark.parametrize('A_, B_', [([1], []), ([1, 1], [1]), ([1, 1, 2], [1, 2])])
def test_no_injection_exists(A_, B_):
    assert list(find_injections(A_, B_)) == []


@pytest.mark.parametrize('A_, B_, expected', [
    ([], [], []),
    ([1], [1], [(1, 1)]),
    ([1], [1, 2], [(1, 1)]),
    ([1, 1], [1, 1], [(1, 1), (1, 1)]),
    ([1, 1], [1, 1, 2], [(1, 1), (1, 1)]),
    ([1, 2], [5, 2, 1, 3], [(1, 1), (2, 2)]),
])
def test_unique_injection_under_equality(A_, B_, expected):
    assert list(find_injections(A_, B_)) == [Counter(expected)]


def _second_char(a, b):
    return a[1] == b[1]


@pytest.mark.parametrize('A_, B_, expected', [
    (["a1", "a1"], ["b1"], []),
    (["a1", "a1"], ["b1", "b2"], []),
    (["a1"], ["b1"], [[("a1", "b1")]]),
    (["a1", "a3"], ["b3", "b1"], [[("a1", "b1"), ("a3", "b3")]]),
    (["a1", "a1"], ["a2", "b1", "c1"], [[("a1", "b1"), ("a1", "c1")]]),
    (["a1", "b1"], ["c1", "c1", "d2"], [[("a1", "c1"), ("b1", "c1")]]),
    (["a1", "b1"], ["x1", "y1", "z2"],
     [[("a1", "x1"), ("b1", "y1")], [("a1", "y1"), ("b1", "x1")]]),
    (["a1", "a1", "a1", "b2"], ["x1", "x1", "x1", "y1", "y1", "z2", "extra"],
     [{("b2", "z2"): 1, ("a1", "x1"): 3},
      {("b2", "z2"): 1, ("a1", "x1"): 2, ("a1", "y1"): 1},
      {("b2", "z2"): 1, ("a1", "x1"): 1, ("a1", "y1"): 2}]),
])
def test_injections_under_custom_match(A_, B_, expected):
    got = Counter(frozenset(c.items()) for c in find_injections(A_, B_, _second_char))
    want = Counter(frozenset(Counter(i).items()) for i in expected)
    assert got == want


def test_injection_counts_for_subsets_and_partitions():
    pairs = list(find_injections(["a1", "a1"], ["a1", "b1", "c1", "d1", "e1"], _second_char))
    assert len(pairs) == 10          # all 2-subsets of 5
    parts = list(find_injections(["a1", "a1", "b1", "b1"], ["w1", "x1", "y1", "z1"], _second_char))
    assert len(parts) == 6           # all 2+2 partitions of 4


def test_equality_laws():
    assert X == X and X != Y
    assert constant(1) == constant(1) and constant(1) != constant(2)
    assert X + Y == Y + X and X + Y != X + Z
    assert X - Y == -Y + X
    assert X / Y == X * (Y ** -1)
    assert log(X) == log(X) and log(X) != exp(X) and log(X) != log(Y)
    assert X * Y == Y * X
    assert X * X.T == X.T * X and X * X.T != X * X
    assert dot(X, Y).T == dot(Y.T, X.T)
    assert dot(X, Y) != dot(Y, X)
    assert sum(X * X.T) == sum(X.T * X)
    assert sum(X * X.T) != trace(X) * trace(X)
    assert trace(dot(X, Y.T)) == sum(Y * X)
    assert dot(dot(X, Y), Z) == dot(X, dot(Y, Z))
    assert hash(dot(dot(X, Y), Z)) == hash(dot(X, dot(Y, Z)))
    assert hash(X * Y) == hash(Y * X)


def test_equality_against_scalar_leftover_is_false_not_an_error():
    # the reference raises AttributeError here (match returns a bare var)
    c = var('c', 0)
    assert (X * c) != X


def test_match_extracts_the_slot():
    assert match(X * Y, X * Z, Z) == Y
    assert match(X * X, X * Z, Z) == X
    assert match(X * X, Y * Z, Z) is None
    assert match(Y * X, X * Z, Z) == Y
    assert match(sum(Y * X), sum(X * Z), Z) == Y
    assert match(dot(X, Y), dot(X, Z), Z) == Y
    assert match(dot(X, X), dot(X, Z), Z) == X
    assert match(dot(X, X.T), dot(X, Z), Z) == X.T
    assert match(dot(X, X.T), dot(X.T, Z), Z) is None
    assert match(dot(X, X * X), dot(X, Z), Z) == X * X
    assert match(dot(X, X * X), dot(Z, X * X), Z) == X
    assert match(dot(X, X * X), dot(X * X, Z), Z) is None
    assert match(dot(X, X * X), dot(X, X * Z), Z) == X
    assert match(trace(dot(X, X)), sum(X * Z), Z) == X.T
    assert match(dot(X, Y).T, dot(X, Z), Z) is None
    assert match(dot(X, Y).T, dot(X.T, Z), Z) is None
    assert match(dot(X, Y).T, dot(Z, X.T), Z) == Y.T
    assert match(dot(X, dot(Y, X)), dot(X, Z), Z) == dot(Y, X)
    assert match(X, Z, Z) == X
    assert match(X * Y, Z, Z) == X * Y


def test_match_pulls_the_statistic_paired_with_a_natural_parameter():
    L = var('L', 2)
    got = match(-0.5 * trace(dot(L, dot(X.T, X))), sum(L * Z), Z)
    assert repr(got) == 'einsum(out_uv = sum_i -0.5 X_iv X_iu)'
    assert got == -0.5 * dot(X.T, X).T


def test_match_errors():
    with pytest.raises(ValueError):
        match(X, Y * W, Z)                       # slot not in template
    with pytest.raises(ValueError):
        match(trace(X), trace(Z), Z)             # slot index used twice


def test_identity_elimination_and_insertion():
    assert dot(X, eye(X.shape[1])) == X
    assert dot(eye(X.shape[0]), X) == X
    assert dot(Y, dot(eye(X.shape[0]), X)) == dot(Y, X)
    assert match(X * Y, dot(X * Y, Z), Z) == eye(X.shape[1])
    assert match(X, dot(X, Z), Z) == eye(X.shape[1])
    assert match(X * y.dimshuffle('x', 0), dot(X, Z), Z) == eye(X.shape[1]) * y.dimshuffle(0, 'x')


def test_construction_errors_match_the_reference():
    with pytest.raises(ValueError):
        dot(X, object())                         # algebra.py:170
    with pytest.raises(ValueError):
        X + x                                    # rank mismatch, algebra.py:190-192
    with pytest.raises(ValueError):
        eye()                                    # algebra.py:254
    with pytest.raises(ValueError):
        einsum([(X, [('out', 0)])])              # algebra.py:363
    with pytest.raises(ValueError):
        einsum([(X, [('out', 0), ('out', 5)])], ndim=2)     # algebra.py:373
    with pytest.raises(ValueError):
        dimshuffle(X, 0, 0)                      # algebra.py:1272
    with pytest.raises(ValueError):
        dimshuffle(X, 0)                         # algebra.py:1275
    with pytest.raises(TypeError):
        (var('q', 2) + var('q', 1).dimshuffle(0, 'x')).input_types   # algebra.py:28


def test_input_types_and_defaults():
    a = var('a', 0, 'int32')
    expr = dot(X, y) * a
    assert expr.input_types == {'X': ('float32', 2), 'y': ('float32', 1), 'a': ('int32', 0)}
    assert var('v', 3).dtype == 'float32'        # algebra.py:109


def test_repeated_child_einsum_keeps_independent_contractions():
    # reference defect (algebra.py:411 keys the renaming by equal factor): both
    # dot(X, Y) occurrences would share one contracted index.
    expr = dot(X, Y) * dot(X, Y)
    assert len(expr.sum_indices) == 2
    from oracle.semantics import evaluate
    rng = np.random.RandomState(0)
    Xv, Yv = rng.randn(4, 4), rng.randn(4, 4)
    np.testing.assert_allclose(evaluate(expr, {'X': Xv, 'Y': Yv}), (Xv @ Yv) ** 2, rtol=1e-12)


def test_basic_plans():
    def planned(e):
        return e._rewrite_as_special_case_ops()
    assert planned(diagonal(X)) == _diagonal(X, 0, 1)
    assert planned(dot(X, Y)) == _tensordot(X, Y, [1], [0])
    assert planned(sum(X, 1)) == _sum(X, 1)
    assert planned(mul(X, Y)) == _mul(X, Y)
    assert planned(dimshuffle(X, 1, 0)) == _dimshuffle(X, 1, 0)
    assert planned(trace(dot(X.T, Y))) == _tensordot(X, Y, [0, 1], [0, 1])
    assert planned((X * Y.T).sum(axis=1)) == _tensordot(X, Y, [1], [0], [0], [1])
    assert planned(dot(Z, x * y)) == _tensordot(Z, _mul(x, y), [1], [0])
    assert planned(dot(X * Y, Z * W)) == _tensordot(_mul(X, Y), _mul(Z, W), [1], [0])
    assert planned(X.sum(1) * y) == _mul(_sum(X, 1), y)
