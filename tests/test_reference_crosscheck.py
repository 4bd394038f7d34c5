"""Seeded random expressions built twice -- with the UNMODIFIED reference ``bayesic.algebra``
(imported through the numpy Theano shim; authoring container only, skipped where
``/root/reference`` is absent) and with ``bayesic_b200.algebra`` -- must agree on the canonical
form (``repr``), on the plan the planner emits, and on the value.  Complements the fixed golden
cases (tests/golden) with shapes of expression nobody wrote down by hand.

The reference's own defects are respected, not reproduced: when its planner raises
the case only checks that ours does not; when its plan has batch axes its evaluator is known to be
wrong or to crash (algebra.py:1364-1380), so the value is checked against the declared semantics
only."""
import json

import numpy as np
import pytest

import bayesic_b200.algebra as A
from oracle.plan_dump import dump_plan
from oracle.reference_loader import load_reference_algebra, reference_available
from oracle.semantics import evaluate

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')

N_CASES = 160


def _recipe(rng, rank, depth):
    """Nested tuple describing an expression of the given rank over X, Y (5x5) and x, y (5)."""
    leaf = depth <= 0 or rng.rand() < 0.25
    if rank == 2:
        if leaf:
            return ('var', rng.choice(['X', 'Y']))
        kind = rng.choice(['T', 'dot', 'mul', 'add', 'outer', 'scale', 'sub'])
        if kind == 'T':
            return ('T', _recipe(rng, 2, depth - 1))
        if kind == 'outer':
            return ('outer', _recipe(rng, 1, depth - 1), _recipe(rng, 1, depth - 1))
        if kind == 'scale':
            return ('scale', float(rng.choice([2.0, -0.5, 3.0])), _recipe(rng, 2, depth - 1))
        return (kind, _recipe(rng, 2, depth - 1), _recipe(rng, 2, depth - 1))
    if rank == 1:
        if leaf:
            return ('var', rng.choice(['x', 'y']))
        kind = rng.choice(['matvec', 'vecmat', 'mul', 'add', 'sumax', 'diag', 'scale'])
        if kind == 'matvec':
            return ('dot', _recipe(rng, 2, depth - 1), _recipe(rng, 1, depth - 1))
        if kind == 'vecmat':
            return ('dot', _recipe(rng, 1, depth - 1), _recipe(rng, 2, depth - 1))
        if kind == 'sumax':
            return ('sumax', int(rng.randint(2)), _recipe(rng, 2, depth - 1))
        if kind == 'diag':
            return ('diag', _recipe(rng, 2, depth - 1))
        if kind == 'scale':
            return ('scale', float(rng.choice([2.0, -0.5])), _recipe(rng, 1, depth - 1))
        return (kind, _recipe(rng, 1, depth - 1), _recipe(rng, 1, depth - 1))
    kind = rng.choice(['inner', 'trace', 'sum1', 'sum2'])
    if kind == 'inner':
        return ('dot', _recipe(rng, 1, depth - 1), _recipe(rng, 1, depth - 1))
    if kind == 'trace':
        return ('trace', _recipe(rng, 2, depth - 1))
    return ('sumall', _recipe(rng, 1 if kind == 'sum1' else 2, depth - 1))


def _build(ns, recipe, variables):
    op = recipe[0]
    if op == 'var':
        return variables[recipe[1]]
    if op == 'T':
        return _build(ns, recipe[1], variables).T
    if op == 'scale':
        return recipe[1] * _build(ns, recipe[2], variables)
    if op == 'sumax':
        return ns.sum(_build(ns, recipe[2], variables), axis=recipe[1])
    if op == 'sumall':
        return ns.sum(_build(ns, recipe[1], variables))
    if op == 'diag':
        return ns.diagonal(_build(ns, recipe[1], variables))
    if op == 'trace':
        return ns.trace(_build(ns, recipe[1], variables))
    a, b = _build(ns, recipe[1], variables), _build(ns, recipe[2], variables)
    if op == 'dot':
        return ns.dot(a, b)
    if op == 'outer':
        return ns.outer(a, b)
    if op == 'mul':
        return a * b
    if op == 'add':
        return a + b
    if op == 'sub':
        return a - b
    raise ValueError(op)


def _variables(ns):
    return {'X': ns.var('X', 2), 'Y': ns.var('Y', 2), 'x': ns.var('x', 1), 'y': ns.var('y', 1)}


def _canon(plan):
    if isinstance(plan, dict):
        out = {k: _canon(v) for k, v in plan.items()}
        if out.get('op') == '_mul':
            out['factors'] = sorted(out['factors'], key=lambda f: json.dumps(f, sort_keys=True))
        return out
    if isinstance(plan, list):
        return [_canon(p) for p in plan]
    return plan


def _has_batch_axes(plan):
    if isinstance(plan, dict):
        if plan.get('op') == '_tensordot' and plan.get('x_batch'):
            return True
        return any(_has_batch_axes(v) for v in plan.values())
    if isinstance(plan, list):
        return any(_has_batch_axes(p) for p in plan)
    return False


def test_random_expressions_agree_with_the_unmodified_reference():
    ref = load_reference_algebra()
    rng = np.random.RandomState(20260)
    data = np.random.RandomState(7)
    inputs = {'X': data.randn(5, 5).astype(np.float32), 'Y': data.randn(5, 5).astype(np.float32),
              'x': data.randn(5).astype(np.float32), 'y': data.randn(5).astype(np.float32)}
    planned = valued = 0
    for case in range(N_CASES):
        recipe = _recipe(rng, int(rng.randint(3)), 3)
        theirs = _build(ref, recipe, _variables(ref))
        ours = _build(A, recipe, _variables(A))
        theirs, ours = ref.wrap_if_literal(theirs), A.wrap_if_literal(ours)
        assert repr(ours) == repr(theirs), (case, recipe)
        assert ours.ndim == theirs.ndim
        want = evaluate(ours, {k: inputs[k] for k in ours.input_types})          # declared semantics, float64
        try:
            their_plan = dump_plan(theirs)
        except Exception:                                                      # reference planner defect
            dump_plan(ours)                                                    # ours must still plan
            continue
        assert _canon(dump_plan(ours)) == _canon(their_plan), (case, recipe)
        planned += 1
        if _has_batch_axes(their_plan):
            continue
        try:
            got = theirs.compile()(**{k: inputs[k] for k in theirs.input_types})
        except Exception:
            continue
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=2e-4, atol=2e-4,
                                   err_msg=repr(ours))
        valued += 1
    assert planned >= N_CASES * 0.8 and valued >= N_CASES * 0.5, (planned, valued)


def test_match_agrees_with_the_unmodified_reference():
    """``match(expression, template, slot)`` (algebra.py:1037-1063) on random expression / template
    pairs: same answer (by canonical form) or None on both sides."""
    ref = load_reference_algebra()
    rng = np.random.RandomState(911)
    found = missed = 0
    for case in range(120):
        left, other = _recipe(rng, 2, 2), _recipe(rng, 2, 2)
        fill = _recipe(rng, int(rng.choice([1, 2])), 1)
        same_template = rng.rand() < 0.6
        results = []
        for ns in (ref, A):
            variables = _variables(ns)
            slot = ns.var('slot', _rank(fill))
            expr = ns.dot(_build(ns, left, variables), _build(ns, fill, variables))
            template = ns.dot(_build(ns, left if same_template else other, variables), slot)
            try:
                out = ns.match(expr, template, slot)
            except Exception as exc:                   # same exception type on both sides
                out = type(exc).__name__
            results.append(None if out is None else (out if isinstance(out, str) else repr(ns.wrap_if_literal(out))))
        assert results[0] == results[1], (case, left, fill, same_template, results)
        if results[0] is None:
            missed += 1
        else:
            found += 1
    assert found >= 40 and missed >= 5, (found, missed)


def _rank(recipe):
    op = recipe[0]
    if op == 'var':
        return 2 if recipe[1] in ('X', 'Y') else 1
    if op in ('T', 'outer'):
        return 2
    if op == 'scale':
        return _rank(recipe[2])
    if op in ('sumax', 'diag'):
        return 1
    if op in ('sumall', 'trace'):
        return 0
    if op == 'dot':
        return _rank(recipe[1]) + _rank(recipe[2]) - 2
    return _rank(recipe[1])


def _commute(rng, recipe):
    """The same expression with the operands of some commutative nodes swapped."""
    op = recipe[0]
    if op == 'var':
        return recipe
    if op in ('mul', 'add') and rng.rand() < 0.7:
        return (op, _commute(rng, recipe[2]), _commute(rng, recipe[1]))
    return (op,) + tuple(_commute(rng, r) if isinstance(r, tuple) else r for r in recipe[1:])


def test_expression_equality_agrees_with_the_unmodified_reference():
    """``==`` on expressions (einsum isomorphism up to factor order and index renaming,
    algebra.py:983-1034): same verdict as the reference on pairs that are equal by commutativity and on
    unrelated pairs."""
    ref = load_reference_algebra()
    rng = np.random.RandomState(31)
    equal = unequal = 0
    for case in range(150):
        rank = int(rng.randint(3))
        first = _recipe(rng, rank, 3)
        second = _commute(rng, first) if rng.rand() < 0.5 else _recipe(rng, rank, 3)
        verdicts = []
        for ns in (ref, A):
            variables = _variables(ns)
            a = ns.wrap_if_literal(_build(ns, first, variables))
            b = ns.wrap_if_literal(_build(ns, second, variables))
            verdicts.append(bool(a == b))
        assert verdicts[0] == verdicts[1], (case, first, second, verdicts)
        equal += verdicts[0]
        unequal += not verdicts[0]
    assert equal >= 40 and unequal >= 20, (equal, unequal)
