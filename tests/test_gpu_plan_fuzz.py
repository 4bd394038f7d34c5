"""GPU parity of the plan executor on seeded random expressions (the recipes of tests/test_reference_crosscheck.py,
plus the pointwise vocabulary): each is compiled with ``bayesic_b200.algebra`` -- the reference's entry point,
``Expression.compile`` -> ``f(**inputs)`` (bayesic/algebra.py:50-58) -- run through the C-ABI on the device in
float32, and compared with the float64 declared semantics and with the numpy evaluation of the very descriptor the
executor was given.  Tolerance: rtol 1e-4 (north-star) with an absolute floor of 1e-5 of the largest entry, which
is where float32 cancellation in a depth-3 product of 5 x 5 matrices sits."""
import numpy as np
import pytest

from oracle.descriptor_eval import evaluate_descriptor
from oracle.semantics import evaluate
from tests.test_lowering_fuzz import _decorated, _expressions

pytestmark = pytest.mark.gpu

DATA = np.random.RandomState(11)
INPUTS = {'X': DATA.randn(5, 5).astype(np.float32), 'Y': DATA.randn(5, 5).astype(np.float32),
          'x': DATA.randn(5).astype(np.float32), 'y': DATA.randn(5).astype(np.float32)}


def _check(e, inputs=None, floor=1e-5):
    inputs = INPUTS if inputs is None else inputs
    used = {k: inputs[k] for k in e.input_types}
    fn = e.compile()
    got = np.asarray(fn(**used), dtype=np.float64)
    want = np.asarray(evaluate(e, used), dtype=np.float64)
    assert got.shape == want.shape, repr(e)
    atol = floor * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=atol, err_msg=repr(e))
    low = fn.plan.lowered
    arrays = [used[n] if n else low.bound_constants[i] for i, n in enumerate(low.input_names)]
    (desc_val,) = evaluate_descriptor(low.nodes, low.outputs, arrays)
    np.testing.assert_allclose(got, np.asarray(desc_val, dtype=np.float64), rtol=1e-4, atol=atol, err_msg=repr(e))
    return fn


def test_random_expressions_on_the_device():
    launched = 0
    for e in _expressions(606, 60, 3):
        launched += _check(e).plan.last_launches >= 1
    assert launched >= 30           # a recipe can be a bare variable (a copy, no kernel); most launch kernels


def test_random_pointwise_expressions_on_the_device():
    for e in _decorated(707, 40):
        _check(e)


def test_random_multi_output_plans_on_the_device():
    """groups of three expressions compiled as ONE plan (compile_many): shared sub-trees, several outputs, and a
    second call on the same handle (the host fast path replays its CUDA graph) give the same values"""
    from bayesic_b200.backend.compiled import compile_many
    exprs = _expressions(808, 45, 3)
    for i in range(0, len(exprs), 3):
        group = exprs[i:i + 3]
        names = sorted(set().union(*[set(e.input_types) for e in group]))
        used = {k: INPUTS[k] for k in names}
        fn = compile_many(group)
        first = [np.asarray(v, dtype=np.float64) for v in fn(**used)]
        second = [np.asarray(v, dtype=np.float64) for v in fn(**used)]
        for e, g, g2 in zip(group, first, second):
            want = np.asarray(evaluate(e, {k: INPUTS[k] for k in e.input_types}), dtype=np.float64)
            atol = 1e-5 * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
            np.testing.assert_allclose(g, want, rtol=1e-4, atol=atol, err_msg=repr(e))
            np.testing.assert_array_equal(g, g2)


def test_random_expressions_on_the_device_at_a_ragged_extent():
    """the same recipes on 33 x 33 operands: extents that are no multiple of any tile edge of the generic kernels
    (absolute floor 1e-4 of the largest entry: three chained float32 products of 33 terms)"""
    data = np.random.RandomState(12)
    inputs = {'X': (data.randn(33, 33) / 6).astype(np.float32), 'Y': (data.randn(33, 33) / 6).astype(np.float32),
              'x': data.randn(33).astype(np.float32), 'y': data.randn(33).astype(np.float32)}
    for e in _expressions(909, 40, 3):
        _check(e, inputs, floor=1e-4)
