"""GPU parity of the plan executor: every golden expression case is compiled with
``bayesic_b200.algebra`` and run through the C-ABI (``bb_plan_execute``) on the device,
then compared with (a) the unmodified reference's own output recorded in tests/golden/,
(b) the float64 declared-semantics value, (c) the numpy evaluation of the very
descriptor the executor was given.  Tolerance: rtol 1e-4 (north-star), atol 1e-5."""
import numpy as np
import pytest

import bayesic_b200.algebra as A
from bayesic_b200.backend.compiled import compile_expressions, compile_many
from oracle.descriptor_eval import evaluate_descriptor
from tests.golden.cases import CASES, make_inputs
from tests.golden_util import load_algebra_golden

pytestmark = pytest.mark.gpu

META, ARR = load_algebra_golden()
INPUTS = make_inputs()
RTOL, ATOL = 1e-4, 1e-5


def _run(build, fuse):
    expr = build(A)
    fn = compile_expressions([expr], single=True, fuse=fuse)
    used = {k: INPUTS[k] for k in expr.input_types}
    return fn, fn(**used), used


@pytest.mark.parametrize('fuse', [True, False], ids=['fused', 'unfused'])
@pytest.mark.parametrize('name,build', CASES, ids=[c[0] for c in CASES])
def test_case_matches_reference_and_oracle(name, build, fuse):
    entry = META[name]
    fn, got, used = _run(build, fuse)
    got = np.asarray(got)
    want = ARR[name + '__f64']
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    if entry['ref_output'] and entry['ref_matches_declared']:
        np.testing.assert_allclose(got, ARR[name + '__ref'], rtol=RTOL, atol=ATOL)
    low = fn.plan.lowered
    arrays = [used[n] if n else low.bound_constants[i] for i, n in enumerate(low.input_names)]
    (desc_val,) = evaluate_descriptor(low.nodes, low.outputs, arrays)
    np.testing.assert_allclose(got, desc_val, rtol=RTOL, atol=ATOL)


def test_kernels_really_launched():
    X = A.var('X', 2)
    fn = A.dot(X.T, X).compile()
    fn(X=INPUTS['D'])
    assert fn.plan.last_launches >= 1


def test_shape_size_eye_like_the_reference():
    X = A.var('X', 2)
    data = [[1, 2], [3, 4], [5, 6]]
    assert X.shape[0].compile()(X=data) == 3
    assert X.shape[1].compile()(X=data) == 2
    assert X.size.compile()(X=data) == 6
    a = A.var('a', 0, 'int32')
    fn = A.eye(a).compile()
    np.testing.assert_equal(fn(a=2), np.eye(2))
    np.testing.assert_equal(fn(a=5), np.eye(5))
    assert A.add(1, 1).compile()() == 2


def test_device_resident_inputs_stay_on_device():
    import torch
    X = A.var('X', 2)
    fn = A.dot(X.T, X).compile()
    Xd = torch.from_numpy(INPUTS['D']).cuda()
    out = fn(X=Xd)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), INPUTS['D'].T.astype(np.float64) @ INPUTS['D'],
                               rtol=RTOL, atol=ATOL)


def test_multi_output_plan_shares_the_pass():
    D, L, eta = A.var('D', 2), A.var('L', 2), A.var('eta', 1)
    exprs = [A.dot(D.T, D), A.sum(D, axis=0),
             -0.5 * A.trace(A.dot(L, A.dot(D.T, D))) + A.dot(D, eta).sum() - 0.5 * D.shape[0] * 1.75]
    fn = compile_many(exprs)
    s2, s1, elbo = fn(D=INPUTS['D'], L=INPUTS['L'], eta=INPUTS['eta'])
    Dv, Lv, ev = (INPUTS[k].astype(np.float64) for k in ('D', 'L', 'eta'))
    np.testing.assert_allclose(s2, Dv.T @ Dv, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(s1, Dv.sum(0), rtol=RTOL, atol=ATOL)
    want = -0.5 * np.trace(Lv @ Dv.T @ Dv) + (Dv @ ev).sum() - 0.5 * Dv.shape[0] * 1.75
    np.testing.assert_allclose(elbo, want, rtol=RTOL, atol=ATOL)
    kinds = [n['kind'] for n in fn.plan.lowered.nodes]
    assert kinds.count(21) == 1          # one SYRK node serves both outputs


def test_call_time_errors():
    X, y = A.var('X', 2), A.var('y', 1)
    fn = A.dot(X, y).compile()
    with pytest.raises(KeyError):
        fn(X=INPUTS['X'])
    with pytest.raises(TypeError):
        fn(X=INPUTS['x'], y=INPUTS['y'])                 # wrong rank
    with pytest.raises(ValueError):
        fn(X=INPUTS['X'], y=np.ones(7, dtype='float32'))  # contracted extents differ
    counts = A.var('counts', 1, dtype='int64')
    total = A.sum(counts, axis=0).compile()
    assert float(total(counts=np.array([3, 4, (1 << 24) - 8], dtype='int64'))) == float((1 << 24) - 1)
    with pytest.raises(ValueError):
        total(counts=np.array([3, (1 << 24) + 1], dtype='int64'))   # not exact in float32: refused, not rounded


def test_larger_random_contractions():
    rng = np.random.RandomState(7)
    X, Y, S = A.var('X', 2), A.var('Y', 2), A.var('S', 3)
    Xv, Yv = rng.randn(300, 70).astype('float32'), rng.randn(70, 129).astype('float32')
    np.testing.assert_allclose(A.dot(X, Y).compile()(X=Xv, Y=Yv), Xv.astype('f8') @ Yv, rtol=RTOL, atol=1e-4)
    Sv = rng.randn(17, 33, 9).astype('float32')
    np.testing.assert_allclose(S.sum(axis=(0, 2)).compile()(S=Sv), Sv.astype('f8').sum((0, 2)), rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(A.sum(S, 1).compile()(S=Sv), Sv.astype('f8').sum(1), rtol=RTOL, atol=1e-4)
    big = rng.randn(20000, 24).astype('float32')
    got = A.dot(X.T, X).compile()(X=big)
    np.testing.assert_allclose(got, big.astype('f8').T @ big, rtol=RTOL, atol=1e-3)
    got = compile_expressions([A.dot(X.T, X)], single=True, fuse=False)(X=big)
    np.testing.assert_allclose(got, big.astype('f8').T @ big, rtol=RTOL, atol=1e-3)


def test_logsoftmax_expression_is_stabilised():
    Lg = A.var('Lg', 2)
    expr = Lg - A.log(A.sum(A.exp(Lg), axis=1)).dimshuffle(0, 'x')
    logits = np.array([[100.0, 101.0, 99.0], [-200.0, -201.0, -202.0]], dtype='float32')
    from scipy.special import log_softmax
    got = expr.compile()(Lg=logits)     # the unfused float32 spelling overflows to -inf here
    np.testing.assert_allclose(got, log_softmax(logits.astype('f8'), axis=1), rtol=RTOL, atol=ATOL)


# ---- large contractions of compiled plans land on the tcgen05 projection / Gram kernels ---------

def _scale_close(got, want, left_norms, right_norms, tol=3e-5):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + tol * np.outer(left_norms, right_norms))


def test_compiled_dot_x_wt_runs_on_the_row_projection_kernel():
    rng = np.random.RandomState(21)
    Xh = (rng.randn(8192, 128) * 1.2 + 0.1).astype(np.float32)
    Wh = (rng.randn(64, 128) / 11.0).astype(np.float32)
    X, W, V = A.var('X', 2), A.var('W', 2), A.var('V', 2)
    want = Xh.astype(np.float64) @ Wh.astype(np.float64).T
    norms = (np.linalg.norm(Xh.astype(np.float64), axis=1), np.linalg.norm(Wh.astype(np.float64), axis=1))
    fn = A.dot(X, W.T).compile()                       # plan: _tensordot(X, _dimshuffle(W,1,0), [1],[0])
    _scale_close(fn(X=Xh, W=Wh), want, *norms)
    assert fn.plan.last_launches <= 3                  # split W, projection (not the split-K SIMT GEMM + reduce)
    fn2 = A.dot(X, V).compile()                        # W given as (features, q): transposed copy first
    _scale_close(fn2(X=Xh, V=np.ascontiguousarray(Wh.T)), want, *norms)


def test_compiled_dot_xt_r_runs_on_the_column_projection_kernel():
    rng = np.random.RandomState(22)
    Xh = (rng.randn(8192, 256) * 1.2 + 0.1).astype(np.float32)
    Rh = rng.randn(8192, 64).astype(np.float32)
    X, R = A.var('X', 2), A.var('R', 2)
    fn = A.dot(X.T, R).compile()                       # plan: _tensordot(_dimshuffle(X,1,0), R, [1],[0])
    want = Xh.astype(np.float64).T @ Rh.astype(np.float64)
    _scale_close(fn(X=Xh, R=Rh), want, np.linalg.norm(Xh.astype(np.float64), axis=0),
                 np.linalg.norm(Rh.astype(np.float64), axis=0))


def test_compiled_syrk_large_d_runs_on_the_cta_pair_kernel():
    rng = np.random.RandomState(23)
    Xh = (rng.randn(5000, 256) * 1.2 + 0.1).astype(np.float32)
    X = A.var('X', 2)
    fn = A.dot(X.T, X).compile()
    want = Xh.astype(np.float64).T @ Xh.astype(np.float64)
    nrm = np.linalg.norm(Xh.astype(np.float64), axis=0)
    _scale_close(fn(X=Xh), want, nrm, nrm)


@pytest.mark.parametrize('n,d', [(20000, 24), (20000, 256), (9000, 1000), (6000, 2048), (20000, 6)])
def test_compiled_matrix_vector_contractions_over_the_data_axis(n, d):
    """dot(X.T, y) / dot(y, X) (contraction over rows) and dot(X, w) / dot(w, X.T) (over features) on a
    tall matrix: served by the one-pass matrix-vector kernels when the layout allows (d % 4 == 0),
    by the split-K GEMM otherwise (d = 6) -- same values either way."""
    rng = np.random.RandomState(n + d)
    Xh = (rng.randn(n, d) * 1.1 + 0.2).astype(np.float32)
    yh, wh = rng.randn(n).astype(np.float32), (rng.randn(d) / np.sqrt(d)).astype(np.float32)
    X, y, w = A.var('X', 2), A.var('y', 1), A.var('w', 1)
    X64 = Xh.astype(np.float64)
    col_norms, row_norms = np.linalg.norm(X64, axis=0), np.linalg.norm(X64, axis=1)
    ynorm, wnorm = np.array([np.linalg.norm(yh.astype(np.float64))]), np.array([np.linalg.norm(wh.astype(np.float64))])
    want_cols, want_rows = X64.T @ yh.astype(np.float64), X64 @ wh.astype(np.float64)
    f = A.dot(X.T, y).compile()
    _scale_close(f(X=Xh, y=yh).reshape(d, 1), want_cols.reshape(d, 1), col_norms, ynorm)
    if d % 4 == 0:
        assert f.plan.last_launches <= 2               # one pass + finalize, not GEMM tiles + reduce
    _scale_close(A.dot(y, X).compile()(X=Xh, y=yh).reshape(d, 1), want_cols.reshape(d, 1), col_norms, ynorm)
    g = A.dot(X, w).compile()
    _scale_close(g(X=Xh, w=wh).reshape(n, 1), want_rows.reshape(n, 1), row_norms, wnorm)
    assert g.plan.last_launches == 1
    _scale_close(A.dot(w, X.T).compile()(X=Xh, w=wh).reshape(n, 1), want_rows.reshape(n, 1), row_norms, wnorm)


def test_repeated_factors_and_terms_are_evaluated_not_merged():
    """Round-1 advisor finding: sharing sub-trees by the expressions' frozenset-based ``==``
    computed 2*X*X*Y for X*X*Y + X*Y.  Sharing is by identity / strict value numbering now."""
    rng = np.random.RandomState(31)
    Xh, Yh = rng.randn(37, 11).astype(np.float32), rng.randn(37, 11).astype(np.float32)
    X, Y = A.var('X', 2), A.var('Y', 2)
    X64, Y64 = Xh.astype(np.float64), Yh.astype(np.float64)
    got = (X * X * Y + X * Y).compile()(X=Xh, Y=Yh)
    np.testing.assert_allclose(got, X64 * X64 * Y64 + X64 * Y64, rtol=RTOL, atol=ATOL)
    a, b = compile_many([X + X + Y, X + Y])(X=Xh, Y=Yh)
    np.testing.assert_allclose(a, 2 * X64 + Y64, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(b, X64 + Y64, rtol=RTOL, atol=ATOL)
    got = (X ** 3 + X ** 2 + X * X * X).compile()(X=np.abs(Xh) + 0.5)
    x = np.abs(X64) + 0.5
    np.testing.assert_allclose(got, 2 * x ** 3 + x ** 2, rtol=RTOL, atol=ATOL)


def test_repeated_host_calls_replay_a_graph_and_track_the_inputs():
    """numpy in / numpy out, the reference's calling convention (algebra.py:50-58): from the second call with the
    same input signature on, the plan is one CUDA-graph launch over persistent buffers -- the values must follow
    the inputs of each call, the device-resident route must agree, and a new shape must be a new signature."""
    import torch
    X, Lm = A.var('X', 2), A.var('L', 2)
    fn = A.trace(A.dot(Lm, A.dot(X.T, X))).compile()
    rng = np.random.RandomState(3)
    for n in (1000, 1000, 1000, 777, 1000):
        Xh = rng.randn(n, 16).astype(np.float32)
        a = rng.randn(16, 16)
        Lh = (a @ a.T / 16 + np.eye(16)).astype(np.float32)
        want = np.einsum('de,nd,ne->', Lh.astype('f8'), Xh.astype('f8'), Xh.astype('f8'))
        got = fn(X=Xh, L=Lh)
        assert isinstance(got, np.ndarray) or np.isscalar(got)
        np.testing.assert_allclose(float(got), want, rtol=1e-5)
        on_dev = fn(X=torch.from_numpy(Xh).cuda(), L=torch.from_numpy(Lh).cuda())
        np.testing.assert_allclose(float(on_dev), want, rtol=1e-5)
        pinned = fn(X=torch.from_numpy(Xh).pin_memory(), L=Lh)
        np.testing.assert_allclose(float(pinned), want, rtol=1e-5)
    assert fn.plan.last_launches > 0
    many = compile_many([A.dot(X.T, X), A.sum(X, axis=0)])
    for _ in range(3):
        Xh = rng.randn(500, 16).astype(np.float32)
        xtx, s1 = many(X=Xh)
        np.testing.assert_allclose(xtx, Xh.astype('f8').T @ Xh.astype('f8'), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(s1, Xh.astype('f8').sum(0), rtol=1e-4, atol=1e-3)
    with pytest.raises(ValueError):
        A.dot(X, A.var('y', 1)).compile()(X=Xh, y=np.ones(7, dtype='float32'))
