"""Expression cases shared by the golden generator (run against the unmodified
reference, ``oracle/make_golden.py``) and by the parity tests (run against
``bayesic_b200.algebra``).  Each case is ``(name, build)`` where ``build(A)`` takes
an algebra namespace -- the reference module or ours -- and returns an expression.

Sources: every numeric test of ``bayesic/tests/test_algebra.py:44-191``, every
plan-shape test ``:376-505``, and the hot-path expressions of SURVEY.md 3.2.
"""
import numpy as np
import numpy.random as npr


def make_inputs():
    """Same generator and draw order as ``bayesic/tests/test_algebra.py:15-38`` for
    the first five arrays, then the hot-path extras."""
    rng = npr.RandomState(1234)

    def randn(*shape):
        return rng.randn(*shape).astype('float32')

    inputs = {}
    inputs['X'] = randn(5, 5)
    inputs['Y'] = randn(5, 5)
    inputs['x'] = randn(5)
    inputs['y'] = randn(5)
    inputs['S'] = randn(3, 5, 7)
    inputs['Z'] = randn(5, 5)
    inputs['W'] = randn(5, 5)
    # hot-path shapes (small): data D[n, d], precision L[d, d], responsibilities R[n, k]
    n, d, k = 40, 6, 4
    inputs['D'] = randn(n, d)
    a = randn(d, d)
    inputs['L'] = (a @ a.T / d + np.eye(d)).astype('float32')
    logits = randn(n, k)
    r = np.exp(logits - logits.max(1, keepdims=True))
    inputs['R'] = (r / r.sum(1, keepdims=True)).astype('float32')
    inputs['Lg'] = (3 * logits).astype('float32')
    inputs['M'] = randn(k, d)
    inputs['eta'] = randn(d)
    inputs['P'] = np.abs(randn(5, 5)) + 0.5       # strictly positive, for log / pow
    inputs['T3'] = randn(4, 5, 5)
    inputs['U3'] = randn(4, 5, 3)
    # cfg4 / cfg5 expressions (appended last so the draws above are unchanged): regression targets t[n],
    # binary labels b01[n], parameter draws Wm[s, d]
    inputs['t'] = randn(n)
    inputs['b01'] = (rng.rand(n) < 0.5).astype('float32')
    inputs['Wm'] = (0.5 * randn(3, d)).astype('float32')
    # cfg3 logits: per-component precision-like matrices Ak[k, d, d] and offsets ck[k]
    a3 = randn(k, d, d)
    inputs['Ak'] = (np.einsum('kij,klj->kil', a3, a3) / d + np.eye(d)).astype('float32')
    inputs['ck'] = randn(k)
    return inputs


YLIT = make_inputs()['Y']


def _vars(A):
    v = {name: A.var(name, 2) for name in ('X', 'Y', 'Z', 'W', 'D', 'L', 'R', 'Lg', 'M', 'P', 'Wm')}
    v.update({name: A.var(name, 1) for name in ('x', 'y', 'eta', 't', 'b01', 'ck')})
    v.update({name: A.var(name, 3) for name in ('S', 'T3', 'U3', 'Ak')})
    return v


def _case(fn):
    def build(A):
        ns = _vars(A)
        return fn(A, **ns)
    return build


def _c(fn):
    # helper: fn(A, v) where v is an attribute bag
    class Bag(object):
        pass

    def build(A):
        bag = Bag()
        bag.__dict__.update(_vars(A))
        return fn(A, bag)
    return build


CASES = [
    # ---- numeric tests of the reference (test_algebra.py:44-191) -------------
    ('add', _c(lambda A, v: v.X + v.Y)),
    ('sub', _c(lambda A, v: v.X - v.Y)),
    ('abs', _c(lambda A, v: abs(v.X))),
    ('scalar_add', _c(lambda A, v: v.X + 1)),
    ('scalar_rsub', _c(lambda A, v: 1 - v.X)),
    ('scalar_mul', _c(lambda A, v: 2 * v.X)),
    ('add_literals', _c(lambda A, v: A.add(1, 1))),
    ('literal_array_mul', _c(lambda A, v: v.X * YLIT)),
    ('dot_mm', _c(lambda A, v: A.dot(v.X, v.Y))),
    ('dot_mv', _c(lambda A, v: v.X.dot(v.y))),
    ('dot_vv', _c(lambda A, v: A.dot(v.x, v.y))),
    ('mul', _c(lambda A, v: v.X * v.Y)),
    ('div', _c(lambda A, v: v.X / v.P)),
    ('pow_tensor', _c(lambda A, v: v.P ** v.Y)),
    ('pow_rscalar', _c(lambda A, v: 2 ** v.X)),
    ('pow_scalar', _c(lambda A, v: v.X ** 2)),
    ('log', _c(lambda A, v: A.log(v.P))),
    ('exp', _c(lambda A, v: A.exp(v.X))),
    ('transpose', _c(lambda A, v: v.X.T)),
    ('dimshuffle3', _c(lambda A, v: A.dimshuffle(v.S, 2, 0, 1))),
    ('bcast_add_col', _c(lambda A, v: v.X + v.x.dimshuffle(0, 'x'))),
    ('bcast_mul_row', _c(lambda A, v: v.X * A.dimshuffle(v.x, 'x', 0))),
    ('trace', _c(lambda A, v: A.trace(v.X))),
    ('diagonal', _c(lambda A, v: A.diagonal(v.X))),
    ('outer', _c(lambda A, v: A.outer(v.x, v.y))),
    ('sum_all', _c(lambda A, v: A.sum(v.S))),
    ('sum_axis0', _c(lambda A, v: A.sum(v.S, axis=0))),
    ('sum_axes02', _c(lambda A, v: v.S.sum(axis=(0, 2)))),
    ('shape0', _c(lambda A, v: v.X.shape[0])),
    ('size', _c(lambda A, v: v.S.size)),
    ('nested_collapse', _c(lambda A, v: A.dot(A.diagonal(A.dot(v.X, A.outer(v.x, v.y))), v.Y))),
    # ---- plan-shape tests of the reference (test_algebra.py:376-505) ---------
    ('plan_sum1', _c(lambda A, v: A.sum(v.X, 1))),
    ('plan_dimshuffle', _c(lambda A, v: A.dimshuffle(v.X, 1, 0))),
    ('plan_chain_right', _c(lambda A, v: A.dot(v.X, A.dot(v.Y, v.Z)))),
    ('plan_chain_left', _c(lambda A, v: A.dot(A.dot(v.X, v.Y), v.Z))),
    ('plan_chain_pairs', _c(lambda A, v: A.dot(A.dot(v.X, v.Y), A.dot(v.Z, v.W)))),
    ('plan_chain_mixed', _c(lambda A, v: A.dot(A.dot(v.X, A.dot(v.Y, v.Z)), v.W))),
    ('plan_no_sum', _c(lambda A, v: v.X * v.Y.T * v.x.dimshuffle(0, 'x'))),
    ('plan_sum_one_term', _c(lambda A, v: v.X.sum(1) * v.y)),
    ('plan_trace_dot', _c(lambda A, v: A.trace(A.dot(v.X.T, v.Y)))),
    ('plan_batched_vecvec', _c(lambda A, v: (v.X * v.Y.T).sum(axis=1))),
    ('plan_group_a', _c(lambda A, v: A.dot(v.Z, v.x * v.y))),
    ('plan_group_b', _c(lambda A, v: A.dot(v.Z * v.x.dimshuffle('x', 0), v.y))),
    ('plan_group_c', _c(lambda A, v: A.dot(v.Z * v.y.dimshuffle('x', 0), v.x))),
    ('plan_group_d', _c(lambda A, v: A.dot(v.X * v.Y, v.Z * v.W))),
    ('plan_group_e', _c(lambda A, v: A.tensordot(
        v.X.dimshuffle(0, 1, 'x') * v.Z.dimshuffle('x', 0, 1),
        v.Y.dimshuffle(0, 1, 'x') * v.W.dimshuffle('x', 0, 1),
        X_sum_axes=[1], Y_sum_axes=[1], X_batch_axes=[0, 2], Y_batch_axes=[0, 2]))),
    # ---- hot-path expressions ---------------------------------
    ('hot_sxx', _c(lambda A, v: A.dot(v.D.T, v.D))),
    ('hot_sx', _c(lambda A, v: A.sum(v.D, axis=0))),
    ('hot_trace_quad', _c(lambda A, v: A.trace(A.dot(v.L, A.dot(v.D.T, v.D))))),
    ('hot_rx', _c(lambda A, v: A.dot(v.R.T, v.D))),
    ('hot_rxx', _c(lambda A, v: A.einsum([
        (v.R, [('sum', 0), ('out', 0)]), (v.D, [('sum', 0), ('out', 1)]),
        (v.D, [('sum', 0), ('out', 2)])], 3))),
    ('hot_nk', _c(lambda A, v: A.sum(v.R, axis=0))),
    ('hot_lin_sum', _c(lambda A, v: A.dot(v.D, v.eta).sum())),
    ('hot_interaction', _c(lambda A, v: A.dot(v.D, v.M.T))),
    ('hot_rowquad', _c(lambda A, v: (A.dot(v.D, v.L) * v.D).sum(axis=1))),
    ('hot_elbo_terms', _c(lambda A, v: -0.5 * A.trace(A.dot(v.L, A.dot(v.D.T, v.D)))
                          + A.dot(v.D, v.eta).sum() - 0.5 * v.D.shape[0] * 1.75)),
    ('hot_logsoftmax', _c(lambda A, v: v.Lg - A.log(A.sum(A.exp(v.Lg), axis=1)).dimshuffle(0, 'x'))),
    ('hot_centered_scatter', _c(lambda A, v: A.dot((v.D - v.eta.dimshuffle('x', 0)).T,
                                                   v.D - v.eta.dimshuffle('x', 0)))),
    # ---- batched contractions (reference evaluator broken; declared semantics) --
    ('batched_mm', _c(lambda A, v: A.tensordot(v.T3, v.U3, [2], [1], [0], [0]))),
    ('batched_mv', _c(lambda A, v: A.tensordot(v.T3, v.T3,
                                               [2], [2], [0, 1], [0, 1]))),
    ('tensordot2', _c(lambda A, v: A.tensordot(v.T3, v.U3, [0, 1], [0, 1]))),
    ('planner_crash_case', _c(lambda A, v: A.dot(A.sum(v.D, 0), A.dot(v.L, v.eta)))),
    # cfg4 (conjugate SVI statistics) and cfg5 (reparameterised logistic gradient) as the user writes them
    ('hot_cfg4_xty', _c(lambda A, v: A.dot(v.D.T, v.t))),
    ('hot_cfg4_yty', _c(lambda A, v: A.dot(v.t, v.t))),
    ('hot_cfg5_z', _c(lambda A, v: A.dot(v.D, v.Wm.T))),
    ('hot_cfg5_loglik', _c(lambda A, v: A.sum(v.b01.dimshuffle(0, 'x') * A.dot(v.D, v.Wm.T)
                                              - A.log(1 + A.exp(A.dot(v.D, v.Wm.T))), axis=0))),
    ('hot_cfg5_grad', _c(lambda A, v: A.dot(v.D.T, v.b01.dimshuffle(0, 'x')
                                            - (1 + A.exp(-1 * A.dot(v.D, v.Wm.T))) ** -1))),
    # cfg3 logits c_k + x.b_k - 1/2 x^T A_k x as passes.GmmStep writes them (the reference plans a batched
    # _tensordot whose evaluator is broken, algebra.py:1370-1373: the golden records plan + declared value)
    ('hot_cfg3_logits', _c(lambda A, v: A.dot(v.D, v.M.T) + (-0.5) * A.einsum(
        [(v.D, [('out', 0), ('sum', 0)]), (v.Ak, [('out', 1), ('sum', 0), ('sum', 1)]),
         (v.D, [('out', 0), ('sum', 1)])], 2) + v.ck.dimshuffle('x', 0))),
]
