"""float64 numpy evaluation of an expression tree by its declared semantics.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Duck-typed on class names so
the same evaluator runs over the reference's node classes
(``/root/reference/bayesic/algebra.py``) and over ``bayesic_b200.algebra``'s.

Restated semantics, with the reference lines they follow:

* ``Einsum``       T[out...] = sum_{sum idx} prod_f f[idx_f]; an out number that no
                   factor carries is an extent-1 axis          algebra.py:314-346
* ``_sum``         X.sum(axis=axes)                             algebra.py:1290-1291
* ``_mul``         broadcasting product                         algebra.py:1302-1306
* ``_dimshuffle``  transpose + insert extent-1 axes for 'x'     algebra.py:1318-1319
* ``_tensordot``   result axes = batch + X_other + Y_other      algebra.py:1161-1171
                   (the reference *evaluator* for the batched case is broken,
                   algebra.py:1370-1373 and :1380; NOT reproduced)
* ``_diagonal``    numpy.diagonal(X, 0, a1, a2): diagonal last  algebra.py:1405-1407
* elementwise      add/mul/log/exp/pow/abs_                     algebra.py:201, 223, 1435-1448
* ``shape``        runtime extent, integer                      algebra.py:154-155
* ``eye``          identity of the first extent                 algebra.py:257-258
"""
import numpy as np

_LETTERS = 'abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ'

_POINTWISE = {
    'add': lambda *xs: _fold(np.add, xs),
    'mul': lambda *xs: _fold(np.multiply, xs),
    'log': np.log, 'exp': np.exp, 'pow': np.power, 'abs_': np.abs,
    'lgamma': lambda x: __import__('scipy.special', fromlist=['gammaln']).gammaln(x),
}


def _fold(fn, xs):
    out = xs[0]
    for x in xs[1:]:
        out = fn(out, x)
    return out


def einsum_f64(factor_arrays, index_tuples, ndim, dtype=np.float64):
    """``T[o0..o{ndim-1}] = sum prod factor[idx]`` with out axes nobody carries
    given extent 1 (algebra.py:340-345)."""
    letters = {}

    def letter(i):
        if i not in letters:
            letters[i] = _LETTERS[len(letters)]
        return letters[i]

    subs = [''.join(letter(tuple(i)) for i in idx) for idx in index_tuples]
    carried = [('out', n) for n in range(ndim) if ('out', n) in letters]
    out_sub = ''.join(letters[i] for i in carried)
    arrays = [np.asarray(a, dtype=dtype) for a in factor_arrays]
    if arrays:
        result = np.einsum(','.join(subs) + '->' + out_sub, *arrays)
    else:
        result = np.asarray(1.0, dtype=dtype)
    sel = tuple(slice(None) if ('out', n) in letters else np.newaxis for n in range(ndim))
    return np.asarray(result)[sel]


def evaluate(expr, inputs, dtype=np.float64):
    """Value of ``expr`` (reference or bayesic_b200 node classes) in ``dtype``."""
    memo = {}

    def ev(node):
        key = id(node)
        if key not in memo:
            memo[key] = _ev(node)
        return memo[key]

    def _ev(node):
        kind = type(node).__name__
        if kind == 'var':
            value = np.asarray(inputs[node.name])
            return value.astype(dtype) if value.dtype.kind == 'f' else value
        if kind == 'constant':
            value = np.asarray(node.value)
            return value.astype(dtype) if value.dtype.kind == 'f' else value
        if kind == 'shape':
            return np.asarray(np.shape(ev(node.parents[0]))[node.axis], dtype=np.int64)
        if kind == 'eye':
            return np.eye(int(ev(node.parents[0])), dtype=dtype)
        if kind == 'Einsum':
            arrays = [ev(f) for f, _ in node.factors_and_indices]
            return einsum_f64(arrays, [idx for _, idx in node.factors_and_indices], node.ndim, dtype)
        if kind in ('elemwise', 'add'):
            name = 'add' if kind == 'add' else node.name
            return _POINTWISE[name](*[ev(p) for p in node.parents])
        if kind == 'logdet':
            return logdet_spd(ev(node.parents[0])).astype(dtype)
        if kind == '_sum':
            return ev(node.parents[0]).sum(axis=tuple(node.axes))
        if kind == '_mul':
            return _fold(np.multiply, [ev(p) for p in node.parents])
        if kind == '_dimshuffle':
            value = ev(node.parents[0])
            value = np.transpose(value, [a for a in node.axes if a != 'x'])
            return value[tuple(np.newaxis if a == 'x' else slice(None) for a in node.axes)]
        if kind == '_diagonal':
            return np.diagonal(ev(node.parents[0]), 0, node.axis1, node.axis2)
        if kind == '_tensordot':
            return tensordot_declared(ev(node.parents[0]), ev(node.parents[1]),
                                      node.X_dot_axes, node.Y_dot_axes,
                                      node.X_batch_axes, node.Y_batch_axes)
        raise TypeError("oracle: unknown node type %s" % kind)

    return ev(expr)


def logdet_spd(X):
    """log|X| over the last two axes for symmetric positive-definite matrices; NaN for a matrix that is
    not positive definite (the log-determinant is defined through the Cholesky factor)."""
    X = np.asarray(X, dtype=np.float64)
    _, value = np.linalg.slogdet(X)
    if X.shape[-1] == 0:
        return value
    definite = np.linalg.eigvalsh(0.5 * (X + np.swapaxes(X, -1, -2))).min(axis=-1) > 0
    return np.where(definite, value, np.nan)


def tensordot_declared(X, Y, x_dot, y_dot, x_batch=(), y_batch=()):
    """Declared semantics of the (batched) tensordot, algebra.py:1161-1171:
    result axes = batch + X's other axes + Y's other axes."""
    x_idx, y_idx = [None] * X.ndim, [None] * Y.ndim
    n = 0
    for xa, ya in zip(x_dot, y_dot):
        x_idx[xa] = y_idx[ya] = _LETTERS[n]
        n += 1
    batch = ''
    for xa, ya in zip(x_batch, y_batch):
        x_idx[xa] = y_idx[ya] = _LETTERS[n]
        batch += _LETTERS[n]
        n += 1
    x_other = y_other = ''
    for a in range(X.ndim):
        if x_idx[a] is None:
            x_idx[a] = _LETTERS[n]
            x_other += _LETTERS[n]
            n += 1
    for a in range(Y.ndim):
        if y_idx[a] is None:
            y_idx[a] = _LETTERS[n]
            y_other += _LETTERS[n]
            n += 1
    return np.einsum('%s,%s->%s' % (''.join(x_idx), ''.join(y_idx), batch + x_other + y_other), X, Y)
