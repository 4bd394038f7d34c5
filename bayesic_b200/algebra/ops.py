"""Public tensor operations: thin constructors of einsum index patterns, the
pointwise functions, and the operator overloads on ``Expression``.

Mirrors the API surface of ``bayesic/algebra.py:1149-1278`` and ``:1416-1478``:
same names, argument order and error behaviour (``ValueError`` on bad axes).
``sum`` shadows the builtin exactly like the reference does.
"""
from .expr import (Expression, wrap_if_literal, with_wrapped_literals, autobroadcast_or_match,
                   elemwise, add, OP_LOG, OP_EXP, OP_POW, OP_ABS, OP_LGAMMA)
from .einsum import einsum, sum_index, out_index

__all__ = ['dot', 'tensordot', 'mul', 'outer', 'sum', 'trace', 'diagonal', 'transpose',
           'dimshuffle', 'div', 'neg', 'sub', 'log', 'exp', 'pow', 'abs_', 'lgamma']


def _outs(start, stop):
    return [out_index(n) for n in range(start, stop)]


@with_wrapped_literals
def dot(X, Y):
    """Contract the last axis of ``X`` with the first of ``Y`` (``algebra.py:1151-1158``)."""
    k = sum_index(0)
    return einsum([(X, _outs(0, X.ndim - 1) + [k]),
                   (Y, [k] + _outs(X.ndim - 1, X.ndim + Y.ndim - 2))],
                  X.ndim + Y.ndim - 2)


def tensordot(X, Y, X_sum_axes, Y_sum_axes, X_batch_axes=(), Y_batch_axes=()):
    """Generalised (batched) tensordot; result axes are
    batch + X's remaining axes + Y's remaining axes (``algebra.py:1161-1194``)."""
    X, Y = wrap_if_literal(X), wrap_if_literal(Y)
    x_idx, y_idx = [None] * X.ndim, [None] * Y.ndim
    for n, axis in enumerate(X_sum_axes):
        x_idx[axis] = sum_index(n)
    for n, axis in enumerate(Y_sum_axes):
        y_idx[axis] = sum_index(n)
    for n, axis in enumerate(X_batch_axes):
        x_idx[axis] = out_index(n)
    for n, axis in enumerate(Y_batch_axes):
        y_idx[axis] = out_index(n)
    nxt = len(X_batch_axes)
    for idx in (x_idx, y_idx):
        for axis in range(len(idx)):
            if idx[axis] is None:
                idx[axis] = out_index(nxt)
                nxt += 1
    return einsum([(X, x_idx), (Y, y_idx)], nxt)


@with_wrapped_literals
def mul(*args):
    """Hadamard product; scalars broadcast, everything else must agree in rank
    (``algebra.py:1197-1207``)."""
    rank = max(a.ndim for a in args)
    return einsum([(a, _outs(0, a.ndim)) for a in (autobroadcast_or_match(a, rank) for a in args)], rank)


@with_wrapped_literals
def outer(X, Y):
    """Tensor product of any two tensors (``algebra.py:1210-1221``)."""
    return einsum([(X, _outs(0, X.ndim)), (Y, _outs(X.ndim, X.ndim + Y.ndim))], X.ndim + Y.ndim)


def sum(X, axis=None):
    """Sum over all axes, or over ``axis`` (int or iterable) -- ``algebra.py:1227-1242``."""
    X = wrap_if_literal(X)
    if isinstance(axis, int):
        axis = [axis]
    if axis is None:
        axis = range(X.ndim)
    idx, n_sum, n_out = [], 0, 0
    for a in range(X.ndim):
        if a in axis:
            idx.append(sum_index(n_sum))
            n_sum += 1
        else:
            idx.append(out_index(n_out))
            n_out += 1
    return einsum([(X, idx)], n_out)


@with_wrapped_literals
def trace(X):
    return einsum([(X, [sum_index(0), sum_index(0)])], 0)


@with_wrapped_literals
def diagonal(X):
    return einsum([(X, [out_index(0), out_index(0)])], 1)


@with_wrapped_literals
def transpose(X):
    return dimshuffle(X, *reversed(range(X.ndim)))


def dimshuffle(X, *axes):
    """Permute axes; ``'x'`` inserts a broadcastable axis (``algebra.py:1262-1277``)."""
    X = wrap_if_literal(X)
    idx = [None] * X.ndim
    for position, axis in enumerate(axes):
        if axis == 'x':
            continue
        if idx[axis] is not None:
            raise ValueError("dimshuffle: same input axis can't occur twice")
        idx[axis] = out_index(position)
    if any(i is None for i in idx):
        raise ValueError("dimshuffle: can't drop an axis")
    return einsum([(X, idx)], len(axes))


# ---- pointwise --------------------------------------------------------------

@with_wrapped_literals
def div(X, Y):
    """``X * Y**-1`` so that division takes part in einsums (``algebra.py:1418-1422``)."""
    return mul(X, Y ** -1)


@with_wrapped_literals
def neg(X):
    return -1 * X


@with_wrapped_literals
def sub(X, Y):
    return add(X, -Y)


def log(X):
    return elemwise(OP_LOG, X)


def exp(X):
    return elemwise(OP_EXP, X)


def pow(X, Y):
    return elemwise(OP_POW, X, Y)


def abs_(X):
    return elemwise(OP_ABS, X)


def lgamma(X):
    """log Gamma(X), pointwise -- not in the reference's vocabulary; an extension for the
    exponential-family log-normalisers."""
    return elemwise(OP_LGAMMA, X)


# ---- operator overloads (algebra.py:1451-1478) -------------------------------

def _flipped(fn):
    def flipped(a, b):
        return fn(b, a)
    return flipped


def _add(*terms):
    return add(*terms)


Expression.__add__ = _add
Expression.__radd__ = _flipped(_add)
Expression.__sub__ = sub
Expression.__rsub__ = _flipped(sub)
Expression.__mul__ = mul
Expression.__rmul__ = _flipped(mul)
Expression.__truediv__ = div
Expression.__rtruediv__ = _flipped(div)
Expression.__pow__ = pow
Expression.__rpow__ = _flipped(pow)
Expression.__matmul__ = dot
Expression.__rmatmul__ = _flipped(dot)
Expression.__neg__ = neg
Expression.__abs__ = abs_
Expression.T = property(transpose)
Expression.dimshuffle = dimshuffle
Expression.sum = sum
Expression.dot = dot
