"""theano.tensor stand-in (numpy-backed lazy graph).  See package docstring."""
import builtins as _b
import numpy as _np


class Variable(object):
    """Symbolic tensor.  ``_ev(env)`` computes its numpy value."""

    def __init__(self, ev, ndim, dtype, name=None):
        self._ev = ev
        self.ndim = ndim
        self.dtype = dtype
        self.name = name

    # -- attributes / methods used by the reference -------------------------
    @property
    def shape(self):
        return _ShapeOf(self)

    def sum(self, axis=None):
        if axis is not None and not isinstance(axis, int):
            axis = tuple(axis)
        n_removed = self.ndim if axis is None else (1 if isinstance(axis, int) else len(axis))
        return Variable(lambda env: self._ev(env).sum(axis=axis), self.ndim - n_removed, self.dtype)

    def dimshuffle(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (list, tuple)):
            axes = tuple(axes[0])

        def ev(env):
            v = self._ev(env)
            perm = [a for a in axes if a != 'x']
            # axes not mentioned must be droppable (size 1) -- not used by the reference
            v = _np.transpose(v, perm)
            idx = tuple(_np.newaxis if a == 'x' else slice(None) for a in axes)
            return v[idx]

        return Variable(ev, len(axes), self.dtype)

    def reshape(self, shape):
        shape = list(shape)

        def ev(env):
            return self._ev(env).reshape([int(_value(s, env)) for s in shape])

        return Variable(ev, len(shape), self.dtype)

    def __mul__(self, other):
        return mul(self, other)

    __rmul__ = __mul__

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


def _value(x, env):
    return x._ev(env) if isinstance(x, Variable) else x


class _ShapeOf(object):
    def __init__(self, var):
        self._var = var

    def __getitem__(self, i):
        v = self._var
        return Variable(lambda env: _np.asarray(v._ev(env).shape[i], dtype='int64'), 0, 'int64')


class TensorType(object):
    """T.TensorType(dtype, broadcastable)(name)  (algebra.py:44)."""

    def __init__(self, dtype, broadcastable):
        self.dtype = dtype
        self.ndim = len(broadcastable)

    def __call__(self, name=None):
        var = Variable(None, self.ndim, self.dtype, name)
        var._ev = lambda env: env[var]
        return var


def constant(value):
    """T.constant(v): has .ndim and .dtype (algebra.py:132-134)."""
    arr = _np.asarray(value)
    if arr.dtype == _np.float64 and _np.isscalar(value):
        # theano stores python floats as floatX-compatible scalars; keep float64
        pass
    return Variable(lambda env: arr, arr.ndim, str(arr.dtype))


class _ScalarOp(object):
    def __init__(self, name):
        self.name = name


class _Elemwise(object):
    """Callable elementwise op with a ``scalar_op.name`` (algebra.py:202)."""

    def __init__(self, name, fn):
        self.scalar_op = _ScalarOp(name)
        self._fn = fn

    def __call__(self, *args):
        ndim = _b.max([a.ndim if isinstance(a, Variable) else _np.ndim(a) for a in args] + [0])
        dtypes = [a.dtype if isinstance(a, Variable) else _np.asarray(a).dtype for a in args]
        dtype = str(_np.result_type(*dtypes)) if dtypes else 'float64'
        fn = self._fn
        return Variable(lambda env: fn(*[_value(a, env) for a in args]), ndim, dtype)


def _nary(binary):
    def fn(*xs):
        out = xs[0]
        for x in xs[1:]:
            out = binary(out, x)
        return out
    return fn


add = _Elemwise('add', _nary(_np.add))
mul = _Elemwise('mul', _nary(_np.multiply))
log = _Elemwise('log', _np.log)
exp = _Elemwise('exp', _np.exp)
pow = _Elemwise('pow', _np.power)
abs_ = _Elemwise('abs_', _np.abs)


def eye(n):
    return Variable(lambda env: _np.eye(int(_value(n, env))), 2, 'float64')


def tensordot(X, Y, axes):
    x_axes, y_axes = axes
    ndim = X.ndim + Y.ndim - 2 * len(x_axes)
    return Variable(
        lambda env: _np.tensordot(X._ev(env), Y._ev(env), (list(x_axes), list(y_axes))),
        ndim, str(_np.result_type(X.dtype, Y.dtype)))


def prod(values):
    values = list(values)

    def ev(env):
        out = 1
        for v in values:
            out = out * int(_value(v, env))
        return out

    return Variable(ev, 0, 'int64')


def batched_dot(X, Y):
    return Variable(lambda env: _np.matmul(X._ev(env), Y._ev(env)), 3,
                    str(_np.result_type(X.dtype, Y.dtype)))


class Diagonal(object):
    """T.Diagonal(offset, axis1, axis2)(X)  (algebra.py:1407).  Theano appends the
    diagonal axis last, like numpy.diagonal."""

    def __init__(self, offset, axis1, axis2):
        self.offset, self.axis1, self.axis2 = offset, axis1, axis2

    def __call__(self, X):
        return Variable(
            lambda env: _np.diagonal(X._ev(env), self.offset, self.axis1, self.axis2),
            X.ndim - 1, X.dtype)
