"""Plan IR: the five special-purpose nodes an einsum is lowered to before it is
executed (reference: ``bayesic/algebra.py:1280-1414``).

In the reference each node's ``_apply_to_parents`` emits a Theano op.  Here the
nodes are pure descriptions; ``bayesic_b200.backend.lowering`` flattens a tree of
them into the plan descriptor that ``bb_plan_create`` (include/bayesic_b200.h)
consumes, and the sm_100a kernels do the arithmetic.

Semantics the executor implements (and the oracle restates in numpy):

``_sum(X, *axes)``            X.sum(axis=axes)                       algebra.py:1284-1294
``_mul(*factors)``            broadcasting product, equal ranks       algebra.py:1297-1309
``_dimshuffle(X, *axes)``     permute axes / insert 'x' extent-1 axes algebra.py:1312-1326
``_tensordot(X, Y, ...)``     contraction; result axes are
                              batch + X_other + Y_other               algebra.py:1161-1171, 1329-1396
``_diagonal(X, a1, a2)``      diagonal of two axes, appended last     algebra.py:1398-1414

The reference's *evaluator* for batched ``_tensordot`` is broken
(``algebra.py:1370-1373`` returns ``None``; ``:1380`` uses ``X`` for ``Y``); the
declared semantics above are what is implemented.
"""
from .expr import Expression

__all__ = ['_sum', '_mul', '_dimshuffle', '_tensordot', '_diagonal']


class _sum(Expression):
    def __init__(self, X, *axes):
        self.axes = tuple(axes)
        self.ndim = X.ndim - len(self.axes)
        Expression.__init__(self, (X,))

    def _equality_by(self):
        return (self.parents[0], frozenset(self.axes))

    def __repr__(self):
        return "_sum(%r%s)" % (self.parents[0], ''.join(', %d' % a for a in self.axes))


class _mul(Expression):
    def __init__(self, *factors):
        self.ndim = factors[0].ndim
        Expression.__init__(self, factors)

    def _equality_by(self):
        return frozenset(self.parents)


class _dimshuffle(Expression):
    def __init__(self, X, *axes):
        self.axes = tuple(axes)
        self.ndim = len(self.axes)
        Expression.__init__(self, (X,))

    def _equality_by(self):
        return (self.parents[0], self.axes)

    def __repr__(self):
        return "_dimshuffle(%r, %s)" % (self.parents[0], ', '.join(map(repr, self.axes)))


class _tensordot(Expression):
    def __init__(self, X, Y, X_dot_axes, Y_dot_axes, X_batch_axes=(), Y_batch_axes=()):
        self.X_dot_axes = list(X_dot_axes)
        self.Y_dot_axes = list(Y_dot_axes)
        self.X_batch_axes = list(X_batch_axes)
        self.Y_batch_axes = list(Y_batch_axes)
        taken_x = set(self.X_dot_axes) | set(self.X_batch_axes)
        taken_y = set(self.Y_dot_axes) | set(self.Y_batch_axes)
        self.X_other_axes = [a for a in range(X.ndim) if a not in taken_x]
        self.Y_other_axes = [a for a in range(Y.ndim) if a not in taken_y]
        self.ndim = len(self.X_batch_axes) + len(self.X_other_axes) + len(self.Y_other_axes)
        Expression.__init__(self, (X, Y))

    def _equality_by(self):
        return (self.parents,
                frozenset(zip(self.X_dot_axes, self.Y_dot_axes)),
                frozenset(zip(self.X_batch_axes, self.Y_batch_axes)))

    def __repr__(self):
        parts = [repr(self.parents[0]), repr(self.parents[1]),
                 repr(self.X_dot_axes), repr(self.Y_dot_axes)]
        if self.X_batch_axes:
            parts += [repr(self.X_batch_axes), repr(self.Y_batch_axes)]
        return "_tensordot(%s)" % ', '.join(parts)


class _diagonal(Expression):
    def __init__(self, X, axis1, axis2):
        self.axis1, self.axis2 = axis1, axis2
        self.ndim = X.ndim - 1
        Expression.__init__(self, (X,))

    def _equality_by(self):
        return (self.parents[0], frozenset((self.axis1, self.axis2)))

    def __repr__(self):
        return "_diagonal(%r, %s, %s)" % (self.parents[0], self.axis1, self.axis2)
