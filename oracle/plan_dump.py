"""Neutral (JSON-able) dump of a planned expression, duck-typed over the
reference's node classes and ``bayesic_b200.algebra``'s, so the two planners can
be compared structurally without relying on ``repr`` (TEST INFRASTRUCTURE)."""
import numpy as np


def dump_plan(expr):
    kind = type(expr).__name__
    if kind == 'Einsum':
        return dump_plan(expr._rewrite_as_special_case_ops())
    if kind == 'var':
        return {'op': 'var', 'name': expr.name}
    if kind == 'constant':
        return {'op': 'constant', 'value': np.asarray(expr.value, dtype='float64').tolist()}
    if kind == 'shape':
        return {'op': 'shape', 'axis': expr.axis, 'of': dump_plan(expr.parents[0])}
    if kind == 'eye':
        return {'op': 'eye', 'extents': [dump_plan(p) for p in expr.parents]}
    if kind in ('elemwise', 'add'):
        return {'op': 'elemwise', 'name': 'add' if kind == 'add' else expr.name,
                'args': [dump_plan(p) for p in expr.parents]}
    if kind == '_sum':
        return {'op': '_sum', 'axes': sorted(expr.axes), 'x': dump_plan(expr.parents[0])}
    if kind == '_mul':
        return {'op': '_mul', 'factors': [dump_plan(p) for p in expr.parents]}
    if kind == '_dimshuffle':
        return {'op': '_dimshuffle', 'axes': list(expr.axes), 'x': dump_plan(expr.parents[0])}
    if kind == '_diagonal':
        return {'op': '_diagonal', 'axes': sorted([expr.axis1, expr.axis2]),
                'x': dump_plan(expr.parents[0])}
    if kind == '_tensordot':
        return {'op': '_tensordot',
                'x': dump_plan(expr.parents[0]), 'y': dump_plan(expr.parents[1]),
                'x_dot': list(expr.X_dot_axes), 'y_dot': list(expr.Y_dot_axes),
                'x_batch': list(expr.X_batch_axes), 'y_batch': list(expr.Y_batch_axes)}
    raise TypeError('plan_dump: unknown node type %s' % kind)
