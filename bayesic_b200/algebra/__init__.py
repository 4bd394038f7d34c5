"""``bayesic.algebra``-compatible symbolic tensor algebra (host side, Theano-free).

``from bayesic_b200.algebra import *`` exposes the same names as
``from bayesic.algebra import *`` in the reference (``bayesic/algebra.py``);
``expr.compile()`` returns ``f(**inputs)`` backed by the sm_100a executor.
"""
from collections import Counter, defaultdict  # noqa: F401  (reference re-exports these via *)

from .expr import (Expression, var, constant, shape, elemwise, add, eye, logdet,  # noqa: F401
                   wrap_if_literal, with_wrapped_literals, autobroadcast_or_match)
from .einsum import einsum, Einsum  # noqa: F401
from .plan_ir import _sum, _mul, _dimshuffle, _tensordot, _diagonal  # noqa: F401
from .planner import plan_einsum, lower_to_plan_ir  # noqa: F401
from .matching import (match, find_duplicate, equivalence_classes, submultisets_of_size,  # noqa: F401
                       find_injection, find_injections, find_bijection, find_bijections)
from .ops import (dot, tensordot, mul, outer, sum, trace, diagonal, transpose, dimshuffle,  # noqa: F401
                  div, neg, sub, log, exp, pow, abs_, lgamma)

__all__ = [
    'Expression', 'var', 'constant', 'shape', 'elemwise', 'add', 'eye',
    'wrap_if_literal', 'with_wrapped_literals', 'autobroadcast_or_match',
    'einsum', 'Einsum', 'match',
    'find_duplicate', 'equivalence_classes', 'submultisets_of_size',
    'find_injection', 'find_injections', 'find_bijection', 'find_bijections',
    'dot', 'tensordot', 'mul', 'outer', 'sum', 'trace', 'diagonal', 'transpose', 'dimshuffle',
    'div', 'neg', 'sub', 'log', 'exp', 'pow', 'abs_', 'lgamma', 'logdet',
    'Counter', 'defaultdict',
]
