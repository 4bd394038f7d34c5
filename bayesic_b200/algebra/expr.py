"""Expression core: symbolic tensors with named inputs, literals, runtime dims.

Host-side Python mirror of the reference's expression API for the hot path
(reference: ``bayesic/algebra.py:17-290``).  Same public names and argument
meaning (``var``, ``constant``, ``shape``, ``elemwise``, ``add``, ``eye``), but no
Theano anywhere: an expression is lowered to a flat plan descriptor and
executed by the sm_100a kernels behind the C-ABI (see ``bayesic_b200/backend``).

Differences from the reference, all deliberate:

* ``Expression.input_types`` raises ``TypeError`` with a well-formed message
  (the reference's format string at ``algebra.py:28`` has three ``%s`` for two
  arguments and itself raises a formatting ``TypeError``).
* ``compile()`` returns ``f(**inputs)`` exactly like ``algebra.py:50-58``; the
  attribute carrying the compiled object is ``f.plan`` (``f.theano_fn`` is kept
  as an alias so reference-era call sites keep working).
"""
import numpy as np

__all__ = [
    'Expression', 'var', 'constant', 'shape', 'elemwise', 'add', 'eye', 'logdet',
    'wrap_if_literal', 'with_wrapped_literals', 'autobroadcast_or_match',
]


class Expression(object):
    """A node of the symbolic graph.  Subclasses provide ``ndim`` and either
    override ``_equality_by`` or ``__eq__``/``__hash__`` (``algebra.py:17-105``)."""

    def __init__(self, parents):
        self.parents = tuple(parents)

    # ---- typing ---------------------------------------------------------
    @property
    def input_types(self):
        """``{input name: (dtype, ndim)}`` gathered over the whole sub-graph
        (``algebra.py:21-30``).  The same name with two types is a ``TypeError``."""
        merged = {}
        for parent in self.parents:
            for name, typ in parent.input_types.items():
                seen = merged.setdefault(name, typ)
                if seen != typ:
                    raise TypeError(
                        "same input %s occurs with different types %s, %s" % (name, seen, typ))
        return merged

    # ---- execution ------------------------------------------------------
    def lower(self):
        """The expression the executor runs: ``self`` with every einsum replaced
        by its plan-IR tree.  (Reference: planning happens inside
        ``Einsum.apply``, ``algebra.py:767-768``, i.e. once per compile.)"""
        from .planner import lower_to_plan_ir
        return lower_to_plan_ir(self)

    def compile(self, **options):
        """Plan, lower and bind to the CUDA backend; returns ``f(**inputs)``
        (``algebra.py:50-58``)."""
        from ..backend.compiled import compile_expressions
        return compile_expressions([self], single=True, **options)

    # ---- printing -------------------------------------------------------
    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ', '.join(map(repr, self.parents)))

    def bracketed_repr(self):
        return repr(self)

    def terms(self):
        """``self == add(*self.terms())``."""
        return [self]

    # ---- structural equality -------------------------------------------
    def _equality_by(self):
        return self.parents

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self._equality_by() == other._equality_by()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._equality_by())

    # ---- shape helpers --------------------------------------------------
    @property
    def shape(self):
        return tuple(shape(self, axis) for axis in range(self.ndim))

    @property
    def size(self):
        from .ops import mul
        return mul(*self.shape)


class var(Expression):
    """A named input of ``ndim`` axes (``algebra.py:108-126``); dtype defaults to
    float32 like the reference."""

    def __init__(self, name, ndim, dtype='float32'):
        self.name, self.ndim, self.dtype = name, ndim, dtype
        Expression.__init__(self, ())

    @property
    def input_types(self):
        return {self.name: (self.dtype, self.ndim)}

    def __repr__(self):
        return self.name

    def _equality_by(self):
        return self.name


def _hashable_literal(value):
    if isinstance(value, np.ndarray):
        return ('ndarray', value.dtype.str, value.shape, value.tobytes())
    return value


class constant(Expression):
    """A literal scalar or ndarray (``algebra.py:129-144``)."""

    def __init__(self, value):
        self.value = value
        arr = np.asarray(value)
        self.ndim = arr.ndim
        self.dtype = str(arr.dtype)
        Expression.__init__(self, ())

    def __repr__(self):
        return repr(self.value)

    def _equality_by(self):
        # the reference compares raw values, which breaks for ndarray literals
        # (ambiguous truth value); compare array literals by content instead.
        return _hashable_literal(self.value)


class shape(Expression):
    """Runtime extent of one axis of an expression; an integer scalar
    (``algebra.py:147-161``)."""

    ndim = 0

    def __init__(self, expression, axis):
        Expression.__init__(self, (expression,))
        self.axis = axis

    def _equality_by(self):
        return (self.parents[0], self.axis)

    def __repr__(self):
        return "%s.shape[%d]" % (self.parents[0].bracketed_repr(), self.axis)


def wrap_if_literal(x):
    """Scalars and ndarrays become ``constant``; expressions pass through
    (``algebra.py:164-170``)."""
    if isinstance(x, Expression):
        return x
    if np.isscalar(x) or isinstance(x, np.ndarray):
        return constant(x)
    raise ValueError("must be a scalar, numpy array or Expression")


def with_wrapped_literals(fn):
    def wrapped(*args):
        return fn(*[wrap_if_literal(a) for a in args])
    wrapped.__name__ = getattr(fn, '__name__', 'wrapped')
    wrapped.__doc__ = fn.__doc__
    return wrapped


def autobroadcast_or_match(X, ndim):
    """Scalars are broadcast up to ``ndim``; anything else must already match --
    broadcasting is explicit via ``dimshuffle`` (``algebra.py:179-192``)."""
    if X.ndim == ndim:
        return X
    if X.ndim == 0:
        from .ops import dimshuffle
        return dimshuffle(X, *(['x'] * ndim))
    raise ValueError(
        "Dimension mismatch, was %d, should be %d. If you want broadcasting "
        "you need to do it explicitly via dimshuffle" % (X.ndim, ndim))


class ElemwiseOp(object):
    """Identity of a pointwise operator.  Plays the role the theano op object
    plays in the reference (``algebra.py:195-209``): it is what elemwise
    expressions are compared by, and it names the device opcode."""

    _registry = {}

    def __init__(self, name, arity):
        self.name, self.arity = name, arity
        ElemwiseOp._registry[name] = self

    def __repr__(self):
        return "<elemwise %s>" % self.name


OP_ADD = ElemwiseOp('add', None)
OP_MUL = ElemwiseOp('mul', None)
OP_LOG = ElemwiseOp('log', 1)
OP_EXP = ElemwiseOp('exp', 1)
OP_POW = ElemwiseOp('pow', 2)
OP_ABS = ElemwiseOp('abs_', 1)
# extension beyond the reference's vocabulary (algebra.py:1435-1448 has log/exp/pow/abs only): the
# log-normalisers of the Gamma / Dirichlet families need log Gamma
OP_LGAMMA = ElemwiseOp('lgamma', 1)


class elemwise(Expression):
    """Pointwise application of ``op`` to equal-rank arguments; scalar arguments
    are auto-broadcast (``algebra.py:195-209``)."""

    def __init__(self, op, *args, name=None):
        args = [wrap_if_literal(a) for a in args]
        self.ndim = max(a.ndim for a in args)
        self.op = op
        self.name = name or op.name
        Expression.__init__(self, [autobroadcast_or_match(a, self.ndim) for a in args])

    def __repr__(self):
        return "%s(%s)" % (self.name, ', '.join(map(repr, self.parents)))

    def _equality_by(self):
        return (self.op, self.parents)


class add(elemwise):
    """n-ary sum.  Nested sums are flattened (associativity) and equality
    ignores term order (commutativity) -- ``algebra.py:212-233``."""

    def __init__(self, *terms):
        flat = []
        for term in terms:
            flat.extend(wrap_if_literal(term).terms())
        elemwise.__init__(self, OP_ADD, *flat)

    def terms(self):
        return self.parents

    def __repr__(self):
        return ' + '.join(map(repr, self.parents))

    def bracketed_repr(self):
        return '(%r)' % self

    def _equality_by(self):
        return frozenset(self.parents)


class eye(Expression):
    """Square identity whose extent is any of the given (runtime-equal) scalar
    expressions (``algebra.py:236-290``).  Two ``eye``s are equal when their
    extent sets intersect, so the hash can only be the class."""

    ndim = 2

    def __init__(self, *shapes):
        if not shapes:
            raise ValueError("need at least one shape for eye")
        Expression.__init__(self, [wrap_if_literal(s) for s in shapes])

    def __eq__(self, other):
        return isinstance(other, self.__class__) and bool(set(self.parents) & set(other.parents))

    def __hash__(self):
        return hash(self.__class__)


class logdet(Expression):
    """``log|X|`` over the last two axes of a stack of symmetric positive-definite matrices:
    ``X[..., d, d] -> [...]``.  Not in the reference's executable vocabulary: its
    ``MultivariateNormal.log_normalizer`` calls a ``T.logdet`` that Theano never had
    (``bayesic/distribution/core.py:49-52``).  Needed by every log-normaliser with a precision or scale
    matrix in it (MVN, Wishart, Gaussian-Wishart); evaluated on the device by a float64 Cholesky
    (``BB_NODE_LOGDET``).  Not multilinear, so it stays an opaque node, like ``log``."""

    def __init__(self, X):
        X = wrap_if_literal(X)
        if X.ndim < 2:
            raise ValueError("logdet needs at least two axes, got %d" % X.ndim)
        self.ndim = X.ndim - 2
        Expression.__init__(self, (X,))
