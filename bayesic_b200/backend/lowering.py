"""Plan-IR tree -> flat plan descriptor consumed by ``bb_plan_create``.

The tree comes from ``Expression.lower()`` (every einsum replaced by its plan,
``bayesic/algebra.py:527-765``).  Lowering walks it once, de-duplicates shared
sub-trees (by object identity and by value-numbering the emitted descriptor rows -- NOT by the
expressions' ``==``, which follows the reference in ignoring repeated factors / terms and
batch-axis order, ``algebra.py:1297-1309``, and would merge ``X*X*Y`` with ``X*Y``), assigns input slots, and recognises three
patterns that the executor serves with fused kernels instead of the reference's
node-by-node evaluation (``algebra.py:34-40``):

* ``_tensordot(_dimshuffle(X,1,0), X, [1],[0])``  ->  SYRK (Sigma x x^T over the
  data axis; tcgen05 kernel when D <= 64);
* ``add(Lg, _mul(_dimshuffle(-1,'x','x'), _dimshuffle(log(_sum(exp(Lg), 1)), 0, 'x')))``
  -> LOGSOFTMAX, the only way the reference's vocabulary can spell mixture
  log-responsibilities (``algebra.py:1435-1448``); computed max-subtracted in one pass;
* ``_tensordot(_mul(_dimshuffle(R,1,'x',0), _dimshuffle(X,'x',1,0)), X, [2],[0])``
  -> WEIGHTED_SCATTER (Sigma r x x^T without the K x D x N intermediate).

Each fused node computes exactly the value of the sub-tree it replaces; fusion can
be switched off (``fuse=False``) and the parity tests run both ways.
"""
import numpy as np

from ..algebra.expr import var, constant, shape, elemwise, add, eye, logdet
from ..algebra.plan_ir import _sum, _mul, _dimshuffle, _tensordot, _diagonal
from . import library as L


class LoweredPlan(object):
    """Flat descriptor + the bookkeeping ``CompiledPlan`` needs."""

    def __init__(self):
        self.nodes = []            # list of dicts: kind, parents, iparams, fparam
        self.outputs = []          # node index per requested expression
        self.input_names = []      # slot -> var name ('' for bound constants)
        self.input_types = {}      # name -> (dtype, ndim)
        self.bound_constants = {}  # slot -> ndarray literal
        self.integer_result = []   # per output: result is integer-valued by construction

    def add(self, kind, parents=(), iparams=(), fparam=0.0):
        if len(parents) > L.MAX_PARENTS:
            raise ValueError("plan node with %d parents (max %d)" % (len(parents), L.MAX_PARENTS))
        if len(iparams) > L.MAX_IPARAMS:
            raise ValueError("plan node with %d integer params (max %d)" % (len(iparams), L.MAX_IPARAMS))
        self.nodes.append({'kind': kind, 'parents': list(parents), 'iparams': [int(i) for i in iparams],
                           'fparam': float(fparam)})
        return len(self.nodes) - 1

    def as_ctypes(self):
        arr = (L.NodeDesc * len(self.nodes))()
        for desc, node in zip(arr, self.nodes):
            desc.kind = node['kind']
            desc.n_parents = len(node['parents'])
            for i, p in enumerate(node['parents']):
                desc.parents[i] = p
            desc.n_iparams = len(node['iparams'])
            for i, v in enumerate(node['iparams']):
                desc.iparams[i] = v
            desc.fparam = node['fparam']
        return arr


def _is_scalar_literal(node, value):
    return isinstance(node, constant) and np.ndim(node.value) == 0 and node.value == value


def _match_syrk(node, same):
    """X if node is ``_tensordot(_dimshuffle(X,1,0), X, [1],[0])`` (or the mirrored
    ``_tensordot(X', X, [0],[0])`` forms), else None."""
    if not isinstance(node, _tensordot) or node.X_batch_axes:
        return None
    a, b = node.parents
    if (isinstance(a, _dimshuffle) and a.axes == (1, 0) and a.parents[0].ndim == 2
            and node.X_dot_axes == [1] and node.Y_dot_axes == [0] and same(a.parents[0], b)):
        return b
    if a.ndim == 2 and node.X_dot_axes == [0] and node.Y_dot_axes == [0] and same(a, b):
        return b
    return None


def _match_weighted_scatter(node, same):
    """(R, X) if node is the reference plan of ``sum_n R[n,k] X[n,d] X[n,e]``."""
    if not isinstance(node, _tensordot) or node.X_batch_axes:
        return None
    if node.X_dot_axes != [2] or node.Y_dot_axes != [0]:
        return None
    lhs, x = node.parents
    if not isinstance(lhs, _mul) or len(lhs.parents) != 2 or x.ndim != 2:
        return None
    found_r, found_x = None, None
    for factor in lhs.parents:
        if not isinstance(factor, _dimshuffle) or factor.parents[0].ndim != 2:
            return None
        if factor.axes == (1, 'x', 0):
            found_r = factor.parents[0]
        elif factor.axes == ('x', 1, 0):
            found_x = factor.parents[0]
    if found_r is None or found_x is None or not same(found_x, x):
        return None
    return found_r, x


def _match_logsoftmax(node, same):
    """Lg if node is ``Lg + (-1 * log(sum(exp(Lg), axis=last))) broadcast back``."""
    if not isinstance(node, add) or len(node.parents) != 2:
        return None
    for lg, corr in (node.parents, node.parents[::-1]):
        if lg.ndim != 2 or not isinstance(corr, _mul) or len(corr.parents) != 2:
            continue
        minus_one, logsum = None, None
        for factor in corr.parents:
            if not isinstance(factor, _dimshuffle):
                break
            inner = factor.parents[0]
            if factor.axes == ('x', 'x') and _is_scalar_literal(inner, -1):
                minus_one = factor
            elif factor.axes == (0, 'x') and isinstance(inner, elemwise) and inner.name == 'log':
                logsum = inner.parents[0]
        if minus_one is None or logsum is None:
            continue
        if not isinstance(logsum, _sum) or logsum.axes != (1,):
            continue
        ex = logsum.parents[0]
        if isinstance(ex, elemwise) and ex.name == 'exp' and same(ex.parents[0], lg):
            return lg
    return None


def _is_integer_valued(node):
    """True when the value is an integer by construction (shape arithmetic, integer
    literals, integer-typed inputs) -- mirrors theano's dtype upcasting closely enough
    for ``X.shape[0]`` / ``X.size`` / ``add(1, 1)`` (test_algebra.py:59-60, 167-173)."""
    if isinstance(node, shape):
        return True
    if isinstance(node, constant):
        return np.asarray(node.value).dtype.kind in 'iub'
    if isinstance(node, var):
        return np.dtype(node.dtype).kind in 'iub'
    if isinstance(node, (_mul, add, _dimshuffle, _sum, _diagonal)):
        return all(_is_integer_valued(p) for p in node.parents)
    if isinstance(node, elemwise):
        return node.name in ('add', 'mul', 'abs_') and all(_is_integer_valued(p) for p in node.parents)
    return False


def lower_plans(plan_trees, input_types, fuse=True):
    """Flatten the given lowered expression trees into one descriptor."""
    out = LoweredPlan()
    out.input_types = dict(input_types)
    slot_of_var = {}
    keep_alive = []    # ids in by_id stay valid while the nodes are referenced
    by_id = {}         # id(Expression) -> node index (shared sub-trees)
    numbered = {}      # (kind, ordered parents, iparams, fparam) -> node index

    def input_slot(name):
        if name not in slot_of_var:
            slot_of_var[name] = len(out.input_names)
            out.input_names.append(name)
        return slot_of_var[name]

    def add_node(kind, parents=(), iparams=(), fparam=0.0, unique=False):
        """Value numbering: two rows merge only when kind, the ORDERED parent list, every integer
        parameter and the immediate agree -- so X*X*Y and X*Y, or tensordots that differ in
        batch-axis order, stay distinct."""
        key = (kind, tuple(parents), tuple(int(i) for i in iparams), float(fparam))
        if not unique and key in numbered:
            return numbered[key]
        idx = out.add(kind, parents, iparams, fparam)
        if not unique:
            numbered[key] = idx
        return idx

    def chain(kind, parents, iparams):
        """n-ary node with more parents than the descriptor holds: fold left."""
        parents = list(parents)
        while len(parents) > L.MAX_PARENTS:
            head = add_node(kind, parents[:L.MAX_PARENTS], iparams)
            parents = [head] + parents[L.MAX_PARENTS:]
        return add_node(kind, parents, iparams)

    def visit(node):
        if id(node) in by_id:
            return by_id[id(node)]
        idx = emit(node)
        by_id[id(node)] = idx
        keep_alive.append(node)
        return idx

    def same(a, b):
        """Strict equality of two sub-trees: they lower to the same descriptor row."""
        return visit(a) == visit(b)

    def emit(node):
        if fuse:
            x = _match_syrk(node, same)
            if x is not None:
                return add_node(L.NODE_SYRK, [visit(x)])
            rx = _match_weighted_scatter(node, same)
            if rx is not None:
                return add_node(L.NODE_WEIGHTED_SCATTER, [visit(rx[0]), visit(rx[1])])
            lg = _match_logsoftmax(node, same)
            if lg is not None:
                return add_node(L.NODE_LOGSOFTMAX, [visit(lg)])
        if isinstance(node, var):
            return add_node(L.NODE_INPUT, iparams=[input_slot(node.name)])
        if isinstance(node, constant):
            value = np.asarray(node.value)
            if value.ndim == 0:
                return add_node(L.NODE_SCALAR, fparam=float(value))
            slot = len(out.input_names)
            out.input_names.append('')
            out.bound_constants[slot] = np.ascontiguousarray(value, dtype=np.float32)
            return add_node(L.NODE_INPUT, iparams=[slot])
        if isinstance(node, shape):
            return add_node(L.NODE_SHAPE, [visit(node.parents[0])], [node.axis])
        if isinstance(node, eye):
            return add_node(L.NODE_EYE, [visit(node.parents[0])])
        if isinstance(node, logdet):
            return add_node(L.NODE_LOGDET, [visit(node.parents[0])])
        if isinstance(node, _sum):
            return add_node(L.NODE_SUM, [visit(node.parents[0])], sorted(node.axes))
        if isinstance(node, _mul):
            return chain(L.NODE_MUL, [visit(p) for p in node.parents], [])
        if isinstance(node, _dimshuffle):
            axes = [-1 if a == 'x' else a for a in node.axes]
            return add_node(L.NODE_DIMSHUFFLE, [visit(node.parents[0])], axes)
        if isinstance(node, _diagonal):
            return add_node(L.NODE_DIAGONAL, [visit(node.parents[0])], [node.axis1, node.axis2])
        if isinstance(node, _tensordot):
            params = ([len(node.X_dot_axes), len(node.X_batch_axes)] + node.X_dot_axes +
                      node.Y_dot_axes + node.X_batch_axes + node.Y_batch_axes)
            return add_node(L.NODE_TENSORDOT, [visit(node.parents[0]), visit(node.parents[1])], params)
        if isinstance(node, elemwise):        # includes add
            name = 'add' if isinstance(node, add) else node.name
            if name not in L.OP_CODES:
                raise NotImplementedError("no device opcode for elementwise op %r" % name)
            parents = [visit(p) for p in node.parents]
            if name in ('add', 'mul'):
                return chain(L.NODE_ELEMWISE, parents, [L.OP_CODES[name]])
            return add_node(L.NODE_ELEMWISE, parents, [L.OP_CODES[name]])
        raise NotImplementedError("cannot lower %s to the plan descriptor" % type(node).__name__)

    for tree in plan_trees:
        out.outputs.append(visit(tree))
        out.integer_result.append(_is_integer_valued(tree))
    # inputs never referenced by the lowered trees still get a slot so callers may pass them
    for name in sorted(out.input_types):
        input_slot(name)
    return out
