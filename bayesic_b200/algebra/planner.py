"""Einsum planner: lowers a canonical ``Einsum`` to a tree of plan-IR nodes.

Restates the strategy of ``bayesic/algebra.py:513-765`` so that plans come out
identical to the reference's (``tests/test_plan_parity.py`` pins this against
plans dumped from the unmodified reference module):

* a factor carrying the same index twice is replaced by its ``_diagonal`` first
  (``algebra.py:513-525``);
* contracted indices are eliminated lowest-number first, so nested ``dot``s keep
  the bracketing the user wrote (``algebra.py:539-584``); indices that occur on
  exactly the same factors go together;
* an index living on a single factor becomes a ``_sum`` pushed inside the product
  (``algebra.py:585-600``) -- this is how ``dot(X, eta).sum()`` turns into the
  sufficient-statistic form ``_tensordot(_sum(X), eta)`` by itself;
* an index shared by several factors becomes a ``_tensordot``.  The factors are
  split into two operands greedily, minimising
  ``#factors * #distinct indices`` summed over both sides
  (``algebra.py:602-654``); non-contracted indices present on both sides are
  batch axes (``algebra.py:668-671``);
* with nothing left to contract the factors are aligned by ``_dimshuffle`` and
  multiplied (``algebra.py:741-765``).

No shapes are known here; the plan is shape-polymorphic (``algebra.py:539-546``).

Reference quirk kept on purpose (plan parity): the greedy split compares every
candidate move against the cost of the *initial* split, never the current one
(``algebra.py:648-654`` does not update ``current_cost``).

Reference defect not kept: ``algebra.py:592`` builds list-typed index sequences
which crash ``Counter`` at ``:636`` on the next elimination step (e.g.
``dot(sum(X, 0), dot(L, m))``); indices are tuples throughout here.
"""
import copy
from collections import Counter

from .expr import constant
from .einsum import Einsum, OUT, out_index
from .plan_ir import _sum, _mul, _dimshuffle, _tensordot, _diagonal

__all__ = ['plan_einsum', 'lower_to_plan_ir']


def plan_einsum(es):
    """Plan-IR tree for one einsum (its factors are left as they are)."""
    return _eliminate_contractions(_diagonalise_repeats(es))


def lower_to_plan_ir(expr, _memo=None):
    """Replace every einsum in ``expr`` (at any depth) by its plan; all other node
    types are kept, rebuilt over lowered parents."""
    memo = {} if _memo is None else _memo
    key = id(expr)
    if key in memo:
        return memo[key][1]
    node = plan_einsum(expr) if isinstance(expr, Einsum) else expr
    if node.parents:
        lowered = tuple(lower_to_plan_ir(p, memo) for p in node.parents)
        if any(a is not b for a, b in zip(lowered, node.parents)):
            node = copy.copy(node)
            node.parents = lowered
    memo[key] = (expr, node)       # keep expr alive so id() stays unique
    return node


# ---------------------------------------------------------------------------

def _first_repeat(idx):
    seen = {}
    for pos, i in enumerate(idx):
        if i in seen:
            return pos, seen[i], i
        seen[i] = pos
    return None


def _diagonalise_repeats(es):
    rewritten = []
    for factor, idx in es.factors_and_indices:
        while True:
            hit = _first_repeat(idx)
            if hit is None:
                break
            later, earlier, index = hit
            factor = _diagonal(factor, later, earlier)
            idx = tuple(i for pos, i in enumerate(idx) if pos not in (later, earlier)) + (index,)
        rewritten.append((factor, idx))
    return Einsum(rewritten, es.ndim)


def _split_cost(side):
    distinct = {i for _, idx in side for i in idx}
    return len(side) * len(distinct)


def _greedy_operand_split(carriers):
    """Split the factors carrying the contracted index into (lhs, rhs) lists."""
    lhs = list(carriers)
    rhs = Counter([lhs.pop()])
    lhs = Counter(lhs)
    threshold = _split_cost(lhs) + _split_cost(rhs)
    while len(lhs) > 1:
        best = None
        for candidate in lhs:
            moved = Counter([candidate])
            trial = (lhs - moved, rhs + moved)
            cost = _split_cost(trial[0]) + _split_cost(trial[1])
            if best is None or cost < best[1]:
                best = (trial, cost)
        if best[1] >= threshold:
            break
        lhs, rhs = best[0]
    return list(lhs.elements()), list(rhs.elements())


def _eliminate_contractions(es):
    sums = es.sum_indices
    if not sums:
        return _aligned_product(es)
    pairs = es.factors_and_indices

    def carriers_of(index):
        return [pos for pos, (_, idx) in enumerate(pairs) if index in idx]

    lead = sums[0]
    where = carriers_of(lead)
    group = [s for s in sums if carriers_of(s) == where]

    if len(where) == 1:
        pos = where[0]
        factor, idx = pairs[pos]
        summed = _sum(factor, *[idx.index(s) for s in group])
        replaced = list(pairs)
        replaced[pos] = (summed, tuple(i for i in idx if i not in group))
        return _eliminate_contractions(Einsum(replaced, es.ndim))

    lhs, rhs = _greedy_operand_split([pairs[pos] for pos in where])
    rest = [pairs[pos] for pos in range(len(pairs)) if pos not in where]
    used_outside = {i for _, idx in rest for i in idx}
    on_lhs = {i for _, idx in lhs for i in idx}
    on_rhs = {i for _, idx in rhs for i in idx}
    shared = on_lhs & on_rhs
    batch = sorted(shared - set(group))

    def operand(side, present, contracted_first):
        # axes the operand must expose: result axes, anything another factor
        # still needs, and the batch axes; contracted axes last on the lhs and
        # first on the rhs (the usual dot convention).
        exposed = sorted(i for i in present
                         if i[0] == OUT or i in used_outside or i in batch)
        exposed = group + exposed if contracted_first else exposed + group
        axis_of = {i: n for n, i in enumerate(exposed)}
        inner = Einsum([(f, tuple(out_index(axis_of[i]) if i in axis_of else i for i in idx))
                        for f, idx in side], len(exposed))
        return (_eliminate_contractions(inner),
                [axis_of[s] for s in group],
                [axis_of[b] for b in batch],
                [i for i in exposed if i not in shared])

    x_plan, x_dot, x_batch, x_other = operand(lhs, on_lhs, False)
    y_plan, y_dot, y_batch, y_other = operand(rhs, on_rhs, True)
    product = _tensordot(x_plan, y_plan, x_dot, y_dot, x_batch, y_batch)
    product_idx = tuple(batch + x_other + y_other)

    # Put the contraction roughly where its factors were, so later steps see the
    # operands in the order the user wrote them (algebra.py:722-739).
    original = list(pairs)
    consumed = lhs + rhs
    centre = sum(original.index(fi) for fi in consumed) / float(len(consumed))
    placed = [(fi, original.index(fi)) for fi in rest] + [((product, product_idx), centre)]
    placed.sort(key=lambda item: item[1])
    return _eliminate_contractions(Einsum([fi for fi, _ in placed], es.ndim))


def _aligned_product(es):
    outs = es.out_indices
    aligned = []
    for factor, idx in es.factors_and_indices:
        axes = [idx.index(o) if o in idx else 'x' for o in outs]
        aligned.append(factor if axes == list(range(factor.ndim)) else _dimshuffle(factor, *axes))
    if not aligned:
        one = constant(1)
        return one if es.ndim == 0 else _dimshuffle(one, *(['x'] * es.ndim))
    if len(aligned) == 1:
        return aligned[0]
    return _mul(*aligned)
