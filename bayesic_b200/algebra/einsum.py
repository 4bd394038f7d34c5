"""The general multilinear form every product-like op funnels into.

Host-side mirror of ``bayesic/algebra.py:314-508`` (``einsum`` / ``Einsum`` and its
canonical form).  Index convention is the reference's and is API surface:
each axis of each factor carries either ``('sum', n)`` (contracted) or
``('out', n)`` (axis ``n`` of the result); an out number that no factor carries
is a broadcastable (extent-1) axis of the result (``algebra.py:340-345``).

Canonical form (``algebra.py:391-508``), applied by :func:`einsum`:

1. factors that are themselves einsums are dissolved into this one;
2. ``eye`` factors touching a contracted index are removed by identifying the
   two indices they tie together;
3. a single factor whose indices are exactly ``out 0..ndim-1`` *is* the result.

Deliberate divergences from the reference (defects):

* child-einsum sum indices are renamed per *occurrence* of the child, not per
  equal child, so ``dot(X, Y) * dot(X, Y)`` keeps two independent contractions
  (the reference keys the renaming by ``(factor, index)`` at ``algebra.py:411``
  and silently ties them together);
* index tuples are always stored as tuples (the reference leaks lists from
  ``algebra.py:592`` that later crash ``Counter`` at ``algebra.py:636``).
"""
from .expr import Expression, eye, wrap_if_literal

__all__ = ['einsum', 'Einsum', 'SUM', 'OUT', 'sum_index', 'out_index']

SUM, OUT = 'sum', 'out'


def sum_index(n):
    return (SUM, n)


def out_index(n):
    return (OUT, n)


def einsum(factors_and_indices, ndim=None):
    """Build the canonical multilinear expression

        T[o0, o1, ...] = sum over all 'sum' indices of  prod_f  f[indices_f]

    ``factors_and_indices`` is a sequence of ``(factor, indices)``; ``ndim`` is the
    number of out axes (default: highest out number + 1).  Reference:
    ``algebra.py:314-347``."""
    return Einsum(factors_and_indices, ndim)._canonicalize()


def _merge_index_classes(pairs):
    """Union the given index pairs into classes; returns ``{index: class}``."""
    cls = {}
    for a, b in pairs:
        merged = cls.get(a, frozenset((a,))) | cls.get(b, frozenset((b,)))
        for member in merged:
            cls[member] = merged
    return cls


class Einsum(Expression):
    """See :func:`einsum`.  ``factors_and_indices`` is a tuple of
    ``(factor, tuple_of_indices)``."""

    def __init__(self, factors_and_indices, ndim=None):
        pairs = tuple((wrap_if_literal(f), tuple(tuple(i) for i in idx))
                      for f, idx in factors_and_indices)
        for factor, idx in pairs:
            if factor.ndim != len(idx):
                raise ValueError("The indices for each factor must have same length as factor.ndim")
        out_numbers = [n for _, idx in pairs for kind, n in idx if kind == OUT]
        if ndim is None:
            if not out_numbers:
                raise ValueError("ndim must be given when there are no out indices")
            ndim = max(out_numbers) + 1
        if any(n < 0 or n >= ndim for n in out_numbers):
            raise ValueError("some output indices are out of range")
        self.ndim = ndim
        self.factors_and_indices = pairs
        Expression.__init__(self, [f for f, _ in pairs])

    # ---- index bookkeeping ----------------------------------------------
    @property
    def out_indices(self):
        return [out_index(n) for n in range(self.ndim)]

    @property
    def sum_indices(self):
        return sorted({i for _, idx in self.factors_and_indices for i in idx if i[0] == SUM})

    def factors(self):
        return self.parents

    @classmethod
    def _wrap_if_not_einsum(cls, expr):
        """``expr`` as an einsum (identity pattern if it is not one already)."""
        if isinstance(expr, cls):
            return expr
        return cls([(expr, tuple(out_index(i) for i in range(expr.ndim)))], expr.ndim)

    # ---- canonical form ---------------------------------------------------
    def _canonicalize(self):
        return self._dissolve_child_einsums()._drop_contracted_eyes()._unwrap_identity()

    def _dissolve_child_einsums(self):
        # New numbering of contracted indices: first the private ones of each
        # child einsum (children in order), then our own in order of appearance.
        # Keeping this order is what preserves the user's bracketing later on
        # (algebra.py:539-546).
        counter = 0
        child_renames = []
        for factor, _ in self.factors_and_indices:
            rename = {}
            if isinstance(factor, Einsum):
                for s in factor.sum_indices:
                    rename[s] = sum_index(counter)
                    counter += 1
            child_renames.append(rename)
        own_rename = {}
        for _, idx in self.factors_and_indices:
            for i in idx:
                if i[0] == SUM and i not in own_rename:
                    own_rename[i] = sum_index(counter)
                    counter += 1

        flat = []
        for (factor, idx_here), rename in zip(self.factors_and_indices, child_renames):
            if isinstance(factor, Einsum):
                inner = factor.factors_and_indices
            else:
                inner = ((factor, tuple(out_index(a) for a in range(factor.ndim))),)
            for leaf, idx_in_child in inner:
                translated = []
                for kind, n in idx_in_child:
                    if kind == OUT:
                        # axis n of the child; what we call that axis
                        ours = idx_here[n]
                        translated.append(own_rename.get(ours, ours))
                    else:
                        translated.append(rename[(kind, n)])
                flat.append((leaf, tuple(translated)))
        return Einsum(flat, self.ndim)

    def _drop_contracted_eyes(self):
        tied = [(s, s) for s in self.sum_indices]
        kept = []
        for factor, idx in self.factors_and_indices:
            if isinstance(factor, eye) and (idx[0][0] == SUM or idx[1][0] == SUM):
                tied.append(idx)
            else:
                kept.append((factor, idx))
        # representative of a class: an out index if there is one, else the
        # lowest contracted index ('out' < 'sum' lexicographically).
        representative = {i: min(c) for i, c in _merge_index_classes(tied).items()}
        survivors = sorted({r[1] for r in representative.values() if r[0] == SUM})
        dense = {sum_index(old): sum_index(new) for new, old in enumerate(survivors)}

        def rename(i):
            i = representative.get(i, i)
            return dense.get(i, i)

        return Einsum([(f, tuple(rename(i) for i in idx)) for f, idx in kept], self.ndim)

    def _unwrap_identity(self):
        if len(self.factors_and_indices) == 1:
            factor, idx = self.factors_and_indices[0]
            if factor.ndim == self.ndim and idx == tuple(self.out_indices):
                return factor
        return self

    # ---- planning -----------------------------------------------------------
    def _rewrite_as_special_case_ops(self):
        """Plan-IR tree (``_tensordot/_sum/_mul/_dimshuffle/_diagonal``) computing
        this einsum; same entry point name as ``algebra.py:527-551``."""
        from .planner import plan_einsum
        return plan_einsum(self)

    # ---- printing -------------------------------------------------------------
    def __repr__(self):
        sums = self.sum_indices

        def letter(index):
            kind, n = index
            if kind == OUT:
                return 'uvwxyz'[n] if n < 6 else 'o%d' % n
            pos = sums.index(index)
            return 'ijklmn'[pos] if pos < 6 else 's%d' % pos

        def show(factor, idx):
            text = factor.bracketed_repr()
            return text if not idx else "%s_%s" % (text, ''.join(letter(i) for i in idx))

        body = ' '.join(show(f, idx) for f, idx in self.factors_and_indices) or '1'
        if sums:
            body = 'sum_%s %s' % (''.join(letter(s) for s in sums), body)
        if self.ndim > 0:
            return 'einsum(out_%s = %s)' % (''.join(letter(o) for o in self.out_indices), body)
        return 'einsum(%s)' % body

    # ---- matching / equality (implemented in matching.py) ---------------------
    def match(self, template, slot):
        from .matching import match_einsum
        return match_einsum(self, template, slot)

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, self.__class__):
            return False
        from .matching import einsums_isomorphic
        return einsums_isomorphic(self, other)

    def __hash__(self):
        from .matching import einsum_structure_hash
        return einsum_structure_hash(self)
