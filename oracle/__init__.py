"""CPU oracle for the hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; nothing under
``bayesic_b200/`` does (``tests/test_no_oracle_in_product.py`` enforces it).

Contents
--------
``semantics.py``        float64 numpy evaluation of an expression tree by its
                        *declared* semantics (einsum index patterns, plan-IR node
                        docs) -- restates ``bayesic/algebra.py:314-346`` and
                        ``:1161-1171, 1284-1414``.
``descriptor_eval.py``  numpy evaluation of the flat plan descriptor that the
                        C-ABI executor consumes, node by node.
``closed_forms.py``     float64 restatements of the exponential-family
                        quantities (``bayesic/distribution/base.py:25-100,
                        263-335``; ``core.py:8-55``) and of the north-star's
                        log-sum-exp / ELBO terms.
``theano_shim/``        numpy stand-in for Theano so the unmodified reference
                        module imports in the authoring container.
``reference_loader.py`` imports ``/root/reference/bayesic/algebra.py`` through the
                        shim (authoring container only).
``make_golden.py``      regenerates ``tests/golden/`` from the reference.

Pinning: the oracle is pinned against the reference itself -- every numeric test
of ``bayesic/tests/test_algebra.py:44-191`` and every plan-shape test
``:376-505`` is replayed through the unmodified reference (via the shim) by
``make_golden.py``; outputs and plans are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks ``semantics.evaluate`` against them.
PARITY UNPINNED (no reference output exists: the reference cannot evaluate
them): batched ``_tensordot`` numerics, anything in
``bayesic/distribution/``, log-sum-exp.  For those the oracle is the float64
restatement of the declared semantics and says so where used.
"""
