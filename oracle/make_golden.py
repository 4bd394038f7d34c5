"""Regenerate ``tests/golden/algebra_golden.{json,npz}`` from the UNMODIFIED
reference (``/root/reference/bayesic/algebra.py`` imported through the numpy Theano
shim).  Run in the authoring container:  ``python -m oracle.make_golden``.

For every case of ``tests/golden/cases.py`` it records
  * ``repr`` of the canonical expression,
  * the reference planner's plan as a neutral nested structure (or the exception
    the reference raises -- it crashes on some inputs),
  * the reference's own numeric output ``expr.compile()(**inputs)`` (or the
    exception: batched ``_tensordot`` evaluation is broken in the reference),
  * the float64 declared-semantics value (``oracle.semantics.evaluate``).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference_algebra  # noqa: E402
from oracle.semantics import evaluate  # noqa: E402
from oracle.plan_dump import dump_plan  # noqa: E402
from tests.golden.cases import CASES, make_inputs  # noqa: E402


def main():
    ref = load_reference_algebra()
    inputs = make_inputs()
    meta, arrays = {}, {}
    for name, build in CASES:
        entry = {}
        expr = build(ref)
        entry['repr'] = repr(expr)
        entry['ndim'] = expr.ndim
        try:
            entry['plan'] = dump_plan(expr)
        except Exception as exc:       # reference planner defects
            entry['plan'] = None
            entry['plan_error'] = '%s: %s' % (type(exc).__name__, exc)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                fn = expr.compile()
                used = {k: inputs[k] for k in expr.input_types}
                out = fn(**used)
            if out is None:
                raise ValueError('reference evaluator returned None')
            arrays[name + '__ref'] = np.asarray(out)
            entry['ref_output'] = True
        except Exception as exc:       # reference evaluator defects
            entry['ref_output'] = False
            entry['ref_error'] = '%s: %s' % (type(exc).__name__, str(exc)[:200])
        arrays[name + '__f64'] = np.asarray(evaluate(expr, inputs))
        if entry['ref_output']:
            # does the reference's own evaluator agree with the declared semantics?
            # (it does not for batched contractions: algebra.py:1380 feeds X twice)
            entry['ref_matches_declared'] = bool(
                arrays[name + '__ref'].shape == arrays[name + '__f64'].shape and
                np.allclose(arrays[name + '__ref'], arrays[name + '__f64'], rtol=1e-4, atol=1e-5,
                            equal_nan=True))
        meta[name] = entry
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    with open(os.path.join(out_dir, 'algebra_golden.json'), 'w') as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(out_dir, 'algebra_golden.npz'), **arrays)
    n_plan = sum(1 for e in meta.values() if e['plan'] is not None)
    n_out = sum(1 for e in meta.values() if e['ref_output'])
    print('cases: %d, reference plans: %d, reference outputs: %d' % (len(meta), n_plan, n_out))
    for name, e in meta.items():
        if e['plan'] is None or not e['ref_output'] or not e.get('ref_matches_declared', True):
            print('  %-22s plan_error=%s ref_error=%s ref_matches_declared=%s' % (
                name, e.get('plan_error'), e.get('ref_error'), e.get('ref_matches_declared')))


if __name__ == '__main__':
    main()
