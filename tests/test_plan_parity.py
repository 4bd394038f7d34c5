"""Our planner must emit the same plan as the unmodified reference planner
(``bayesic/algebra.py:513-765``) for every golden case the reference can plan."""
import json

import pytest

import bayesic_b200.algebra as A
from oracle.plan_dump import dump_plan
from tests.golden.cases import CASES
from tests.golden_util import load_algebra_golden

META, _ = load_algebra_golden()


def _canon(plan):
    """_mul factors and _sum/_diagonal axes compare as sets in the reference
    (algebra.py:1293-1294, 1308-1309, 1413-1414)."""
    if isinstance(plan, dict):
        out = {k: _canon(v) for k, v in plan.items()}
        if out.get('op') == '_mul':
            out['factors'] = sorted(out['factors'], key=lambda f: json.dumps(f, sort_keys=True))
        return out
    if isinstance(plan, list):
        return [_canon(p) for p in plan]
    return plan


@pytest.mark.parametrize('name,build', CASES, ids=[c[0] for c in CASES])
def test_plan_matches_reference(name, build):
    entry = META[name]
    expr = build(A)
    assert repr(expr) == entry['repr']
    assert expr.ndim == entry['ndim']
    ours = dump_plan(expr)
    if entry['plan'] is None:
        pytest.skip('reference planner crashes here: %s' % entry['plan_error'])
    assert _canon(ours) == _canon(entry['plan'])


def test_planner_survives_the_reference_crash_case():
    D, L, eta = A.var('D', 2), A.var('L', 2), A.var('eta', 1)
    plan = A.dot(A.sum(D, 0), A.dot(L, eta))._rewrite_as_special_case_ops()
    assert plan == A._tensordot(A._sum(D, 0), A._tensordot(L, eta, [1], [0]), [0], [0])
