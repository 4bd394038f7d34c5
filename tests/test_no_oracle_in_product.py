"""The oracle is test infrastructure: nothing under bayesic_b200/ may import, call or execute
anything under oracle/, and the product must fail loudly when the CUDA path is unavailable."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_sources():
    for base, _, files in os.walk(os.path.join(ROOT, 'bayesic_b200')):
        if os.sep + 'build' in base:
            continue
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h')):
                yield os.path.join(base, name)


def test_product_never_touches_the_oracle():
    pattern = re.compile(r'^\s*(from|import)\s+oracle\b|oracle[./](semantics|closed_forms|descriptor_eval)',
                         re.M)
    offenders = [p for p in _product_sources() if pattern.search(open(p).read())]
    assert offenders == []


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present; the fallback check is for CPU-only boxes')
    import bayesic_b200.algebra as A
    import bayesic_b200.stats as S
    X = A.var('X', 2)
    fn = A.dot(X.T, X).compile()                # planning + descriptor creation are host-only
    with pytest.raises(RuntimeError):
        fn(X=np.ones((4, 4), dtype='float32'))  # ... but evaluation must refuse to run on CPU
    with pytest.raises(RuntimeError):
        S.gaussian_suffstats(np.ones((4, 4), dtype='float32'))
    # the layers built on top refuse likewise: update kernels, streaming, the passes
    import bayesic_b200.passes as P
    import bayesic_b200.streaming as T
    import bayesic_b200.updates as U
    cpu = torch.zeros(2, 3, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        U.gmm_global_update(torch.zeros(2, dtype=torch.float64), cpu, torch.zeros(2, 3, 3, dtype=torch.float64),
                            1.0, 1.0, 4.0, torch.zeros(3, dtype=torch.float64), torch.eye(3, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        U.svi_natural_blend(cpu, cpu, cpu, 1.0, 0.5)
    with pytest.raises(RuntimeError):
        T.streamed_pass(S.regression_suffstats, (np.ones((4, 4), dtype='float32'), np.ones(4, dtype='float32')), 2)
    with pytest.raises(RuntimeError):
        P.gaussian_pass(np.ones((4, 4), dtype='float32'), np.eye(4), np.zeros(4), 0.0, 0.0)


def test_library_exports_every_declared_symbol():
    from bayesic_b200.backend import library
    header = open(os.path.join(ROOT, 'include', 'bayesic_b200.h')).read()
    declared = set(re.findall(r'BB_API\s+[\w\s\*]+?\b(bb_\w+)\s*\(', header))
    assert declared, 'no BB_API declarations found'
    assert declared == set(library.SIGNATURES)
    lib = library.load()                        # raises if any symbol is missing
    assert lib.bb_abi_version() == 2
    for name in declared:
        assert hasattr(lib, name)
