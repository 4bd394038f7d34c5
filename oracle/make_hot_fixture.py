"""Regenerate ``tests/golden/hot_kernels_reference.npz``: outputs of the UNMODIFIED reference
(``bayesic.algebra`` through the numpy Theano shim, float32 like the reference computes) for the
hot-path expressions at sizes where the tcgen05 kernels -- not the generic ones -- serve them.
Inputs are not stored: ``tests/golden/hot_inputs.py`` regenerates them from a seed on both sides.
Authoring container only:  ``python -m oracle.make_hot_fixture``."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference_algebra  # noqa: E402
from tests.golden.hot_inputs import hot_inputs  # noqa: E402


def main():
    ref = load_reference_algebra()
    inp = hot_inputs()
    out = {}

    def run(expr, **names):
        return np.asarray(expr.compile()(**{k: inp[v] for k, v in names.items()}), dtype=np.float32)

    D, R, t, Lg = ref.var('D', 2), ref.var('R', 2), ref.var('t', 1), ref.var('Lg', 2)
    W, b = ref.var('W', 2), ref.var('b', 1)
    # cfg2: Gaussian statistics, n = 8192, d = 64
    out['cfg2_sxx'] = run(ref.dot(D.T, D), D='X2')
    out['cfg2_sx'] = run(ref.sum(D, axis=0), D='X2')
    # cfg3: responsibility-weighted statistics, n = 4096, d = 16, k = 8; log-softmax n = 512, k = 128
    out['cfg3_nk'] = run(ref.sum(R, axis=0), R='R3')
    out['cfg3_rx'] = run(ref.dot(R.T, D), R='R3', D='X3')
    out['cfg3_rxx'] = run(ref.einsum([(R, [('sum', 0), ('out', 0)]), (D, [('sum', 0), ('out', 1)]),
                                      (D, [('sum', 0), ('out', 2)])], 3), R='R3', D='X3')
    out['cfg3_logsoftmax_rows96'] = run(Lg - ref.log(ref.sum(ref.exp(Lg), axis=1)).dimshuffle(0, 'x'), Lg='Lg3')[:96]
    # cfg4: regression statistics, n = 2048, d = 256 (first 32 rows of X^T X are kept)
    out['cfg4_xtx_rows32'] = run(ref.dot(D.T, D), D='X4')[:32]
    out['cfg4_xty'] = run(ref.dot(D.T, t), D='X4', t='t4')
    out['cfg4_yty'] = run(ref.dot(t, t), t='t4')
    # cfg5: logistic pass, n = 2048, d = 128, s = 64
    Z = ref.dot(D, W.T)
    out['cfg5_loglik'] = run(ref.sum(b.dimshuffle(0, 'x') * Z - ref.log(1 + ref.exp(Z)), axis=0), D='X5', W='W5', b='b5')
    out['cfg5_grad'] = run(ref.dot(D.T, b.dimshuffle(0, 'x') - (1 + ref.exp(-1 * Z)) ** -1), D='X5', W='W5', b='b5')
    # cfg3 at K = 256: responsibilities by the reference's own spelling of the softmax (algebra.py:1435-1448), then the
    # three statistics plans over them; n = 2048, d = 16 (components 0-15 and 240-255 of sum r x x^T are kept)
    inp['R3b'] = run(ref.exp(Lg - ref.log(ref.sum(ref.exp(Lg), axis=1)).dimshuffle(0, 'x')), Lg='Lg3b')
    out['cfg3b_nk'] = run(ref.sum(R, axis=0), R='R3b')
    out['cfg3b_rx'] = run(ref.dot(R.T, D), R='R3b', D='X3b')
    rxx_b = run(ref.einsum([(R, [('sum', 0), ('out', 0)]), (D, [('sum', 0), ('out', 1)]),
                            (D, [('sum', 0), ('out', 2)])], 3), R='R3b', D='X3b')
    out['cfg3b_rxx_32'] = np.concatenate([rxx_b[:16], rxx_b[240:]], axis=0)
    # cfg4 at D = 136 (not a multiple of 256: the CTA-pair Gram kernel pads the feature axis on chip)
    out['cfg4b_xtx'] = run(ref.dot(D.T, D), D='X4b')
    path = os.path.join(ROOT, 'tests', 'golden', 'hot_kernels_reference.npz')
    np.savez_compressed(path, **out)
    print('wrote %s: %s' % (path, {k: v.shape for k, v in out.items()}))


if __name__ == '__main__':
    main()
