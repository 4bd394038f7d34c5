"""The hot kernels against outputs of the UNMODIFIED reference on the same seeded inputs
(``tests/golden/hot_kernels_reference.npz``, made by ``oracle/make_hot_fixture.py`` in the authoring
container) at sizes where the tcgen05 kernels -- not the generic executor kernels the small golden
cases exercise -- are what runs.  The reference computes in float32 (numpy/BLAS behind the Theano
shim), so the comparison is rtol 1e-4 plus an absolute term relative to the natural scale of each
output (sqrt of the two squared norms being contracted); the CPU test applies the same comparison
to the float64 oracle, which pins the oracle to the reference at these sizes as well."""
import os

import numpy as np
import pytest

from oracle import closed_forms as O
from tests.golden.hot_inputs import hot_inputs

FIXTURE = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hot_kernels_reference.npz'))
INPUTS = hot_inputs()


def _norms(a, axis):
    return np.linalg.norm(np.asarray(a, dtype=np.float64), axis=axis)


def _close(name, got, scale, rtol=1e-4, tol=1e-4):
    want = np.asarray(FIXTURE[name], dtype=np.float64)
    got = np.asarray(got, dtype=np.float64).reshape(want.shape)
    assert np.all(np.abs(got - want) <= rtol * np.abs(want) + tol * np.broadcast_to(scale, want.shape)), \
        (name, float(np.abs(got - want).max()))


def _check_all(results):
    """``results``: dict of every fixture entry computed by the implementation under test."""
    X2, X3, R3, X4, t4, X5 = (INPUTS[k] for k in ('X2', 'X3', 'R3', 'X4', 't4', 'X5'))
    c2 = _norms(X2, 0)
    _close('cfg2_sxx', results['cfg2_sxx'], np.outer(c2, c2))
    _close('cfg2_sx', results['cfg2_sx'], c2 * np.sqrt(X2.shape[0]))
    r3, c3 = _norms(R3, 0), _norms(X3, 0)
    _close('cfg3_nk', results['cfg3_nk'], r3 * np.sqrt(R3.shape[0]))
    _close('cfg3_rx', results['cfg3_rx'], np.outer(r3, c3))
    rx_scale = np.sqrt(np.einsum('nk,nd,ne->kde', R3.astype('f8'), X3.astype('f8') ** 2, np.ones_like(X3, dtype='f8'))
                       * np.einsum('nk,nd,ne->kde', R3.astype('f8'), np.ones_like(X3, dtype='f8'), X3.astype('f8') ** 2))
    _close('cfg3_rxx', results['cfg3_rxx'], rx_scale)
    _close('cfg3_logsoftmax_rows96', results['cfg3_logsoftmax_rows96'], 1.0, tol=1e-4)
    c4 = _norms(X4, 0)
    _close('cfg4_xtx_rows32', results['cfg4_xtx_rows32'], np.outer(c4[:32], c4))
    _close('cfg4_xty', results['cfg4_xty'], c4 * np.linalg.norm(t4.astype('f8')))
    _close('cfg4_yty', results['cfg4_yty'], float(t4.astype('f8') @ t4.astype('f8')))
    _close('cfg5_loglik', results['cfg5_loglik'], float(X5.shape[0]))
    _close('cfg5_grad', results['cfg5_grad'], np.outer(_norms(X5, 0), np.full(64, np.sqrt(X5.shape[0]))))
    # cfg3 at K = 256: responsibilities by the reference's own softmax spelling, then its statistics plans
    X3b = INPUTS['X3b'].astype('f8')
    Lg = INPUTS['Lg3b'].astype('f8')
    R3b = np.exp(Lg - np.log(np.exp(Lg).sum(1, keepdims=True)))
    r3b, x2b = np.sqrt((R3b ** 2).sum(0)), R3b.T @ (X3b ** 2)
    keep = np.r_[0:16, 240:256]
    _close('cfg3b_nk', results['cfg3b_nk'], r3b * np.sqrt(R3b.shape[0]))
    _close('cfg3b_rx', results['cfg3b_rx'], np.outer(r3b, _norms(X3b, 0)))
    _close('cfg3b_rxx_32', results['cfg3b_rxx_32'], np.sqrt(np.einsum('kd,ke->kde', x2b, x2b))[keep])
    c4b = _norms(INPUTS['X4b'], 0)
    _close('cfg4b_xtx', results['cfg4b_xtx'], np.outer(c4b, c4b))


def test_oracle_matches_the_reference_outputs_at_kernel_sizes():
    X2, X3, R3, Lg3, X4, t4, X5, W5, b5 = (INPUTS[k].astype(np.float64) for k in
                                           ('X2', 'X3', 'R3', 'Lg3', 'X4', 't4', 'X5', 'W5', 'b5'))
    _, s1, s2 = O.gaussian_suffstats(X2)
    nk, rx, rxx = O.weighted_suffstats(X3, R3)
    log_r, _ = O.log_responsibilities(Lg3)
    xtx, xty, yty = O.regression_suffstats(X4, t4)
    Z = X5 @ W5.T
    loglik = (b5[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)
    grad = X5.T @ (b5[:, None] - 1.0 / (1.0 + np.exp(-Z)))
    log_rb, _ = O.log_responsibilities(INPUTS['Lg3b'].astype(np.float64))
    nkb, rxb, rxxb = O.weighted_suffstats(INPUTS['X3b'].astype(np.float64), np.exp(log_rb))
    _check_all({'cfg2_sxx': s2, 'cfg2_sx': s1, 'cfg3_nk': nk, 'cfg3_rx': rx, 'cfg3_rxx': rxx,
                'cfg3_logsoftmax_rows96': log_r[:96], 'cfg4_xtx_rows32': xtx[:32], 'cfg4_xty': xty, 'cfg4_yty': yty,
                'cfg5_loglik': loglik, 'cfg5_grad': grad,
                'cfg3b_nk': nkb, 'cfg3b_rx': rxb, 'cfg3b_rxx_32': np.concatenate([rxxb[:16], rxxb[240:]], axis=0),
                'cfg4b_xtx': O.regression_suffstats(INPUTS['X4b'].astype(np.float64), np.zeros(1536))[0]})


@pytest.mark.gpu
def test_hot_kernels_match_the_reference_outputs():
    import torch
    from bayesic_b200 import stats as S
    dev = {k: torch.from_numpy(v).cuda() for k, v in INPUTS.items()}
    host = lambda t: t.detach().cpu().numpy()
    _, s1, s2 = S.gaussian_suffstats(dev['X2'])                               # tcgen05 TF32 kernel (cfg2)
    nk, rx, rxx = S.weighted_suffstats(dev['X3'], dev['R3'])                  # BF16 pair-block kernel (cfg3)
    log_r, _, _ = S.log_responsibilities(dev['Lg3'])                          # vectorised log-softmax (cfg3b)
    xtx, xty, yty = S.regression_suffstats(dev['X4'], dev['t4'])              # CTA-pair Gram kernel (cfg4)
    loglik, grad = S.logistic_reparam_stats(dev['X5'], dev['b5'], dev['W5'])  # single-kernel logistic pass (cfg5)
    # cfg3 at K = 256: softmax written as BF16 operand tiles, statistics on CTA pairs (the default local-step route)
    rsplit, _, _ = S.responsibilities_split(dev['Lg3b'])
    nkb, rxb, rxxb = S.weighted_suffstats_split(dev['X3b'], rsplit, 256)
    _check_all({'cfg2_sxx': host(s2), 'cfg2_sx': host(s1), 'cfg3_nk': host(nk), 'cfg3_rx': host(rx),
                'cfg3_rxx': host(rxx), 'cfg3_logsoftmax_rows96': host(log_r)[:96],
                'cfg4_xtx_rows32': host(xtx)[:32], 'cfg4_xty': host(xty), 'cfg4_yty': host(yty).reshape(()),
                'cfg5_loglik': host(loglik), 'cfg5_grad': host(grad),
                'cfg3b_nk': host(nkb), 'cfg3b_rx': host(rxb),
                'cfg3b_rxx_32': np.concatenate([host(rxxb)[:16], host(rxxb)[240:]], axis=0),
                'cfg4b_xtx': host(S.regression_suffstats(dev['X4b'])[0])})
