"""Developer script: one table of parity figures per BASELINE configuration -- every output of the
CUDA path against the float64 oracle on the same seeded inputs: the largest elementwise relative
error (with the absolute floor the tests use) and the norm-wise error ||got - ref|| / ||ref||
(SURVEY.md 8(d)).  Run on the GPU box; the output is kept under profiles/."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.algebra as A  # noqa: E402
import bayesic_b200.passes as P  # noqa: E402
import bayesic_b200.stats as S  # noqa: E402
import bayesic_b200.updates as U  # noqa: E402
from oracle import closed_forms as O  # noqa: E402


def row(cfg, name, got, ref):
    got = np.asarray(got.detach().cpu().numpy() if hasattr(got, 'detach') else got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    scale = max(np.abs(ref).max(), 1e-300)
    elem = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * scale))
    norm = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
    print('%-6s %-34s n=%-9d max elementwise rel %.2e   ||err||/||ref|| %.2e' % (cfg, name, ref.size, elem, norm), flush=True)


def spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def main():
    rng = np.random.RandomState(2024)
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    # cfg1: the reference's own expression at its own size
    n, d = 10000, 16
    X = rng.randn(n, d).astype(np.float32)
    L = spd(rng, d).astype(np.float32)
    Xv, Lv = A.var('X', 2), A.var('L', 2)
    got = A.trace(A.dot(Lv, A.dot(Xv.T, Xv))).compile()(X=X, L=L)
    row('cfg1', 'trace(dot(L, dot(X.T, X)))', got, np.einsum('de,nd,ne->', L.astype('f8'), X.astype('f8'), X.astype('f8')))
    # cfg2
    n, d = 1 << 18, 64
    X = (rng.randn(n, d) * 1.3 + 0.4).astype(np.float32)
    ex = O.gaussian_wishart_expectations(rng.randn(d) * 0.1, 2.0, np.linalg.inv(spd(rng, d)) / (d + 4.0), d + 4.0)
    cnt, s1, s2, ell = P.gaussian_pass(torch.from_numpy(X).cuda(), *ex)
    rn, r1, r2 = O.gaussian_suffstats(X)
    row('cfg2', 'sum x', s1, r1)
    row('cfg2', 'sum x x^T', s2, r2)
    row('cfg2', 'expected log-likelihood', ell, O.gaussian_expected_loglik(rn, r1, r2, *ex))
    # cfg3
    n, d, k = 1 << 15, 64, 256
    centres = rng.randn(k, d) * 2
    X = (centres[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    m, beta, nu = centres + 0.1 * rng.randn(k, d), np.full(k, 2.0), np.full(k, d + 3.0)
    W = np.stack([np.linalg.inv(spd(rng, d)) / (d + 3.0) for _ in range(k)])
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    want = O.gmm_vmp_step(X, log_pi, m, beta, W, nu)
    step = P.GmmStep()
    Ak, bk, ck = step.expectations(log_pi, m, beta, W, nu)
    dev = lambda a: torch.from_numpy(a).cuda()
    got = step(dev(X), dev(Ak), dev(bk), dev(ck))
    row('cfg3', 'log responsibilities', got['log_resp'], want['log_resp'])
    row('cfg3', 'sum_n logsumexp', got['sum_lse'], want['sum_lse'])
    row('cfg3', 'N_k', got['nk'], want['nk'])
    row('cfg3', 'sum r x', got['rx'], want['rx'])
    row('cfg3', 'sum r x x^T', got['rxx'], want['rxx'])
    got2 = step(dev(X), dev(Ak), dev(bk), dev(ck), want_log_resp=False)
    row('cfg3', 'sum r x x^T (R never written)', got2['rxx'], want['rxx'])
    upd = U.gmm_global_update(t64(want['nk']), t64(want['rx']), t64(want['rxx']), 1.0, 1.0, d + 2.0,
                              t64(np.zeros(d)), t64(np.eye(d)))
    ref = O.gmm_global_update(want['nk'], want['rx'], want['rxx'], 1.0, 1.0, d + 2.0, np.zeros(d), np.eye(d))
    row('cfg3', 'global update: m_k', upd['m'], ref['m'])
    row('cfg3', 'global update: W_k^-1', upd['W_inv'], ref['W_inv'])
    row('cfg3', 'global update: KL terms', upd['kl'], ref['kl'])
    # cfg4
    n, d = 1 << 15, 1024
    X = rng.randn(n, d).astype(np.float32)
    y = (X @ (rng.randn(d) / np.sqrt(d)) + 0.1 * rng.randn(n)).astype(np.float32)
    xtx, xty, yty = S.regression_suffstats(dev(X), dev(y))
    rxtx, rxty, ryty = O.regression_suffstats(X, y)
    row('cfg4', 'X^T X', xtx, rxtx)
    row('cfg4', 'X^T y', xty, rxty)
    row('cfg4', 'y^T y', yty, ryty)
    # cfg5
    n, d, s = 1 << 16, 512, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    mu, ls, eps = rng.randn(d) * 0.05, np.full(d, -2.0), rng.randn(s, d)
    want = O.logistic_reparam_gradient(X, y, mu, ls, eps)
    got = P.LogisticReparamGrad()(dev(X), dev(y), t64(mu), t64(ls), t64(eps))
    row('cfg5', 'ELBO estimate', got['elbo'], want['elbo'])
    row('cfg5', 'G = X^T (y - sigmoid(Z))', got['G'], want['G'])
    row('cfg5', 'grad mu', got['grad_mu'], want['grad_mu'])
    row('cfg5', 'grad log sigma', got['grad_log_sigma'], want['grad_log_sigma'])


if __name__ == '__main__':
    main()
