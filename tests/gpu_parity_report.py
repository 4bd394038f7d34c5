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


def _metrics(got, ref, cs):
    got = np.asarray(got.detach().cpu().numpy() if hasattr(got, 'detach') else got, dtype=np.float64).reshape(-1)
    scale = max(np.abs(ref).max(), 1e-300)
    elem = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * scale))
    norm = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
    cauchy = np.max(np.abs(got - ref) / cs) if cs is not None else float('nan')
    return elem, norm, cauchy


def row(cfg, name, got, ref, f32=None, cs=None):
    """One line per output: this repo's CUDA path AND (f32 column) a float32 numpy/BLAS evaluation of the
    same plan -- the reference's own arithmetic (Theano's CPU backend computes tensordot with float32
    BLAS, bayesic/algebra.py:1347-1351) -- both against the float64 oracle.  `elem` = max |err| /
    max(|ref|, 1e-3 max|ref|); `norm` = ||err|| / ||ref||; `cs` = max |err| / sqrt(S_aa S_bb), the
    Cauchy-Schwarz scale of the entry, where one applies."""
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    cs = None if cs is None else np.maximum(np.asarray(cs, dtype=np.float64).reshape(-1), 1e-300)
    e, nrm, c = _metrics(got, ref, cs)
    text = '%-6s %-30s n=%-8d ours: elem %.2e norm %.2e cs %.2e' % (cfg, name, ref.size, e, nrm, c)
    if f32 is not None:
        fe, fn, fc = _metrics(np.asarray(f32, dtype=np.float64), ref, cs)
        text += '   | float32 reference: elem %.2e norm %.2e cs %.2e | ours/f32 elem %.2f' % (fe, fn, fc, e / max(fe, 1e-300))
    print(text, flush=True)


def spd(rng, d):
    a = rng.randn(d, d)
    return a @ a.T / d + np.eye(d)


def main():
    print('# library: %s' % os.environ.get('BB_LIB_PATH', 'bayesic_b200/lib/libbayesic_b200.so (default build)'))
    rng = np.random.RandomState(2024)
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    # cfg1: the reference's own expression at its own size
    n, d = 10000, 16
    X = rng.randn(n, d).astype(np.float32)
    L = spd(rng, d).astype(np.float32)
    Xv, Lv = A.var('X', 2), A.var('L', 2)
    got = A.trace(A.dot(Lv, A.dot(Xv.T, Xv))).compile()(X=X, L=L)
    row('cfg1', 'trace(dot(L, dot(X.T, X)))', got, np.einsum('de,nd,ne->', L.astype('f8'), X.astype('f8'), X.astype('f8')),
        f32=np.trace(L @ (X.T @ X)))
    # cfg2
    n, d = 1 << 18, 64
    X = (rng.randn(n, d) * 1.3 + 0.4).astype(np.float32)
    ex = O.gaussian_wishart_expectations(rng.randn(d) * 0.1, 2.0, np.linalg.inv(spd(rng, d)) / (d + 4.0), d + 4.0)
    cnt, s1, s2, ell = P.gaussian_pass(torch.from_numpy(X).cuda(), *ex)
    rn, r1, r2 = O.gaussian_suffstats(X)
    f2, f1 = X.T @ X, X.sum(0)
    dg = np.sqrt(np.diag(r2))
    row('cfg2', 'sum x', s1, r1, f32=f1, cs=np.sqrt(n) * dg)
    row('cfg2', 'sum x x^T', s2, r2, f32=f2, cs=np.outer(dg, dg))
    row('cfg2', 'expected log-likelihood', ell, O.gaussian_expected_loglik(rn, r1, r2, *ex),
        f32=O.gaussian_expected_loglik(rn, f1.astype('f8'), f2.astype('f8'), *ex))
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
    # float32 reference arithmetic: the logit expression c + x.b - x^T A x / 2 and the statistics, all in float32
    XA = np.einsum('nd,kde->nke', X, Ak)
    lg32 = ck[None] + X @ bk.T - np.float32(0.5) * np.einsum('nke,ne->nk', XA, X)
    mx = lg32.max(1, keepdims=True)
    lse32 = (mx + np.log(np.exp(lg32 - mx).sum(1, keepdims=True))).astype(np.float32)
    lr32 = lg32 - lse32
    R32 = np.exp(lr32)
    Rw = want['log_resp']
    R64 = np.exp(Rw)
    phi = (X[:, :, None] * X[:, None, :]).reshape(n, d * d)
    rxx32 = (R32.T @ phi).reshape(k, d, d)
    rx32 = R32.T @ X
    x2 = (X.astype('f8') ** 2)
    cs_rx = np.sqrt(np.outer(R64.sum(0), np.ones(d)) * (R64.T @ x2))              # sqrt(sum r) sqrt(sum r x^2)
    cs_rxx = np.sqrt(np.einsum('kd,ke->kde', R64.T @ x2, R64.T @ x2))
    row('cfg3', 'log responsibilities', got['log_resp'], want['log_resp'], f32=lr32)
    row('cfg3', 'sum_n logsumexp', got['sum_lse'], want['sum_lse'], f32=lse32.astype('f8').sum())
    row('cfg3', 'N_k', got['nk'], want['nk'], f32=R32.sum(0))
    row('cfg3', 'sum r x', got['rx'], want['rx'], f32=rx32, cs=cs_rx)
    row('cfg3', 'sum r x x^T', got['rxx'], want['rxx'], f32=rxx32, cs=cs_rxx)
    got2 = step(dev(X), dev(Ak), dev(bk), dev(ck), want_log_resp=False)
    row('cfg3', 'sum r x x^T (local_step, R in place)', got2['rxx'], want['rxx'], f32=rxx32, cs=cs_rxx)
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
    dg = np.sqrt(np.diag(rxtx))
    row('cfg4', 'X^T X', xtx, rxtx, f32=X.T @ X, cs=np.outer(dg, dg))
    row('cfg4', 'X^T y', xty, rxty, f32=X.T @ y, cs=dg * np.sqrt(ryty))
    row('cfg4', 'y^T y', yty, ryty, f32=y @ y)
    # cfg5
    n, d, s = 1 << 16, 512, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    mu, ls, eps = rng.randn(d) * 0.05, np.full(d, -2.0), rng.randn(s, d)
    want = O.logistic_reparam_gradient(X, y, mu, ls, eps)
    got = P.LogisticReparamGrad()(dev(X), dev(y), t64(mu), t64(ls), t64(eps))
    # float32 reference arithmetic for the data-axis contractions (draws and gradient assembly in float64, as ours)
    Wd = (mu[None] + np.exp(ls)[None] * eps).astype(np.float32)
    Z32 = X @ Wd.T
    resid32 = (y[:, None] - 1.0 / (1.0 + np.exp(-Z32))).astype(np.float32)
    G32 = X.T @ resid32
    Z64 = X.astype('f8') @ Wd.astype('f8').T
    resid64 = y[:, None] - 1.0 / (1.0 + np.exp(-Z64))
    cs_G = np.linalg.norm(X.astype('f8'), axis=0)[:, None] * np.linalg.norm(resid64, axis=0)[None, :]
    gm32 = G32.astype('f8').mean(1) - mu
    row('cfg5', 'ELBO estimate', got['elbo'], want['elbo'])
    row('cfg5', 'G = X^T (y - sigmoid(Z))', got['G'], want['G'], f32=G32, cs=cs_G)
    row('cfg5', 'grad mu', got['grad_mu'], want['grad_mu'], f32=gm32, cs=cs_G.mean(1))
    row('cfg5', 'grad log sigma', got['grad_log_sigma'], want['grad_log_sigma'])


if __name__ == '__main__':
    main()
