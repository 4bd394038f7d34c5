"""Developer check + timing of the tcgen05 projection kernels (not a pytest file)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.stats as S  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    torch.manual_seed(0)
    ok = True
    # exact integer cases: any layout mistake is a whole-number error
    for n, d, q in [(128, 64, 64), (300, 128, 64), (1000, 512, 64), (257, 256, 128), (130, 64, 16), (5000, 128, 48)]:
        X = torch.randint(-4, 5, (n, d), device='cuda').float()
        W = torch.randint(-3, 4, (q, d), device='cuda').float()
        Z = S.row_projection(X, W)
        ref = X.double() @ W.double().T
        err = float((Z.double() - ref).abs().max())
        print('rowproj int  n=%d d=%d q=%d  max|err| %g' % (n, d, q, err), flush=True)
        ok &= err == 0.0
    for n, d, q in [(16, 128, 64), (100, 128, 64), (1000, 512, 64), (5000, 256, 128), (333, 128, 256), (4097, 384, 64)]:
        X = torch.randint(-4, 5, (n, d), device='cuda').float()
        R = torch.randint(-3, 4, (n, q), device='cuda').float()
        G = S.column_projection(X, R)
        ref = X.double().T @ R.double()
        err = float((G - ref).abs().max())
        print('colproj int  n=%d d=%d q=%d  max|err| %g' % (n, d, q, err), flush=True)
        ok &= err == 0.0
    # random data: error relative to the natural scale
    for n, d, q in [(20000, 512, 64), (70001, 128, 64)]:
        X = torch.randn(n, d, device='cuda') * 1.3 + 0.2
        W = torch.randn(q, d, device='cuda') / d ** 0.5
        Z = S.row_projection(X, W)
        ref = X.double() @ W.double().T
        scale = (X.double().norm(dim=1)[:, None] * W.double().norm(dim=1)[None, :])
        err = float(((Z.double() - ref).abs() / scale).max())
        print('rowproj rand n=%d d=%d q=%d  max err/(|x||w|) %.3g' % (n, d, q, err), flush=True)
        ok &= err < 2e-5
        R = torch.randn(n, q, device='cuda')
        G = S.column_projection(X, R)
        ref = X.double().T @ R.double()
        scale = (X.double().norm(dim=0)[:, None] * R.double().norm(dim=0)[None, :])
        err = float(((G - ref).abs() / scale).max())
        print('colproj rand n=%d d=%d q=%d  max err/(|x||r|) %.3g' % (n, d, q, err), flush=True)
        ok &= err < 3e-5
    # fused logistic pass vs float64 torch
    n, d, s = 50000, 512, 64
    X = torch.randn(n, d, device='cuda')
    W = torch.randn(s, d, device='cuda') / d ** 0.5
    y = (torch.rand(n, device='cuda') < torch.sigmoid(X @ W[0])).float()
    loglik, G = S.logistic_reparam_stats(X, y, W)
    Zd = X.double() @ W.double().T
    ll_ref = (y.double()[:, None] * Zd - torch.nn.functional.softplus(Zd)).sum(0)
    G_ref = X.double().T @ (y.double()[:, None] - torch.sigmoid(Zd))
    e1 = float(((loglik - ll_ref).abs() / ll_ref.abs()).max())
    e2 = float((G - G_ref).abs().max() / G_ref.abs().max())
    print('logistic pass n=%d: loglik max rel err %.3g, G max err / max|G| %.3g' % (n, e1, e2), flush=True)
    ok &= e1 < 1e-5 and e2 < 1e-4
    print('ALL OK' if ok else 'FAILURES', flush=True)
    if not ok:
        return 1
    # cfg5 timing
    n, d, s = 1 << 22, 512, 64
    X = torch.randn(n, d, device='cuda')
    W = torch.randn(s, d, device='cuda') / d ** 0.5
    y = (torch.rand(n, device='cuda') < 0.5).float()
    R = torch.randn(n, s, device='cuda')
    ms = timeit(lambda: S.row_projection(X, W))
    print('rowproj  cfg5: %.3f ms  %.0f GB/s (X read + Z write)' % (ms, (n * d * 4 + n * s * 4) / ms / 1e6), flush=True)
    ms = timeit(lambda: S.column_projection(X, R))
    print('colproj  cfg5: %.3f ms  %.0f GB/s (X + R read)' % (ms, (n * d * 4 + n * s * 4) / ms / 1e6), flush=True)
    ms = timeit(lambda: S.logistic_reparam_stats(X, y, W))
    print('logistic cfg5: %.3f ms  %.1f M rows/s  (algorithmic %d B/row -> %.0f GB/s)'
          % (ms, n / ms / 1e3, d * 4 + 4, n * (d * 4 + 4) / ms / 1e6), flush=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
