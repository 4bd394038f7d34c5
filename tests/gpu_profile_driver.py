"""Developer driver for ncu: runs ONE of the hot kernels a few times at (scaled) BASELINE extents.
    python tests/gpu_profile_driver.py gram|rowproj|colproj|logistic|weighted|weighted_split|softmax_split|logits|suffstats|logsoftmax"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.stats as S  # noqa: E402


def main(which):
    torch.manual_seed(0)
    if which == 'gram':                      # cfg4: minibatch 1 Mi, D = 1024
        n, d = 1 << 20, 1024
        X = torch.randn(n, d, device='cuda')
        y = torch.randn(n, device='cuda')
        fn = lambda: S.regression_suffstats(X, y)
    elif which in ('rowproj', 'colproj', 'logistic'):   # cfg5: minibatch 4 Mi, D = 512, S = 64
        n, d, s = 1 << 22, 512, 64
        X = torch.randn(n, d, device='cuda')
        W = torch.randn(s, d, device='cuda') / d ** 0.5
        y = (torch.rand(n, device='cuda') < 0.5).float()
        R = torch.randn(n, s, device='cuda')
        fn = {'rowproj': lambda: S.row_projection(X, W), 'colproj': lambda: S.column_projection(X, R),
              'logistic': lambda: S.logistic_reparam_stats(X, y, W)}[which]
    elif which in ('weighted', 'logits'):    # cfg3 at 1 Mi rows: K = 256, D = 64
        n, d, k = 1 << 20, 64, 256
        X = torch.randn(n, d, device='cuda')
        if which == 'weighted':
            R = torch.softmax(torch.randn(n, k, device='cuda') * 2, 1)
            fn = lambda: S.weighted_suffstats(X, R)
        else:
            U = torch.eye(d, device='cuda').repeat(k, 1, 1).contiguous() + 0.01 * torch.randn(k, d, d, device='cuda').triu()
            t = torch.randn(k, d, device='cuda')
            c = torch.randn(k, device='cuda')
            fn = lambda: S.mixture_logits(X, U, t, c, upper_triangular=True)
    elif which == 'weighted_split':          # cfg3 at 1 Mi rows, responsibilities pre-split into operand tiles (CTA pairs)
        n, d, k = 1 << 20, 64, 256
        X = torch.randn(n, d, device='cuda')
        rsplit, _, _ = S.responsibilities_split(torch.randn(n, k, device='cuda') * 2)
        fn = lambda: S.weighted_suffstats_split(X, rsplit, k)
    elif which == 'softmax_split':
        Lg = torch.randn(1 << 21, 256, device='cuda') * 3
        fn = lambda: S.responsibilities_split(Lg)
    elif which == 'suffstats':               # cfg2: N = 16 Mi, D = 64
        X = torch.randn(1 << 24, 64, device='cuda')
        fn = lambda: S.gaussian_suffstats(X)
    elif which == 'logsoftmax':              # cfg3b at 4 Mi rows
        Lg = torch.randn(1 << 22, 256, device='cuda') * 3
        fn = lambda: S.log_responsibilities(Lg)
    else:
        raise SystemExit('unknown kernel %r' % which)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print('%s: %.3f ms per call' % (which, e0.elapsed_time(e1) / 5))


if __name__ == '__main__':
    main(sys.argv[1])
