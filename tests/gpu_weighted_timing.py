"""Developer timing script for the weighted-statistics kernels (not a pytest file)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.stats as S  # noqa: E402


def main():
    for n, d, k in [(1 << 20, 64, 256), (1 << 20, 64, 16), (1 << 22, 64, 64)]:
        X = torch.randn(n, d, device='cuda')
        R = torch.softmax(torch.randn(n, k, device='cuda') * 2, 1)
        nk, rx, rxx = S.weighted_suffstats(X, R)
        torch.cuda.synchronize()
        # quick check of two components against float64 torch
        for c in (0, k - 1):
            ref = (X.double() * R[:, c:c + 1].double()).T @ X.double()
            err = float((rxx[c] - ref).abs().max() / ref.abs().max())
            assert err < 5e-5, err
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            S.weighted_suffstats(X, R)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('weighted n=%d d=%d k=%d: %.2f ms  %.1f Mpts/s  %.1f TFLOP/s useful (2 K D^2 per row)  relerr %.1e'
              % (n, d, k, ms, n / ms / 1e3, 2 * k * d * d * n / ms / 1e9, err), flush=True)
        del X, R


if __name__ == '__main__':
    main()
