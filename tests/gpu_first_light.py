"""Diagnostic script for a GPU box (not a pytest file): runs each kernel family once against
the oracle and prints errors + timings.  `python tests/gpu_first_light.py`."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.stats as S  # noqa: E402
from oracle import closed_forms as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    rng = np.random.RandomState(0)
    for n, d in [(128, 64), (129, 64), (1, 64), (1000, 16), (10000, 16), (65536, 64), (100003, 32),
                 (5000, 60), (777, 20), (513, 96), (40, 3)]:
        X = (rng.randn(n, d) * 1.5 + 0.7).astype(np.float32)
        try:
            _, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
            torch.cuda.synchronize()
            _, r1, r2 = O.gaussian_suffstats(X)
            print('suffstats n=%d d=%d  relerr S1 %.2e  S2 %.2e' % (n, d, rel(s1.cpu(), r1), rel(s2.cpu(), r2)),
                  flush=True)
        except Exception as exc:
            print('suffstats n=%d d=%d FAILED: %s' % (n, d, exc), flush=True)
    # truncation semantics: values whose low mantissa bits would round up under RN-to-tf32
    X = (1.0 + 2.0 ** -11 + 2.0 ** -12 + rng.rand(4096, 64) * 2.0 ** -14).astype(np.float32)
    _, s1, s2 = S.gaussian_suffstats(torch.from_numpy(X).cuda())
    _, r1, r2 = O.gaussian_suffstats(X)
    print('truncation probe relerr S2 %.3e (3xTF32 ok if ~1e-6; RN-on-B would give ~2e-4)' % rel(s2.cpu(), r2))
    # host path
    X = rng.randn(300000, 64).astype(np.float32)
    _, h1, h2 = S.gaussian_suffstats(X, chunk_rows=65536)
    _, r1, r2 = O.gaussian_suffstats(X)
    print('host-streamed suffstats relerr S1 %.2e S2 %.2e' % (rel(h1, r1), rel(h2, r2)), flush=True)
    # timing at cfg2 size
    n, d = 1 << 24, 64
    Xd = torch.randn(n, d, device='cuda', dtype=torch.float32)
    out = (torch.empty(d, dtype=torch.float64, device='cuda'), torch.empty((d, d), dtype=torch.float64, device='cuda'))
    for _ in range(3):
        S.gaussian_suffstats(Xd, out=out)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    iters = 10
    for _ in range(iters):
        S.gaussian_suffstats(Xd, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    print('cfg2 suffstats: %.3f ms/pass  %.2f Gpts/s  %.1f GB/s' % (ms, n / ms / 1e6, n * d * 4 / ms / 1e6), flush=True)
    tr = float((Xd.double() ** 2).sum())
    print('trace check rel %.2e' % (abs(float(out[1].diagonal().sum()) - tr) / tr))
    s1ref = Xd.double().sum(0)
    print('S1 check rel %.2e' % float((out[0] - s1ref).abs().max() / s1ref.abs().max()))
    del Xd
    # logsoftmax
    for n, k in [(4096, 256), (1000, 128), (33, 7), (5, 1000), (64, 1024), (7, 33)]:
        Lg = (rng.randn(n, k) * 3).astype(np.float32)
        lr, lse, tot = S.log_responsibilities(torch.from_numpy(Lg).cuda())
        ref_lr, ref_lse = O.log_responsibilities(Lg)
        print('logsoftmax n=%d k=%d  max abs err %.2e  lse relerr %.2e  sum relerr %.2e' % (
            n, k, float(np.max(np.abs(lr.cpu().numpy() - ref_lr))), rel(lse.cpu(), ref_lse),
            abs(float(tot) - ref_lse.sum()) / abs(ref_lse.sum())), flush=True)
    n, k = 1 << 22, 256
    Lg = torch.randn(n, k, device='cuda') * 3
    outb = torch.empty_like(Lg)
    for _ in range(3):
        S.log_responsibilities(Lg, out=outb)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        S.log_responsibilities(Lg, out=outb)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print('logsoftmax 4Mi x 256: %.3f ms  %.1f GB/s' % (ms, n * (2 * k + 1) * 4 / ms / 1e6), flush=True)
    del Lg, outb
    # weighted stats
    for n, d, k in [(1000, 64, 8), (4099, 16, 5), (300, 6, 3)]:
        X = rng.randn(n, d).astype(np.float32)
        R = rng.dirichlet(np.ones(k), size=n).astype(np.float32)
        nk, rx, rxx = S.weighted_suffstats(torch.from_numpy(X).cuda(), torch.from_numpy(R).cuda())
        rnk, rrx, rrxx = O.weighted_suffstats(X, R)
        print('weighted n=%d d=%d k=%d  relerr Nk %.2e  rx %.2e  rxx %.2e' % (
            n, d, k, rel(nk.cpu(), rnk), rel(rx.cpu(), rrx), rel(rxx.cpu(), rrxx)), flush=True)
    print('launches so far', S.launch_count())


if __name__ == '__main__':
    main()
