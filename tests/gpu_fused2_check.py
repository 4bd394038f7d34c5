"""Developer check for the cfg5 single-kernel pass (run on the GPU box): parity against float64 numpy
at a few shapes, then timing at BASELINE size.  (The first design, logistic_fused_sm100.cu, was removed in round 2.)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesic_b200 import stats  # noqa: E402


def check(n, d, s=64, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(s, d) / np.sqrt(d)).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    Z = X.astype(np.float64) @ W.astype(np.float64).T
    ll = (y[:, None] * Z - np.logaddexp(0, Z)).sum(0)
    R = y[:, None] - 1 / (1 + np.exp(-Z))
    G = X.astype(np.float64).T @ R
    got_ll, got_G = stats.logistic_reparam_stats(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(),
                                                 torch.from_numpy(W).cuda())
    torch.cuda.synchronize()
    scale = np.sqrt((X.astype(np.float64) ** 2).sum(0))[:, None] * np.sqrt((R ** 2).sum(0))[None, :]
    eg = np.abs(got_G.cpu().numpy() - G) / np.maximum(scale, 1e-30)
    el = np.abs(got_ll.cpu().numpy() - ll) / np.abs(ll).max()
    ok = eg.max() < 3e-5 and el.max() < 2e-5
    print("n=%d d=%d  G err/scale %.2e  loglik rel %.2e  %s" % (n, d, eg.max(), el.max(), 'ok' if ok else 'FAIL'),
          flush=True)
    return ok


def main():
    print("BB_FUSED_V2 =", os.environ.get('BB_FUSED_V2'), flush=True)
    ok = True
    shapes = [(64, 128), (4096, 128), (5000, 256), (10007, 384), (20000, 512), (64 * 148 * 3 + 17, 512)]
    for n, d in ([] if os.environ.get('FUSED2_TIMING_ONLY') else shapes):
        ok &= check(n, d)
    if not ok:
        sys.exit(1)
    n, d, s = 4 * 1024 * 1024, 512, 64
    X = torch.randn(n, d, device='cuda')
    y = (torch.rand(n, device='cuda') < 0.5).float()
    W = torch.randn(s, d, device='cuda') / d ** 0.5
    for _ in range(3):
        stats.logistic_reparam_stats(X, y, W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        stats.logistic_reparam_stats(X, y, W)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("cfg5 4Mi x 512: %.3f ms/pass  %.2f G rows/s  %.0f GB/s algorithmic" % (ms, n / ms / 1e6, n * (4 * d + 4) / ms / 1e6),
          flush=True)


if __name__ == '__main__':
    main()
