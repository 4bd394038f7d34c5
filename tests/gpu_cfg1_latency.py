"""Developer script: BASELINE cfg1 (the reference's own CPU-runnable case) -- latency of
trace(dot(L, dot(X.T, X))), N = 10 000, D = 16, through expr.compile() -> f(**inputs)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.algebra as A  # noqa: E402


def main():
    rng = np.random.RandomState(1234)
    n, d = 10000, 16
    Xh = rng.randn(n, d).astype(np.float32)
    a = rng.randn(d, d)
    Lh = (a @ a.T / d + np.eye(d)).astype(np.float32)
    X, L = A.var('X', 2), A.var('L', 2)
    fn = A.trace(A.dot(L, A.dot(X.T, X))).compile()
    want = float(np.einsum('de,nd,ne->', Lh.astype('f8'), Xh.astype('f8'), Xh.astype('f8')))
    got = float(fn(X=Xh, L=Lh))
    print('value %.6e  oracle %.6e  rel err %.2e  launches %d' % (got, want, abs(got - want) / abs(want),
                                                              fn.plan.last_launches))
    Xd, Ld = torch.from_numpy(Xh).cuda(), torch.from_numpy(Lh).cuda()
    Xp = torch.from_numpy(Xh).pin_memory()
    for label, kw in (('host numpy in/out', dict(X=Xh, L=Lh)), ('host, X page-locked', dict(X=Xp, L=Lh)),
                      ('device resident', dict(X=Xd, L=Ld))):
        for _ in range(20):
            out = fn(**kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            out = fn(**kw)
            if label.startswith('device'):
                pass
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print('%-20s %.1f us/pass  %.1f M pts/s' % (label, dt * 1e6, n / dt / 1e6))
    # CPU port of the reference plan for the same expression
    for _ in range(5):
        np.tensordot(Lh.T, np.tensordot(Xh.T, Xh, ([1], [0])), ([0, 1], [0, 1]))
    t0 = time.perf_counter()
    for _ in range(200):
        np.tensordot(Lh.T, np.tensordot(Xh.T, Xh, ([1], [0])), ([0, 1], [0, 1]))
    dt = (time.perf_counter() - t0) / 200
    print('%-20s %.1f us/pass  %.1f M pts/s (cores %d)' % ('numpy port (CPU)', dt * 1e6, n / dt / 1e6, os.cpu_count()))


if __name__ == '__main__':
    main()
