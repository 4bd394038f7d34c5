"""Developer timing of the BASELINE cfg3/cfg4/cfg5 passes (not a pytest file)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bayesic_b200.passes as P  # noqa: E402
import bayesic_b200.stats as S  # noqa: E402


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cfg4():
    n, d = 1 << 20, 1024
    X = torch.randn(n, d, device='cuda')
    y = X @ (torch.randn(d, device='cuda') / d ** 0.5) + 0.1 * torch.randn(n, device='cuda')
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device='cuda')
    eta1, eta2 = t(np.zeros(d)), t(-0.5 * np.eye(d))
    for fused in (True, False):
        step = P.LinRegSviStep(fused=fused)
        ms = timeit(lambda: step(X, y, eta1, eta2, 100.0, 10 * n, 0.3, eta1, eta2))
        print('cfg4 LinRegSviStep fused=%s: %.2f ms/step  %.1f M rows/s' % (fused, ms, n / ms / 1e3), flush=True)
    ms = timeit(lambda: S.regression_suffstats(X, y), reps=10)
    print('cfg4 regression_suffstats alone: %.3f ms  %.1f M rows/s' % (ms, n / ms / 1e3), flush=True)


def cfg5():
    n, d, s = 1 << 22, 512, 64
    X = torch.randn(n, d, device='cuda')
    y = (torch.rand(n, device='cuda') < 0.5).float()
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device='cuda')
    mu, ls, eps = t(np.zeros(d)), t(np.full(d, -2.0)), t(np.random.RandomState(0).randn(s, d))
    step = P.LogisticReparamGrad()
    ms = timeit(lambda: step(X, y, mu, ls, eps), reps=2)
    print('cfg5 LogisticReparamGrad: %.2f ms/step  %.1f M rows/s (launches %d)'
          % (ms, n / ms / 1e3, step.fn.plan.last_launches), flush=True)


def cfg3():
    n, d, k = 1 << 21, 64, 256
    X = torch.randn(n, d, device='cuda')
    Ak = torch.eye(d, device='cuda').repeat(k, 1, 1).contiguous()
    bk = torch.randn(k, d, device='cuda')
    ck = torch.randn(k, device='cuda')
    step = P.GmmStep()
    ms = timeit(lambda: step(X, Ak, bk, ck), reps=2)
    print('cfg3 GmmStep (N = %d): %.2f ms/step  %.2f M rows/s' % (n, ms, n / ms / 1e3), flush=True)
    ms = timeit(lambda: step(X, Ak, bk, ck, want_log_resp=False), reps=2)
    print('cfg3 GmmStep without materialising R (N = %d): %.2f ms/step  %.2f M rows/s' % (n, ms, n / ms / 1e3), flush=True)
    U, t, c = step.whiten(Ak, bk, ck)
    for materialise in (None, True, False):
        ms = timeit(lambda: step.local_step(X, U, t, c, materialise=materialise), reps=3)
        print('cfg3 local_step(materialise=%s) (N = %d): %.2f ms/step  %.2f M rows/s' % (materialise, n, ms, n / ms / 1e3), flush=True)
    ms = timeit(lambda: S.mixture_logits(X, U, t, c, upper_triangular=True), reps=3)
    print('  mixture logits kernel: %.2f ms  %.1f M rows/s  %.0f TFLOP/s issued bf16'
          % (ms, n / ms / 1e3, 3 * 2.0 * k * d * d * n / ms / 1e9), flush=True)
    R = torch.softmax(torch.randn(n, k, device='cuda'), 1)
    ms = timeit(lambda: S.weighted_suffstats(X, R), reps=5)
    print('  weighted stats: %.2f ms' % ms, flush=True)
    ms = timeit(lambda: S.log_responsibilities(R), reps=5)
    print('  log-softmax: %.2f ms' % ms, flush=True)
    ms = timeit(lambda: S.responsibilities(R), reps=5)
    print('  softmax (responsibilities): %.2f ms' % ms, flush=True)
    lg = torch.randn(n, k, device='cuda')
    lse = torch.logsumexp(lg, 1)
    ms = timeit(lambda: S.responsibilities_split(lg), reps=5)
    print('  softmax -> pre-split operand tiles: %.2f ms' % ms, flush=True)
    rsplit, _, _ = S.responsibilities_split(lg)
    ms = timeit(lambda: S.weighted_suffstats_split(X, rsplit, k), reps=5)
    print('  weighted stats from pre-split tiles: %.2f ms' % ms, flush=True)
    ms = timeit(lambda: S.weighted_suffstats_from_logits(X, lg, lse), reps=2)
    print('  weighted stats from logits: %.2f ms' % ms, flush=True)


def cfg3_full():
    """BASELINE cfg3 at its full extent: N = 64 Mi rows, K = 256, D = 64 (16 GiB of data, 64 GiB of logits)."""
    n, d, k = 1 << 26, 64, 256
    X = torch.randn(n, d, device='cuda')
    Ak = torch.eye(d, device='cuda').repeat(k, 1, 1).contiguous()
    bk = torch.randn(k, d, device='cuda')
    ck = torch.randn(k, device='cuda')
    step = P.GmmStep()
    U, t, c = step.whiten(Ak, bk, ck)

    def local_step():
        logits, lse, _ = S.mixture_logits(X, U, t, c, upper_triangular=True)
        return S.weighted_suffstats_from_logits(X, logits, lse)
    ms = timeit(local_step, reps=2)
    print('cfg3 full size, local step without materialising R (N = %d): %.1f ms  %.2f M rows/s'
          % (n, ms, n / ms / 1e3), flush=True)


def cfg3_full_slabs():
    """BASELINE cfg3 at its full extent (N = 64 Mi rows resident, K = 256, D = 64) through the default local step,
    in slabs of 8 Mi rows (the logits and operand-tile buffers are 8 GiB each and are reused); the statistics of the
    slabs add up on the device."""
    n, d, k, slab = 1 << 26, 64, 256, 1 << 23
    X = torch.randn(n, d, device='cuda')
    step = P.GmmStep()
    U = (torch.eye(d, device='cuda') * 1.2).repeat(k, 1, 1).contiguous()
    t = torch.randn(k, d, device='cuda')
    c = torch.randn(k, device='cuda')

    def full_pass():
        tot = None
        for lo in range(0, n, slab):
            out = step.local_step(X[lo:lo + slab], U, t, c)
            part = (out['nk'], out['rx'], out['rxx'], out['sum_lse'])
            del out
            tot = part if tot is None else tuple(a + b for a, b in zip(tot, part))
        return tot
    full_pass()
    torch.cuda.synchronize()
    ms = timeit(full_pass, reps=2, warm=0)
    nk = full_pass()[0]
    print('cfg3 full size, default local step in 8 Mi-row slabs (N = %d): %.1f ms  %.2f M rows/s  (sum N_k / N = %.6f)'
          % (n, ms, n / ms / 1e3, float(nk.sum()) / n), flush=True)


def loops():
    """Whole iterations built from the update kernels (bayesic_b200/updates.py)."""
    import bayesic_b200.updates as U
    n, d, k = 1 << 21, 64, 256
    g = torch.Generator(device='cuda').manual_seed(0)
    centres = torch.randn(k, d, device='cuda', generator=g) * 3
    X = centres[torch.randint(0, k, (n,), device='cuda', generator=g)] + torch.randn(n, d, device='cuda', generator=g)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device='cuda')
    vmp = U.GmmVmp(k, d, 1.0, 1.0, d + 2.0, t(np.zeros(d)), t(np.eye(d)))
    R0 = torch.softmax(torch.randn(1 << 16, k, device='cuda', generator=g), 1)
    vmp.initialise(*S.weighted_suffstats(X[:1 << 16], R0))
    ms = timeit(lambda: vmp.step(X), reps=3)
    print('cfg3 GmmVmp.step, local + all-reduce layout + global update (N = %d): %.2f ms/iteration  %.2f M rows/s'
          % (n, ms, n / ms / 1e3), flush=True)
    nk, rx, rxx = S.weighted_suffstats(X[:1 << 16], R0)
    ms = timeit(lambda: U.gmm_global_update(nk, rx, rxx, 1.0, 1.0, d + 2.0, t(np.zeros(d)), t(np.eye(d))), reps=10)
    print('  gmm_global_update kernel (K = %d, D = %d): %.3f ms' % (k, d, ms), flush=True)
    n, d, s = 1 << 22, 512, 64
    X = torch.randn(n, d, device='cuda')
    y = (torch.rand(n, device='cuda') < 0.5).float()
    loop = U.LogisticReparamSgd(t(np.zeros(d)), t(np.full(d, -2.0)), t(np.random.RandomState(0).randn(s, d)))
    ms = timeit(lambda: loop.step(X, y), reps=3)
    print('cfg5 LogisticReparamSgd.step, draws + pass + gradient + Adam: %.2f ms/iteration  %.1f M rows/s'
          % (ms, n / ms / 1e3), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['cfg4', 'cfg5', 'cfg3', 'loops']      # also: cfg3_full, cfg3_full_slabs
    for name in which:
        globals()[name]()
