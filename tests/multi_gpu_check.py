"""Developer / driver script (torchrun, one process per GPU, real NVLink peers): the SHARDED passes of the
BASELINE configurations against the float64 oracle on the concatenated rows, then their timing at the
configurations' full per-minibatch sizes with the library's own peer-memory all-reduce and with NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29531 tests/multi_gpu_check.py [--no-timing]

What is combined: the iid-summed statistics of bayesic/distribution/base.py:328-332.
Rank 0 prints one line per check ("ok"/"FAIL") and one JSON line of timings; exit code 1 on any failure."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bayesic_b200.parallel as PAR  # noqa: E402
import bayesic_b200.stats as S  # noqa: E402
from oracle import closed_forms as O  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
    failures = []

    def report(name, ok, detail=''):
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print('%-58s %s %s' % (name, 'ok' if float(flag) else 'FAIL', detail), flush=True)
        if not float(flag):
            failures.append(name)

    def scaled_err(got, want, scale):
        return float(np.max(np.abs(got.cpu().numpy() - want) / np.maximum(scale, 1e-300)))

    def replicated(t):
        """bit-identical on every rank"""
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t.contiguous())
        return all(torch.equal(g, gathered[0]) for g in gathered)

    peer = PAR.peer_comm_for(1 << 16, dev) is not None
    if rank == 0:
        print('# world %d, peer-memory all-reduce %s' % (world, 'available' if peer else 'UNAVAILABLE (NCCL)'), flush=True)

    # ---- parity of the sharded passes (every rank generates the full arrays from one seed, takes its shard) ----
    rng = np.random.RandomState(77)
    # cfg2: one-launch pass
    n, d = 40000 * world + 17, 64
    X = (rng.randn(n, d) * 1.3 + 0.4).astype(np.float32)
    a = rng.randn(d, d)
    e_lambda, e_lambda_mu = a @ a.T / d + np.eye(d), rng.randn(d)
    lo, hi = PAR.shard_bounds(n, world, rank)
    gp = PAR.GaussianPass(d, dev)
    el, elm = torch.from_numpy(e_lambda).to(dev), torch.from_numpy(e_lambda_mu).to(dev)
    for _ in range(3):
        cnt, s1, s2, ll = gp.run(torch.from_numpy(X[lo:hi]).to(dev), el, elm, 0.7, -1.3)
    gp.check()
    wn, w1, w2 = O.gaussian_suffstats(X)
    want_ll = O.gaussian_expected_loglik(wn, w1, w2, e_lambda, e_lambda_mu, 0.7, -1.3)
    dg = np.sqrt(np.diag(w2))
    e2 = scaled_err(s2, w2, np.outer(dg, dg))
    report('cfg2 one-launch pass: sum x x^T', e2 < 1e-5 and float(cnt) == n, 'cs err %.2e' % e2)
    report('cfg2 one-launch pass: expected log-likelihood', abs(float(ll) - want_ll) < 1e-5 * abs(want_ll),
           'rel %.2e' % (abs(float(ll) - want_ll) / abs(want_ll)))
    report('cfg2 one-launch pass: replicated bit for bit', replicated(s2) and replicated(ll))
    # cfg4
    n, d = 3000 * world + 5, 256
    X = rng.randn(n, d).astype(np.float32)
    y = (X @ (rng.randn(d) / np.sqrt(d)) + 0.1 * rng.randn(n)).astype(np.float32)
    lo, hi = PAR.shard_bounds(n, world, rank)
    got = PAR.regression_suffstats_sharded(torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(y[lo:hi]).to(dev))
    wxtx, wxty, wyty = O.regression_suffstats(X, y)
    dg = np.sqrt(np.diag(wxtx))
    e = scaled_err(got['xtx'], wxtx, np.outer(dg, dg))
    report('cfg4 sharded {X^T X, X^T y, y^T y}', e < 2e-5 and scaled_err(got['xty'], wxty, dg * np.sqrt(wyty)) < 2e-5
           and float(got['count']) == n, 'cs err %.2e' % e)
    report('cfg4 sharded: replicated bit for bit', replicated(got['xtx']))
    # cfg3
    n, d, k = 2500 * world + 3, 64, 256
    X = rng.randn(n, d).astype(np.float32)
    logits = (rng.randn(n, k) * 2.0).astype(np.float32)
    lo, hi = PAR.shard_bounds(n, world, rank)
    log_resp, lse, sum_lse = S.log_responsibilities(torch.from_numpy(logits[lo:hi]).to(dev))
    got = PAR.mixture_suffstats_sharded(torch.from_numpy(X[lo:hi]).to(dev), torch.exp(log_resp), sum_lse)
    R64 = np.exp(logits.astype('f8') - np.log(np.exp(logits.astype('f8')).sum(1, keepdims=True)))
    X64 = X.astype('f8')
    x2 = R64.T @ (X64 ** 2)
    wrxx = np.einsum('nk,nd,ne->kde', R64, X64, X64)
    e = scaled_err(got['rxx'], wrxx, np.sqrt(np.einsum('kd,ke->kde', x2, x2)))
    e_nk = float(np.max(np.abs(got['nk'].cpu().numpy() / R64.sum(0) - 1)))
    report('cfg3 sharded {N_k, sum r x, sum r x x^T, sum lse}', e < 2e-5 and e_nk < 2e-5 and float(got['count']) == n,
           'cs err %.2e, N_k rel %.2e' % (e, e_nk))
    report('cfg3 sharded: replicated bit for bit', replicated(got['rxx']))
    # cfg3, the whole local step at K = 256 (pre-split responsibilities, CTA-pair statistics kernel)
    import bayesic_b200.passes as P
    n, d, k = 1500 * world + 7, 64, 256
    centres = rng.randn(k, d) * 1.5
    X = (centres[rng.randint(k, size=n)] + rng.randn(n, d)).astype(np.float32)
    log_pi = np.log(rng.dirichlet(np.ones(k)))
    m, beta, nu = centres + 0.05 * rng.randn(k, d), rng.rand(k) * 5 + 1, d + 2 + rng.rand(k) * 5
    Wk = np.stack([np.linalg.inv((lambda a: a @ a.T / d + np.eye(d))(rng.randn(d, d))) / nu[j] for j in range(k)])
    want = O.gmm_vmp_step(X, log_pi, m, beta, Wk, nu)
    step = P.GmmStep()
    Ak, bk, ck = (torch.from_numpy(a).to(dev) for a in step.expectations(log_pi, m, beta, Wk, nu))
    Uw, tw, cw = step.whiten(Ak, bk, ck)
    lo, hi = PAR.shard_bounds(n, world, rank)
    got = PAR.mixture_local_step_sharded(torch.from_numpy(X[lo:hi]).to(dev), Uw, tw, cw)
    x2 = np.exp(want['log_resp']).T @ (X.astype('f8') ** 2)
    e = scaled_err(got['rxx'], want['rxx'], np.sqrt(np.einsum('kd,ke->kde', x2, x2)))
    e_lse = abs(float(got['sum_lse']) - want['sum_lse']) / abs(want['sum_lse'])
    report('cfg3 sharded local step (logits -> operand tiles -> statistics)', e < 2e-5 and e_lse < 1e-5 and
           float(got['count']) == n, 'cs err %.2e, sum lse rel %.2e' % (e, e_lse))
    report('cfg3 sharded local step: replicated bit for bit', replicated(got['rxx']))
    # cfg5
    n, d, s = 2000 * world + 9, 512, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    W = (rng.randn(s, d) / np.sqrt(d)).astype(np.float32)
    lo, hi = PAR.shard_bounds(n, world, rank)
    got = PAR.logistic_reparam_sharded(torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(y[lo:hi]).to(dev),
                                       torch.from_numpy(W).to(dev))
    Z = X.astype('f8') @ W.astype('f8').T
    resid = y[:, None] - 1.0 / (1.0 + np.exp(-Z))
    wG = X.astype('f8').T @ resid
    wll = (y[:, None] * Z - np.logaddexp(0.0, Z)).sum(0)
    e = scaled_err(got['G'], wG, np.linalg.norm(X.astype('f8'), axis=0)[:, None] * np.linalg.norm(resid, axis=0)[None, :])
    e_ll = float(np.max(np.abs(got['loglik'].cpu().numpy() / wll - 1)))
    report('cfg5 sharded {loglik[S], G[D, S]}', e < 2e-5 and e_ll < 2e-5 and float(got['count']) == n,
           'cs err %.2e, loglik rel %.2e' % (e, e_ll))
    report('cfg5 sharded: replicated bit for bit', replicated(got['G']))
    comm = PAR.peer_comm_for(1 << 16, dev)
    if comm is not None:
        comm.check()

    # ---- timing at the configurations' full minibatch sizes, sharded over the ranks (strong scaling) ----
    timings = {}
    if '--no-timing' not in sys.argv:
        def timed(fn, reps=10):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        gen = torch.Generator(device=dev).manual_seed(99 + rank)

        def both(name, fn, rows_total):
            rec = {}
            for mode in ('peer', 'nccl'):
                if mode == 'nccl':
                    os.environ['BB_P2P_ALLREDUCE'] = '0'
                else:
                    os.environ.pop('BB_P2P_ALLREDUCE', None)
                    if not peer:
                        continue
                ms = timed(fn)
                rec[mode + '_ms'] = ms
                rec[mode + '_rows_per_s'] = rows_total / (ms * 1e-3)
            os.environ.pop('BB_P2P_ALLREDUCE', None)
            timings[name] = rec

        # cfg4: minibatch 1 Mi x 1024
        n, d = 1 << 20, 1024
        lo, hi = PAR.shard_bounds(n, world, rank)
        X = torch.randn(hi - lo, d, device=dev, generator=gen)
        y = torch.randn(hi - lo, device=dev, generator=gen)
        both('cfg4_regression_1Mi_x_1024', lambda: PAR.regression_suffstats_sharded(X, y), n)
        timings['cfg4_regression_1Mi_x_1024']['local_only_ms'] = timed(lambda: S.regression_suffstats(X, y))
        del X, y
        # cfg5: minibatch 4 Mi x 512, S = 64
        n, d, s = 1 << 22, 512, 64
        lo, hi = PAR.shard_bounds(n, world, rank)
        X = torch.randn(hi - lo, d, device=dev, generator=gen)
        y = (torch.rand(hi - lo, device=dev, generator=gen) < 0.5).float()
        W = torch.randn(s, d, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) / d ** 0.5
        both('cfg5_logistic_4Mi_x_512_S64', lambda: PAR.logistic_reparam_sharded(X, y, W), n)
        timings['cfg5_logistic_4Mi_x_512_S64']['local_only_ms'] = timed(lambda: S.logistic_reparam_stats(X, y, W))
        del X, y, W
        # cfg3: 2 Mi rows per step in total (cfg3's 64 Mi rows stream through in such steps), K = 256, D = 64
        n, d, k = 1 << 21, 64, 256
        lo, hi = PAR.shard_bounds(n, world, rank)
        X = torch.randn(hi - lo, d, device=dev, generator=gen)
        R = torch.softmax(torch.randn(hi - lo, k, device=dev, generator=gen), dim=1)
        both('cfg3_weighted_stats_2Mi_x_64_K256', lambda: PAR.mixture_suffstats_sharded(X, R), n)
        timings['cfg3_weighted_stats_2Mi_x_64_K256']['local_only_ms'] = timed(lambda: S.weighted_suffstats(X, R))
        del R
        Uw = (torch.eye(d, device=dev) * 1.2).repeat(k, 1, 1).contiguous()
        tw = torch.randn(k, d, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
        cw = torch.randn(k, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        both('cfg3_local_step_2Mi_x_64_K256', lambda: PAR.mixture_local_step_sharded(X, Uw, tw, cw), n)
        timings['cfg3_local_step_2Mi_x_64_K256']['local_only_ms'] = timed(lambda: P.GmmStep.local_step(X, Uw, tw, cw))
        del X
        # the bare collectives on the three payloads
        for name, numel in (('allreduce_cfg3_payload', PAR.PackedStats.mixture(256, 64).numel),
                            ('allreduce_cfg4_payload', PAR.PackedStats.regression(1024).numel),
                            ('allreduce_cfg5_payload', PAR.PackedStats.logistic(512, 64).numel),
                            ('allreduce_cfg2_payload', PAR.PackedStats.gaussian(64).numel)):
            rec = {'float64': numel}
            buf = torch.zeros(numel, dtype=torch.float64, device=dev)
            rec['nccl_ms'] = timed(lambda: dist.all_reduce(buf), reps=20)
            c = PAR.peer_comm_for(numel, dev)
            if c is not None:
                rec['peer_ms'] = timed(lambda: c.allreduce(numel), reps=20)
                c.check()
            timings[name] = rec
        if rank == 0:
            print(json.dumps({'world': world, 'timings': timings}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('multi_gpu_check: %s' % ('all ok' if not failures else 'FAILED: ' + ', '.join(failures)), flush=True)
    return 1 if failures else 0


if __name__ == '__main__':
    sys.exit(main())
