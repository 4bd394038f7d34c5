"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, packed layout and the
single all-reduce.  Local statistics come from the oracle here (no GPU in this container); the
CUDA kernel itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest

from bayesic_b200.parallel import shard_bounds, PackedStats, allreduce_packed


def test_shard_bounds_partition_the_axis():
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_packed_layout_views_alias_the_buffer():
    layout = PackedStats.gaussian(3)
    assert layout.numel == 9 + 3 + 1
    buf = np.zeros(layout.numel)
    v = layout.views(buf)
    v['s2'][1, 2] = 5.0
    v['count'][0] = 7.0
    assert buf[5] == 5.0 and buf[-1] == 7.0
    mix = PackedStats.mixture(4, 3)
    assert mix.numel == 4 * 9 + 12 + 4 + 2
    with pytest.raises(ValueError):
        layout.views(np.zeros(5))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, queue):
    import torch
    import torch.distributed as dist
    from oracle.closed_forms import gaussian_suffstats
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(42)                 # every rank regenerates the same data
        X = (rng.randn(n, d) + 0.3).astype(np.float32)
        lo, hi = shard_bounds(n, world, rank)
        cnt, s1, s2 = gaussian_suffstats(X[lo:hi])      # local partial statistics (oracle)
        layout = PackedStats.gaussian(d)
        buf = layout.allocate()
        views = layout.views(buf)
        views['s2'].copy_(torch.from_numpy(s2))
        views['s1'].copy_(torch.from_numpy(s1))
        views['count'].fill_(float(cnt))
        allreduce_packed(buf)                           # ONE collective
        queue.put((rank, buf.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_the_full_statistics():
    import torch.multiprocessing as mp
    from oracle.closed_forms import gaussian_suffstats
    world, n, d = 2, 1001, 6
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(42)
    X = (rng.randn(n, d) + 0.3).astype(np.float32)
    cnt, s1, s2 = gaussian_suffstats(X)
    layout = PackedStats.gaussian(d)
    for rank in range(world):
        v = layout.views(results[rank])
        np.testing.assert_allclose(v['s2'], s2, rtol=1e-12)
        np.testing.assert_allclose(v['s1'], s1, rtol=1e-12)
        assert v['count'][0] == cnt
    np.testing.assert_array_equal(results[0], results[1])    # replicas agree bit for bit


def _worker_all_layouts(rank, world, port, queue):
    """Every packed layout through the same one-collective path, local parts from the oracle."""
    import torch
    import torch.distributed as dist
    from oracle import closed_forms as O
    from bayesic_b200.parallel import _reduce_into
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(7)
        n, d, k, s = 501, 4, 3, 2
        X = rng.randn(n, d).astype(np.float32)
        y = rng.randn(n).astype(np.float32)
        R = rng.dirichlet(np.ones(k), size=n).astype(np.float32)
        lo, hi = shard_bounds(n, world, rank)
        t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
        xtx, xty, yty = O.regression_suffstats(X[lo:hi], y[lo:hi])
        reg = _reduce_into(PackedStats.regression(d), None, {'xtx': t(xtx), 'xty': t(xty), 'yty': t([yty])},
                           hi - lo, None, None)
        nk, rx, rxx = O.weighted_suffstats(X[lo:hi], R[lo:hi])
        mix = _reduce_into(PackedStats.mixture(k, d), None, {'nk': t(nk), 'rx': t(rx), 'rxx': t(rxx)},
                           hi - lo, None, None)
        queue.put((rank, {name: v.numpy().copy() for name, v in reg.items()},
                   {name: v.numpy().copy() for name, v in mix.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_of_the_regression_and_mixture_layouts():
    import torch.multiprocessing as mp
    from oracle import closed_forms as O
    world = 2
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_all_layouts, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(7)
    n, d, k = 501, 4, 3
    X = rng.randn(n, d).astype(np.float32)
    y = rng.randn(n).astype(np.float32)
    R = rng.dirichlet(np.ones(k), size=n).astype(np.float32)
    xtx, xty, yty = O.regression_suffstats(X, y)
    nk, rx, rxx = O.weighted_suffstats(X, R)
    for _, reg, mix in results:
        np.testing.assert_allclose(reg['xtx'], xtx, rtol=1e-12)
        np.testing.assert_allclose(reg['xty'], xty, rtol=1e-12)
        np.testing.assert_allclose(reg['yty'][0], yty, rtol=1e-12)
        assert reg['count'][0] == n and mix['count'][0] == n
        np.testing.assert_allclose(mix['nk'], nk, rtol=1e-12)
        np.testing.assert_allclose(mix['rx'], rx, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(mix['rxx'], rxx, rtol=1e-12, atol=1e-12)


def test_reference_arm_under_torchrun_prints_one_line_from_rank_zero():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone runs the
    CPU baseline and prints ONE JSON line with "impl": "reference"; the other rank exits 0 silently."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29593', os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2',
           '--steps', '1', '--warmup', '1']
    run = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [ln for ln in run.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    record = json.loads(lines[0])
    assert record['impl'] == 'reference' and record['n_gpus'] == 2 and record['gpu_launches'] == 0
    assert record['e2e']['h2d_bytes_per_step'] == 0 and record['cpu_baseline']['kind'] in ('port', 'reference')
    assert record['value'] > 0 and record['higher_is_better'] is True
