"""Data-parallel plumbing: shard the data axis N over one process per GPU and combine the
per-GPU partial statistics with ONE all-reduce per minibatch.

The reference has no distributed code at all; what shards is the operation its distribution
sketch defines -- ``ExpFamIndependentObservations.sufficient_statistics`` sums the per-point
statistics over the iid axes (``bayesic/distribution/base.py:328-332``) -- so every output of
the pass is a sum over N and partial sums combine exactly.  Global
parameters are replicated and updated identically on every rank from the reduced statistics;
per-point outputs (log-responsibilities) stay sharded.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport; the payload is one
packed float64 buffer so the collective is a single latency-bound call.
"""
import numpy as np

__all__ = ['shard_bounds', 'PackedStats', 'allreduce_packed', 'PeerComm', 'peer_comm_for', 'GaussianPass',
           'gaussian_suffstats_sharded',
           'regression_suffstats_sharded', 'mixture_suffstats_sharded', 'mixture_local_step_sharded',
           'logistic_reparam_sharded']


def shard_bounds(n, world_size, rank):
    """Contiguous, balanced shard ``[start, stop)`` of ``range(n)`` for ``rank``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d of %d" % (rank, world_size))
    base, extra = divmod(int(n), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class PackedStats(object):
    """Layout of the per-minibatch all-reduce payload.  Fields are float64 and contiguous:
    ``PackedStats([('s2', (d, d)), ('s1', (d,)), ('count', (1,))])``."""

    def __init__(self, fields):
        self.fields = []
        offset = 0
        for name, shape in fields:
            size = int(np.prod(shape)) if len(shape) else 1
            self.fields.append((name, tuple(shape), offset, size))
            offset += size
        self.numel = offset

    @classmethod
    def gaussian(cls, d):
        return cls([('s2', (d, d)), ('s1', (d,)), ('count', (1,))])

    @classmethod
    def mixture(cls, k, d):
        return cls([('rxx', (k, d, d)), ('rx', (k, d)), ('nk', (k,)), ('sum_lse', (1,)), ('count', (1,))])

    @classmethod
    def regression(cls, d):
        return cls([('xtx', (d, d)), ('xty', (d,)), ('yty', (1,)), ('count', (1,))])

    @classmethod
    def logistic(cls, d, s):
        return cls([('G', (d, s)), ('loglik', (s,)), ('count', (1,))])

    def allocate(self, device=None):
        import torch
        return torch.zeros(self.numel, dtype=torch.float64, device=device)

    def views(self, buffer):
        """``{name: view}`` into a packed buffer (torch tensor or numpy array); no copies."""
        if buffer.shape[0] != self.numel:
            raise ValueError("packed buffer has %d elements, layout needs %d" % (buffer.shape[0], self.numel))
        return {name: buffer[offset:offset + size].reshape(shape)
                for name, shape, offset, size in self.fields}


def allreduce_packed(buffer, group=None):
    """Sum the packed statistics over all ranks, in place.  A no-op without an initialised
    process group (single GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    return buffer


def _symmetric(nbytes_or_numel, dtype, device, group):
    """A zeroed tensor in torch's symmetric memory, mapped into every rank of ``group`` -> (tensor,
    [device pointer of rank r's copy as seen from this process]).  CUDA IPC plumbing only."""
    import torch
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(int(nbytes_or_numel), dtype=dtype, device=device)
    t.zero_()
    handle = symm.rendezvous(t, group)
    ptrs = [int(p) for p in handle.buffer_ptrs]
    return t, ptrs, handle


def _ptr_array(ptrs):
    import ctypes
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class PeerComm(object):
    """All-reduce(sum) of a packed float64 payload over NVLink peer memory as ONE kernel per rank
    (``bb_comm_*``, ``csrc/p2p_reduce.cu``: two-shot, counterpart-CTA handshakes, no NCCL call).

    ``input`` / ``output`` are float64[capacity] tensors in symmetric memory: write this rank's
    partial statistics into ``input`` (or views of it: ``input_views(layout)``), call
    ``allreduce(count)``, read the sum from ``output``.  The output may be overwritten by the next
    all-reduce, so its consumers must be enqueued on the same stream before that call.  Construction is
    collective (every rank of ``group``) and raises where symmetric memory is unavailable -- callers
    fall back to ``allreduce_packed`` (NCCL)."""

    def __init__(self, capacity, device, group=None, spin_limit_ms=2000.0):
        import ctypes
        import torch
        import torch.distributed as dist
        from .backend import library as L
        lib = L.load()
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.capacity = (int(capacity) + 1) // 2 * 2
        self._comm = None
        with torch.cuda.device(self.device):
            self.input, in_ptrs, self._h_in = _symmetric(self.capacity, torch.float64, self.device, group)
            self.output, out_ptrs, self._h_out = _symmetric(self.capacity, torch.float64, self.device, group)
            self.flags, flag_ptrs, self._h_flags = _symmetric(lib.bb_comm_flag_bytes(self.world) // 4, torch.int32,
                                                              self.device, group)
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                    # every rank's flags are zero before anyone publishes
            handle = ctypes.c_void_p()
            L.check(lib.bb_comm_create(self.rank, self.world, _ptr_array(in_ptrs), _ptr_array(out_ptrs),
                                       _ptr_array(flag_ptrs), self.capacity, float(spin_limit_ms),
                                       ctypes.byref(handle)), 'bb_comm_create')
            self._comm = handle

    def __del__(self):
        try:
            if self._comm is not None:
                from .backend import library as L
                L.load().bb_comm_destroy(self._comm)
        except Exception:
            pass

    def input_views(self, layout):
        return layout.views(self.input[:layout.numel])

    def output_views(self, layout):
        return layout.views(self.output[:layout.numel])

    def allreduce(self, count):
        """Sum ``input[:count]`` over the ranks into ``output[:count]`` (stream-ordered, no sync)."""
        from . import stats
        from .backend import library as L
        L.check(L.load().bb_comm_allreduce_sum(self._comm, int(count), stats._stream(self.device)),
                'bb_comm_allreduce_sum')
        return self.output[:int(count)]

    def check(self):
        """Synchronise and raise if a peer was lost (the output is NaN-poisoned in that case)."""
        import ctypes
        from . import stats
        from .backend import library as L
        status = ctypes.c_int32(0)
        L.check(L.load().bb_comm_status(self._comm, ctypes.byref(status), stats._stream(self.device)),
                'bb_comm_status')
        if status.value:
            raise RuntimeError("bayesic_b200: peer rank %d did not answer the all-reduce within the spin limit; "
                               "the reduced statistics are NaN" % (status.value - 1))


_peer_comms = {}


def peer_comm_for(numel, device, group=None):
    """The process-wide ``PeerComm`` for this (group, device) with room for ``numel`` float64, created
    collectively on first use, or None where peer memory is unavailable (every rank agrees: the
    decision is itself all-reduced) or switched off (``BB_P2P_ALLREDUCE=0``)."""
    import os
    import torch
    import torch.distributed as dist
    if os.environ.get('BB_P2P_ALLREDUCE', '1') == '0' or not (dist.is_available() and dist.is_initialized()):
        return None
    if device is None or dist.get_world_size(group) < 2 or torch.device(device).type != 'cuda':
        return None
    key = (id(group) if group is not None else 0, str(device))
    entry = _peer_comms.get(key)
    if entry is not None and (entry is False or entry.capacity >= numel):
        return entry or None
    ok = 1.0
    comm = None
    try:
        comm = PeerComm(max(int(numel), 1 << 16), device, group)
    except Exception:                              # noqa: BLE001 -- any setup problem means "use NCCL"
        ok = 0.0
    flag = torch.tensor([ok], dtype=torch.float64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    _peer_comms[key] = comm if float(flag) == 1.0 else False
    return _peer_comms[key] or None


class GaussianPass(object):
    """The cfg2 step -- {count, sum x, sum x x^T} of this rank's rows, their combine over the ranks and
    the expected log-likelihood of the reduced statistics -- as ONE kernel launch per step
    (``bb_gaussian_pass_*``; the exchange is a per-slice push into the peers' receive buffers inside
    the statistics kernel).  With one process (or ``group`` of size 1) nothing is exchanged.

    ``run`` writes into the persistent float64 tensors ``s1[d]``, ``s2[d, d]``, ``count[1]``,
    ``loglik[1]`` (reduced over the ranks, bit-identical on every rank) and returns them."""

    def __init__(self, d, device, group=None, spin_limit_ms=2000.0):
        import ctypes
        import torch
        import torch.distributed as dist
        from .backend import library as L
        lib = L.load()
        self.d = int(d)
        self.device = torch.device(device)
        self._pass = None
        self.world, self.rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            group = group if group is not None else dist.group.WORLD
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        with torch.cuda.device(self.device):
            handle = ctypes.c_void_p()
            L.check(lib.bb_gaussian_pass_create(self.d, ctypes.byref(handle)), 'bb_gaussian_pass_create')
            self._pass = handle
            self.s1 = torch.zeros(self.d, dtype=torch.float64, device=self.device)
            self.s2 = torch.zeros((self.d, self.d), dtype=torch.float64, device=self.device)
            self.count = torch.zeros(1, dtype=torch.float64, device=self.device)
            self.loglik = torch.zeros(1, dtype=torch.float64, device=self.device)
            if self.world > 1:
                recv_bytes, flag_bytes = ctypes.c_int64(0), ctypes.c_int64(0)
                L.check(lib.bb_gaussian_pass_peer_bytes(self.d, self.world, ctypes.byref(recv_bytes),
                                                        ctypes.byref(flag_bytes)), 'bb_gaussian_pass_peer_bytes')
                self._recv, recv_ptrs, self._h_recv = _symmetric(recv_bytes.value // 8, torch.float64, self.device, group)
                self._flags, flag_ptrs, self._h_flags = _symmetric(flag_bytes.value // 4, torch.int32, self.device, group)
                torch.cuda.synchronize(self.device)
                dist.barrier(group)
                L.check(lib.bb_gaussian_pass_attach_peers(self._pass, self.rank, self.world, _ptr_array(recv_ptrs),
                                                          _ptr_array(flag_ptrs), float(spin_limit_ms)),
                        'bb_gaussian_pass_attach_peers')

    def __del__(self):
        try:
            if self._pass is not None:
                from .backend import library as L
                L.load().bb_gaussian_pass_destroy(self._pass)
        except Exception:
            pass

    def run(self, X, e_lambda=None, e_lambda_mu=None, e_mu_l_mu=0.0, e_logdet=0.0, n_total=None):
        """``X``: this rank's rows, CUDA float32 [n, d].  ``e_lambda`` / ``e_lambda_mu``: float64 CUDA
        tensors (omit both to skip the log-likelihood).  Stream-ordered, no sync."""
        from . import stats
        from .backend import library as L
        X = stats._as_device_f32(X, 2, 'X')
        n, d = X.shape
        if d != self.d:
            raise ValueError("GaussianPass was created for d = %d, got %d" % (self.d, d))
        want = e_lambda is not None
        L.check(L.load().bb_gaussian_pass_run(
            self._pass, X.data_ptr() if n else None, n, e_lambda.data_ptr() if want else None,
            e_lambda_mu.data_ptr() if want else None, float(e_mu_l_mu), float(e_logdet),
            float(n if n_total is None else n_total), self.s1.data_ptr(), self.s2.data_ptr(), self.count.data_ptr(),
            self.loglik.data_ptr() if want else None, stats._stream(self.device)), 'bb_gaussian_pass_run')
        return self.count, self.s1, self.s2, (self.loglik if want else None)

    def check(self):
        """Synchronise and raise if a peer was lost (the outputs are NaN-poisoned in that case)."""
        import ctypes
        from . import stats
        from .backend import library as L
        status = ctypes.c_int32(0)
        L.check(L.load().bb_gaussian_pass_status(self._pass, ctypes.byref(status), stats._stream(self.device)),
                'bb_gaussian_pass_status')
        if status.value:
            raise RuntimeError("bayesic_b200: peer rank %d did not answer within the spin limit; the statistics "
                               "are NaN" % (status.value - 1))


def gaussian_suffstats_sharded(X_local, layout=None, buffer=None, group=None):
    """``{count, s1, s2}`` of the WHOLE data set given this rank's shard ``X_local[n_local, d]``
    (CUDA float32): local one-pass kernel, then one all-reduce.  Returns the dict of views into
    the packed float64 buffer."""
    from . import stats
    n_local, d = X_local.shape
    layout = layout or PackedStats.gaussian(d)
    if buffer is None:
        buffer = layout.allocate(X_local.device)
    views = layout.views(buffer)
    stats.gaussian_suffstats(X_local, out=(views['s1'], views['s2']))
    views['count'].fill_(float(n_local))
    allreduce_packed(buffer, group)
    return views


def _reduce_into(layout, device, parts, n_local, buffer, group):
    """Pack this rank's parts, sum over the ranks, return views of the reduced payload.  On GPUs with
    peer memory the combine is the library's own one-kernel all-reduce (``PeerComm``); NCCL / gloo
    through ``torch.distributed`` otherwise."""
    comm = peer_comm_for(layout.numel, device, group) if buffer is None else None
    if comm is not None:
        views = comm.input_views(layout)
    else:
        if buffer is None:
            buffer = layout.allocate(device)
        views = layout.views(buffer)
    for name, value in parts.items():
        views[name].copy_(value.reshape(views[name].shape))
    for name, shape, offset, size in layout.fields:
        if name not in parts and name != 'count':
            views[name].zero_()
    views['count'].fill_(float(n_local))
    if comm is not None:
        comm.allreduce(layout.numel)
        return {k: v.clone() for k, v in comm.output_views(layout).items()}
    allreduce_packed(buffer, group)
    return views


def regression_suffstats_sharded(X_local, y_local, buffer=None, group=None):
    """cfg4: ``{xtx, xty, yty, count}`` of the whole minibatch from this rank's rows: local fused
    pass (tcgen05 CTA pairs when D % 4 == 0, D > 64), then one all-reduce of D^2 + D + 2 float64."""
    from . import stats
    n_local, d = X_local.shape
    xtx, xty, yty = stats.regression_suffstats(X_local, y_local)
    return _reduce_into(PackedStats.regression(d), X_local.device, {'xtx': xtx, 'xty': xty, 'yty': yty},
                        n_local, buffer, group)


def mixture_suffstats_sharded(X_local, R_local, sum_lse_local=None, buffer=None, group=None):
    """cfg3: ``{nk, rx, rxx, sum_lse, count}`` of the whole data set from this rank's rows and
    responsibilities (which stay sharded); one all-reduce of K (D^2 + D + 1) + 2 float64."""
    from . import stats
    n_local, d = X_local.shape
    k = R_local.shape[1]
    nk, rx, rxx = stats.weighted_suffstats(X_local, R_local)
    parts = {'nk': nk, 'rx': rx, 'rxx': rxx}
    if sum_lse_local is not None:
        parts['sum_lse'] = sum_lse_local
    return _reduce_into(PackedStats.mixture(k, d), X_local.device, parts, n_local, buffer, group)


def mixture_local_step_sharded(X_local, U, t, c, buffer=None, group=None):
    """cfg3, the whole VMP local step on this rank's rows from the replicated whitened parameters
    (``passes.GmmStep.local_step``: logits, responsibilities as operand tiles, statistics on CTA pairs where the
    shapes allow) and one all-reduce of ``{nk, rx, rxx, sum_lse, count}``; the per-row outputs stay sharded."""
    from .passes import GmmStep
    n_local, d = X_local.shape
    k = U.shape[0]
    out = GmmStep.local_step(X_local, U, t, c)
    parts = {'nk': out['nk'], 'rx': out['rx'], 'rxx': out['rxx'], 'sum_lse': out['sum_lse']}
    reduced = _reduce_into(PackedStats.mixture(k, d), X_local.device, parts, n_local, buffer, group)
    reduced['lse'] = out['lse']
    return reduced


def logistic_reparam_sharded(X_local, y_local, W, buffer=None, group=None):
    """cfg5: ``{loglik[S], G[D, S], count}`` summed over all ranks' rows for the replicated
    parameter draws ``W[S, D]``; one all-reduce of S (D + 1) + 1 float64."""
    from . import stats
    n_local, d = X_local.shape
    s = W.shape[0]
    loglik, G = stats.logistic_reparam_stats(X_local, y_local, W)
    return _reduce_into(PackedStats.logistic(d, s), X_local.device, {'loglik': loglik, 'G': G}, n_local,
                        buffer, group)
