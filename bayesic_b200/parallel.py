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

__all__ = ['shard_bounds', 'PackedStats', 'allreduce_packed', 'PeerReducer', 'gaussian_suffstats_sharded',
           'regression_suffstats_sharded', 'mixture_suffstats_sharded', 'logistic_reparam_sharded']


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


class PeerReducer(object):
    """One-shot all-reduce over NVLink peer memory fused with the consumer (``bb_allreduce_sum_p2p``,
    ``csrc/p2p_reduce.cu``) for the small per-minibatch payloads of this path: one single-CTA
    kernel per rank replaces "NCCL all-reduce, then the ELBO kernel".

    The peer mapping comes from ``torch.distributed._symmetric_memory`` (CUDA IPC plumbing): a
    double-buffered float64 payload tensor and a uint32 flag tensor per rank, each mapped into every
    process.  Write this rank's partial statistics into ``slot()`` (views of the payload for the
    coming epoch), then call ``reduce`` / ``reduce_loglik``; ``reduced`` holds the sum.  Raises at
    construction if symmetric memory is unavailable -- callers fall back to ``allreduce_packed``."""

    def __init__(self, layout, device, group=None, spin_limit_ms=2000.0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.layout = layout
        self.numel = int(layout.numel)
        self.stride = (self.numel + 31) // 32 * 32
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.spin_limit_ms = float(spin_limit_ms)
        with torch.cuda.device(self.device):
            self.payload = symm.empty(2 * self.stride, dtype=torch.float64, device=self.device)
            self.flags = symm.empty(max(64, self.world), dtype=torch.int32, device=self.device)
            self.payload.zero_()
            self.flags.zero_()
            self._payload_handle = symm.rendezvous(self.payload, group)
            self._flags_handle = symm.rendezvous(self.flags, group)
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                    # every rank's flags are zero before anyone publishes
            self.reduced = torch.zeros(self.numel, dtype=torch.float64, device=self.device)
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0

    def slot(self):
        """Views (by field name) of this rank's payload slot for the NEXT reduction."""
        lo = ((self.epoch + 1) & 1) * self.stride
        return self.layout.views(self.payload[lo:lo + self.numel])

    def set_constant(self, name, value):
        """Write a field that does not change from step to step (e.g. this rank's row count) into
        both payload slots once, instead of refilling it every step."""
        for parity in (0, 1):
            lo = parity * self.stride
            self.layout.views(self.payload[lo:lo + self.numel])[name].fill_(value)

    def _call(self, loglik):
        from . import stats
        from .backend import library as L
        lib = L.load()
        self.epoch += 1
        if loglik is None:
            e_lambda = e_lambda_mu = elbo = None
            e_mu_l_mu = e_logdet = 0.0
            d = 0
        else:
            e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, d, elbo = loglik
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(lib.bb_allreduce_sum_p2p(self._payload_handle.buffer_ptrs_dev, self._flags_handle.buffer_ptrs_dev,
                                         self.rank, self.world, self.numel, self.stride, self.epoch & 0xFFFFFFFF,
                                         self.spin_limit_ms, self.reduced.data_ptr(), self.status.data_ptr(),
                                         ptr(e_lambda), ptr(e_lambda_mu), float(e_mu_l_mu), float(e_logdet), int(d),
                                         ptr(elbo), stats._stream(self.device)), 'bb_allreduce_sum_p2p')
        return self.layout.views(self.reduced)

    def reduce(self):
        """Sum the ranks' current slots; returns views of the reduced buffer."""
        return self._call(None)

    def reduce_loglik(self, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, d, out):
        """Same, and the Gaussian expected log-likelihood of the reduced statistics into ``out``
        (float64[1]) in the same kernel (layout ``PackedStats.gaussian``)."""
        return self._call((e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, d, out))


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
    if buffer is None:
        buffer = layout.allocate(device)
    views = layout.views(buffer)
    for name, value in parts.items():
        views[name].copy_(value.reshape(views[name].shape))
    views['count'].fill_(float(n_local))
    allreduce_packed(buffer, group)
    return views


def regression_suffstats_sharded(X_local, y_local, buffer=None, group=None):
    """cfg4: ``{xtx, xty, yty, count}`` of the whole minibatch from this rank's rows: local fused
    pass (tcgen05 CTA pairs when D % 256 == 0), then one all-reduce of D^2 + D + 2 float64."""
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


def logistic_reparam_sharded(X_local, y_local, W, buffer=None, group=None):
    """cfg5: ``{loglik[S], G[D, S], count}`` summed over all ranks' rows for the replicated
    parameter draws ``W[S, D]``; one all-reduce of S (D + 1) + 1 float64."""
    from . import stats
    n_local, d = X_local.shape
    s = W.shape[0]
    loglik, G = stats.logistic_reparam_stats(X_local, y_local, W)
    return _reduce_into(PackedStats.logistic(d, s), X_local.device, {'loglik': loglik, 'G': G}, n_local,
                        buffer, group)
