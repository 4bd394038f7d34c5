"""Minibatch streaming (the reference's intent at README.md:71-73): run an additive
pass over data that is NOT resident on the device.  Every output of the passes in ``stats`` is a
sum over the data axis, so a host-resident data set is cut into row chunks that flow through two
pinned staging buffers and two device buffers: the host->device copy of chunk i + 1 runs on a copy
stream while the kernels of chunk i run on the caller's stream, and the float64 partial results are
added on the device.  ``gather_rows`` is the on-device alternative for resident data: pick a
minibatch by index without a round trip through the host.

torch provides pinned memory, streams and events here -- plumbing; the arithmetic is the same CUDA
kernels as the resident path (no CPU fallback: without the library or a GPU everything raises).
"""

import numpy as np

from . import stats
from .backend import library as L

__all__ = ['streamed_pass', 'gather_rows']


def _as_host_tensor(a):
    torch = stats._torch()
    if isinstance(a, np.ndarray):
        if a.dtype != np.float32:
            raise TypeError("streamed_pass: host arrays must be float32, got %s" % a.dtype)
        return torch.from_numpy(np.ascontiguousarray(a))
    if isinstance(a, torch.Tensor) and not a.is_cuda:
        if a.dtype != torch.float32:
            raise TypeError("streamed_pass: host tensors must be float32, got %s" % a.dtype)
        return a.contiguous()
    raise TypeError("streamed_pass: arrays must be numpy arrays or CPU torch tensors")


def streamed_pass(pass_fn, arrays, chunk_rows, device=None):
    """``sum over chunks of pass_fn(*device_chunks)`` for host arrays that share axis 0.

    ``pass_fn`` takes one CUDA float32 tensor per array (a chunk of at most ``chunk_rows`` rows) and
    returns a tensor or a tuple of tensors that are additive over rows (e.g.
    ``stats.regression_suffstats``); the chunk results are accumulated in float64 on the device and
    returned in the same structure.  Copies and kernels of successive chunks overlap."""
    torch = stats._torch()
    hosts = [_as_host_tensor(a) for a in arrays]
    n = hosts[0].shape[0]
    if any(h.shape[0] != n for h in hosts):
        raise ValueError("streamed_pass: arrays disagree on the data axis")
    if chunk_rows < 1:
        raise ValueError("streamed_pass: chunk_rows must be positive")
    dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        rows = min(int(chunk_rows), max(n, 1))
        staging = [[None if h.is_pinned() else torch.empty((rows,) + tuple(h.shape[1:]), dtype=torch.float32).pin_memory()
                    for h in hosts] for _ in range(2)]
        on_dev = [[torch.empty((rows,) + tuple(h.shape[1:]), dtype=torch.float32, device=dev) for h in hosts]
                  for _ in range(2)]
        # The device buffers come from the caching allocator on the compute stream: earlier work on that
        # stream (e.g. the previous call's last kernels, whose buffers the allocator may hand back here)
        # must be finished before the copy stream writes them, and the blocks must not be reused before
        # the copy stream is done with them.
        copy.wait_stream(compute)
        for pair in on_dev:
            for t in pair:
                t.record_stream(copy)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        staged = [torch.cuda.Event() for _ in range(2)]
        totals, single = None, False
        n_chunks = (n + rows - 1) // rows

        def enqueue_copy(i):
            b = i & 1
            lo, hi = i * rows, min(n, (i + 1) * rows)
            if i >= 2:
                staged[b].synchronize()              # the staging buffer's previous copy has left the host
                copy.wait_event(consumed[b])         # and the device buffer's previous chunk has been used
            with torch.cuda.stream(copy):
                for j, h in enumerate(hosts):
                    src = h[lo:hi]
                    if staging[b][j] is not None:
                        staging[b][j][:hi - lo].copy_(src)
                        src = staging[b][j][:hi - lo]
                    on_dev[b][j][:hi - lo].copy_(src, non_blocking=True)
                staged[b].record(copy)
                copied[b].record(copy)

        if n_chunks > 0:
            enqueue_copy(0)
        for i in range(n_chunks):
            b = i & 1
            if i + 1 < n_chunks:
                enqueue_copy(i + 1)
            lo, hi = i * rows, min(n, (i + 1) * rows)
            compute.wait_event(copied[b])
            out = pass_fn(*[t[:hi - lo] for t in on_dev[b]])
            consumed[b].record(compute)
            if not isinstance(out, (tuple, list)):
                out, single = (out,), True
            if totals is None:
                totals = [o.to(torch.float64).clone() for o in out]
            else:
                for acc, o in zip(totals, out):
                    acc.add_(o.to(torch.float64))
        if totals is None:                           # no rows: the pass on an empty chunk defines the zeros
            out = pass_fn(*[t[:0] for t in on_dev[0]])
            if not isinstance(out, (tuple, list)):
                out, single = (out,), True
            totals = [o.to(torch.float64).clone() for o in out]
        copy.synchronize()                           # the sources (pinned or staged) are no longer being read
    return totals[0] if single else tuple(totals)


def gather_rows(X, index, check=True):
    """``X[index]`` for a resident float32 matrix and an int64 CUDA index vector, as one kernel
    (``bb_gather_rows``): the minibatch of an SVI step picked on the device.  Out-of-range indices
    are counted on the device (their rows are NaN); with ``check=True`` the count is read back (one
    4-byte copy, a host sync) and raised as ``IndexError``, with ``check=False`` nothing syncs and
    ``(rows, count_tensor)`` is returned."""
    torch = stats._torch()
    lib = L.load()
    X = stats._as_device_f32(X, 2, 'X')
    if not isinstance(index, torch.Tensor) or not index.is_cuda or index.dtype != torch.int64 or index.dim() != 1:
        raise TypeError("index must be a 1-d int64 CUDA torch.Tensor")
    index = index.contiguous()
    n, d = X.shape
    m = index.shape[0]
    dev = X.device
    with torch.cuda.device(dev):
        out = torch.empty((m, d), dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.bb_gather_rows(X.data_ptr(), n, d, index.data_ptr(), m, out.data_ptr(), bad.data_ptr(),
                                   stats._stream(dev)), 'bb_gather_rows')
        if not check:
            return out, bad
        n_bad = int(bad.item())
    if n_bad:
        raise IndexError("gather_rows: %d of %d indices are outside [0, %d)" % (n_bad, m, n))
    return out
