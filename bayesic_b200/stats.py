"""Fused sufficient-statistic / responsibility passes (Python face of the C-ABI).

These are the iid-summed exponential-family statistics the reference's distribution
sketch asks for -- ``ExpFamIndependentObservations.sufficient_statistics``
(``bayesic/distribution/base.py:328-332``) of ``MultivariateNormal``'s ``(x, x x^T)``
(``distribution/core.py:41-44``) -- computed in ONE pass over the data, plus the
mixture-responsibility pass and the ELBO term assembled from the statistics.

Inputs may be CUDA ``torch.Tensor``s (resident data, results stay on the device as
float64 tensors) or numpy arrays (streamed host->device in chunks, numpy results) --
the latter is the reference-facing call (numpy in, numpy out, ``algebra.py:55-56``).
No CPU fallback: everything here raises without the CUDA library and a GPU.
"""
import ctypes

import numpy as np

from .backend import library as L

__all__ = ['gaussian_suffstats', 'gaussian_expected_loglik', 'gaussian_suffstats_loglik', 'log_responsibilities',
           'responsibilities', 'responsibilities_split', 'weighted_suffstats_split', 'split_responsibilities_supported',
           'weighted_suffstats', 'regression_suffstats', 'row_projection', 'column_projection',
           'logistic_reparam_stats', 'logistic_reparam_supported', 'mixture_logits',
           'mixture_logits_supported', 'weighted_suffstats_from_logits', 'launch_count']

_scratch = {}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bayesic_b200: no CUDA device (there is no CPU fallback)")
    return torch


def _workspace(nbytes, device):
    """Grow-only kernel scratch, one per (device, CUDA stream, host thread): calls issued on two
    streams or from two threads never share a buffer.  A buffer that is outgrown is handed back to
    torch's stream-ordered caching allocator (allocated and used on the same stream, so its reuse is
    ordered after the kernels that still read it)."""
    import threading
    torch = _torch()
    key = (device, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _stream(device):
    return ctypes.c_void_p(_torch().cuda.current_stream(device).cuda_stream)


def launch_count():
    """Kernels launched by the library on this thread so far."""
    return int(L.load().bb_launch_count())


def _as_device_f32(t, ndim, what):
    torch = _torch()
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch.Tensor" % what)
    if t.dim() != ndim:
        raise TypeError("%s: expected ndim %d, got %d" % (what, ndim, t.dim()))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def gaussian_suffstats(X, out=None, chunk_rows=0):
    """``(n, sum_x[d], sum_xxT[d, d])`` of the rows of ``X[n, d]`` in float64.

    CUDA tensor in -> float64 CUDA tensors out (no synchronisation).
    numpy array in -> numpy float64 out through ``bb_suffstats_gaussian_host`` (chunked
    host->device streaming; pass pinned memory for full PCIe speed)."""
    torch = _torch()
    lib = L.load()
    if isinstance(X, torch.Tensor) and X.is_cuda:
        X = _as_device_f32(X, 2, 'X')
        n, d = X.shape
        with torch.cuda.device(X.device):
            if out is None:
                s1 = torch.empty(d, dtype=torch.float64, device=X.device)
                s2 = torch.empty((d, d), dtype=torch.float64, device=X.device)
            else:
                s1, s2 = out
            need = lib.bb_suffstats_gaussian_workspace(n, d)
            ws = _workspace(need, X.device)
            L.check(lib.bb_suffstats_gaussian(X.data_ptr(), n, d, s1.data_ptr(), s2.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream(X.device)),
                    'bb_suffstats_gaussian')
        return n, s1, s2
    if isinstance(X, torch.Tensor):          # pinned / pageable host tensor
        host = X.detach()
        if host.dtype != torch.float32 or not host.is_contiguous() or host.dim() != 2:
            raise TypeError("host X must be a contiguous float32 matrix")
        n, d = host.shape
        ptr = host.data_ptr()
    else:
        host = np.ascontiguousarray(X, dtype=np.float32)
        if host.ndim != 2:
            raise TypeError("X: expected ndim 2, got %d" % host.ndim)
        n, d = host.shape
        ptr = host.ctypes.data
    s1 = np.empty(d, dtype=np.float64)
    s2 = np.empty((d, d), dtype=np.float64)
    device = torch.device('cuda', torch.cuda.current_device())
    L.check(lib.bb_suffstats_gaussian_host(ptr, n, d, s1.ctypes.data, s2.ctypes.data,
                                           int(chunk_rows), _stream(device)),
            'bb_suffstats_gaussian_host')
    return n, s1, s2


def gaussian_suffstats_loglik(X, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, n_total=None, out=None):
    """``gaussian_suffstats`` and ``gaussian_expected_loglik`` in one call for resident data
    (``bb_suffstats_gaussian_loglik``: the log-likelihood is evaluated by the last block of the
    statistics' finalize kernel -- two launches instead of three).  ``e_lambda`` / ``e_lambda_mu``
    are float64 CUDA tensors; ``out = (s1, s2, loglik)`` reuses buffers.  Returns
    ``(n, s1, s2, loglik)``."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    n, d = X.shape
    dev = X.device
    for name, t, shape in (('e_lambda', e_lambda, (d, d)), ('e_lambda_mu', e_lambda_mu, (d,))):
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64 or tuple(t.shape) != shape:
            raise TypeError("%s must be a float64 CUDA tensor of shape %s" % (name, shape))
    e_lambda, e_lambda_mu = e_lambda.contiguous(), e_lambda_mu.contiguous()
    with torch.cuda.device(dev):
        if out is None:
            s1 = torch.empty(d, dtype=torch.float64, device=dev)
            s2 = torch.empty((d, d), dtype=torch.float64, device=dev)
            loglik = torch.empty(1, dtype=torch.float64, device=dev)
        else:
            s1, s2, loglik = out
        ws = _workspace(lib.bb_suffstats_gaussian_workspace(n, d), dev)
        L.check(lib.bb_suffstats_gaussian_loglik(X.data_ptr(), n, d, s1.data_ptr(), s2.data_ptr(),
                                                 float(n if n_total is None else n_total), e_lambda.data_ptr(),
                                                 e_lambda_mu.data_ptr(), float(e_mu_l_mu), float(e_logdet),
                                                 loglik.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
                'bb_suffstats_gaussian_loglik')
    return n, s1, s2, loglik


def gaussian_expected_loglik(n, s1, s2, e_lambda, e_lambda_mu, e_mu_l_mu, e_logdet, out=None):
    """E_q[sum_n log N(x_n | mu, Lambda^-1)] from device float64 statistics and the
    expectations of the natural parameters; returns a 1-element float64 CUDA tensor."""
    torch = _torch()
    lib = L.load()
    dev = None
    for t in (s2, s1, e_lambda, e_lambda_mu):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            dev = t.device
            break
    if dev is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    d = int(s2.shape[0])

    def f64(t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t, dtype=np.float64))
        return t.to(device=dev, dtype=torch.float64).contiguous()

    s1, s2, e_lambda, e_lambda_mu = f64(s1), f64(s2), f64(e_lambda), f64(e_lambda_mu)
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.bb_gaussian_expected_loglik(s1.data_ptr(), s2.data_ptr(), float(n),
                                                e_lambda.data_ptr(), e_lambda_mu.data_ptr(),
                                                float(e_mu_l_mu), float(e_logdet), d, out.data_ptr(),
                                                _stream(dev)), 'bb_gaussian_expected_loglik')
    return out


def log_responsibilities(logits, want_lse=True, want_sum=True, out=None):
    """``log r = logits - logsumexp(logits, axis=1)`` in one pass.  Returns
    ``(log_resp[n, k] float32, lse[n] float32 | None, sum_lse float64[1] | None)``."""
    torch = _torch()
    lib = L.load()
    logits = _as_device_f32(logits, 2, 'logits')
    n, k = logits.shape
    dev = logits.device
    with torch.cuda.device(dev):
        log_resp = out if out is not None else torch.empty_like(logits)
        lse = torch.empty(n, dtype=torch.float32, device=dev) if want_lse else None
        total = torch.empty(1, dtype=torch.float64, device=dev) if want_sum else None
        L.check(lib.bb_logsoftmax_rows(logits.data_ptr(), n, k, log_resp.data_ptr(),
                                       lse.data_ptr() if want_lse else None,
                                       total.data_ptr() if want_sum else None, _stream(dev)),
                'bb_logsoftmax_rows')
    return log_resp, lse, total


def responsibilities(logits, want_lse=True, want_sum=True, out=None):
    """``r = exp(logits - logsumexp(logits, axis=1))`` in one pass (``out`` may be ``logits`` itself).  Returns
    ``(resp[n, k] float32, lse[n] float32 | None, sum_lse float64[1] | None)``."""
    torch = _torch()
    lib = L.load()
    logits = _as_device_f32(logits, 2, 'logits')
    n, k = logits.shape
    dev = logits.device
    with torch.cuda.device(dev):
        resp = out if out is not None else torch.empty_like(logits)
        lse = torch.empty(n, dtype=torch.float32, device=dev) if want_lse else None
        total = torch.empty(1, dtype=torch.float64, device=dev) if want_sum else None
        L.check(lib.bb_softmax_rows(logits.data_ptr(), n, k, resp.data_ptr(), lse.data_ptr() if want_lse else None,
                                    total.data_ptr() if want_sum else None, _stream(dev)), 'bb_softmax_rows')
    return resp, lse, total


def split_responsibilities_supported(d, k):
    """Shapes the pre-split route (``responsibilities_split`` + ``weighted_suffstats_split``) serves."""
    return k in (256, 512, 768, 1024) and d % 8 == 0 and 8 <= d <= 64


def responsibilities_split(logits, want_lse=True, want_sum=True):
    """``r = exp(logits - logsumexp(logits, axis=1))`` written directly as the error-compensated BF16 operand tiles
    of the weighted-statistics kernel (``bb_softmax_rows_split``; opaque uint8 tensor, same bytes as float32 R).
    Returns ``(rsplit, lse[n] float32 | None, sum_lse float64[1] | None)``."""
    torch = _torch()
    lib = L.load()
    logits = _as_device_f32(logits, 2, 'logits')
    n, k = logits.shape
    dev = logits.device
    with torch.cuda.device(dev):
        nbytes = int(lib.bb_softmax_rows_split_bytes(n, k))
        if nbytes <= 0 and n > 0:
            raise ValueError("responsibilities_split: needs k in {256, 512, 768, 1024} (got %d)" % k)
        rsplit = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        lse = torch.empty(n, dtype=torch.float32, device=dev) if want_lse else None
        total = torch.empty(1, dtype=torch.float64, device=dev) if want_sum else None
        L.check(lib.bb_softmax_rows_split(logits.data_ptr(), n, k, rsplit.data_ptr(), lse.data_ptr() if want_lse else None,
                                          total.data_ptr() if want_sum else None, _stream(dev)), 'bb_softmax_rows_split')
    return rsplit, lse, total


def weighted_suffstats_split(X, rsplit, k):
    """``(N_k, sum_rx, sum_rxx)`` from ``X[n, d]`` and the pre-split responsibilities of ``responsibilities_split``."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    n, d = X.shape
    dev = X.device
    if rsplit.numel() < int(lib.bb_softmax_rows_split_bytes(n, k)):
        raise ValueError("weighted_suffstats_split: rsplit does not hold %d rows of %d components" % (n, k))
    with torch.cuda.device(dev):
        nk = torch.empty(k, dtype=torch.float64, device=dev)
        rx = torch.empty((k, d), dtype=torch.float64, device=dev)
        rxx = torch.empty((k, d, d), dtype=torch.float64, device=dev)
        ws = _workspace(lib.bb_suffstats_weighted_workspace(n, d, k), dev)
        L.check(lib.bb_suffstats_weighted_split(X.data_ptr(), rsplit.data_ptr(), n, d, k, nk.data_ptr(), rx.data_ptr(),
                                                rxx.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
                'bb_suffstats_weighted_split')
    return nk, rx, rxx


def weighted_suffstats(X, R):
    """``(N_k[k], sum_rx[k, d], sum_rxx[k, d, d])`` float64 CUDA tensors, one pass over (R, X)."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    R = _as_device_f32(R, 2, 'R')
    if X.shape[0] != R.shape[0]:
        raise ValueError("X and R disagree on the data axis (%d vs %d)" % (X.shape[0], R.shape[0]))
    n, d = X.shape
    k = R.shape[1]
    dev = X.device
    with torch.cuda.device(dev):
        nk = torch.empty(k, dtype=torch.float64, device=dev)
        rx = torch.empty((k, d), dtype=torch.float64, device=dev)
        rxx = torch.empty((k, d, d), dtype=torch.float64, device=dev)
        need = lib.bb_suffstats_weighted_workspace(n, d, k)
        ws = _workspace(need, dev)
        L.check(lib.bb_suffstats_weighted(X.data_ptr(), R.data_ptr(), n, d, k, nk.data_ptr(),
                                          rx.data_ptr(), rxx.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _stream(dev)), 'bb_suffstats_weighted')
    return nk, rx, rxx


def regression_suffstats(X, y=None):
    """``(X^T X [d, d], X^T y [d], y^T y [1])`` float64 CUDA tensors in one pass over ``X[n, d]``
    (``y`` omitted: just the Gram matrix).  ``d % 4 == 0``, ``64 < d <= 4096`` runs the tcgen05 CTA-pair kernel."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    n, d = X.shape
    dev = X.device
    if y is not None:
        y = _as_device_f32(y, 1, 'y')
        if y.shape[0] != n:
            raise ValueError("X and y disagree on the data axis (%d vs %d)" % (n, y.shape[0]))
    with torch.cuda.device(dev):
        xtx = torch.empty((d, d), dtype=torch.float64, device=dev)
        xty = torch.empty(d, dtype=torch.float64, device=dev) if y is not None else None
        yty = torch.empty(1, dtype=torch.float64, device=dev) if y is not None else None
        need = lib.bb_suffstats_regression_workspace(n, d)
        ws = _workspace(need, dev)
        L.check(lib.bb_suffstats_regression(X.data_ptr(), y.data_ptr() if y is not None else None, n, d,
                                            xtx.data_ptr(), xty.data_ptr() if y is not None else None,
                                            yty.data_ptr() if y is not None else None,
                                            ws.data_ptr(), ws.numel(), _stream(dev)),
                'bb_suffstats_regression')
    return xtx, xty, yty


def row_projection(X, W):
    """``Z[n, q] = X @ W.T`` (float32 CUDA tensor) on the tcgen05 row-projection kernel."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    W = _as_device_f32(W, 2, 'W')
    n, d = X.shape
    q = W.shape[0]
    if W.shape[1] != d:
        raise ValueError("X and W disagree on the feature axis (%d vs %d)" % (d, W.shape[1]))
    dev = X.device
    with torch.cuda.device(dev):
        Z = torch.empty((n, q), dtype=torch.float32, device=dev)
        ws = _workspace(lib.bb_rowproj_workspace(n, d, q), dev)
        L.check(lib.bb_rowproj(X.data_ptr(), W.data_ptr(), n, d, q, Z.data_ptr(), ws.data_ptr(), ws.numel(),
                               _stream(dev)), 'bb_rowproj')
    return Z


def column_projection(X, R):
    """``G[d, q] = X.T @ R`` over the data axis (float64 CUDA tensor) on the tcgen05 kernel."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    R = _as_device_f32(R, 2, 'R')
    n, d = X.shape
    q = R.shape[1]
    if R.shape[0] != n:
        raise ValueError("X and R disagree on the data axis (%d vs %d)" % (n, R.shape[0]))
    dev = X.device
    with torch.cuda.device(dev):
        G = torch.empty((d, q), dtype=torch.float64, device=dev)
        ws = _workspace(lib.bb_colproj_workspace(n, d, q), dev)
        L.check(lib.bb_colproj(X.data_ptr(), R.data_ptr(), n, d, q, G.data_ptr(), ws.data_ptr(), ws.numel(),
                               _stream(dev)), 'bb_colproj')
    return G


def logistic_reparam_supported(d, s):
    """Shapes the fused logistic pass serves (others go through the compiled plan)."""
    return (d % 128 == 0 and s % 64 == 0 and s * d <= 32768 and (d // 128) * (s // 64) <= 4)


def logistic_reparam_stats(X, y, W):
    """``(loglik[s], G[d, s])`` float64 CUDA tensors for S parameter draws ``W[s, d]``:
    ``loglik[s] = sum_n y z - log(1 + exp z)``, ``G = X.T @ (y - sigmoid(Z))``, ``Z = X @ W.T``."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    y = _as_device_f32(y, 1, 'y')
    W = _as_device_f32(W, 2, 'W')
    n, d = X.shape
    s = W.shape[0]
    if W.shape[1] != d or y.shape[0] != n:
        raise ValueError("logistic_reparam_stats: inconsistent shapes")
    dev = X.device
    with torch.cuda.device(dev):
        loglik = torch.empty(s, dtype=torch.float64, device=dev)
        G = torch.empty((d, s), dtype=torch.float64, device=dev)
        ws = _workspace(lib.bb_logistic_reparam_workspace(n, d, s), dev)
        L.check(lib.bb_logistic_reparam_pass(X.data_ptr(), y.data_ptr(), W.data_ptr(), n, d, s,
                                             loglik.data_ptr(), G.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _stream(dev)), 'bb_logistic_reparam_pass')
    return loglik, G


def mixture_logits_supported(d, k):
    """Shapes the tcgen05 mixture-logit kernel serves."""
    return d % 8 == 0 and 8 <= d <= 64 and k % 4 == 0 and 4 <= k <= 4096


def mixture_logits(X, U, t, c, want_lse=True, want_sum=True, upper_triangular=False):
    """``logits[n, k] = c[k] - 0.5 * ||U[k] @ x_n - t[k]||**2`` (float32 CUDA tensor) and,
    optionally, the row log-sum-exp ``lse[n]`` and its sum (float64[1]).
    ``upper_triangular=True`` promises ``U[k, j, i] == 0`` for ``i < j`` (Cholesky factors) and
    lets the kernel skip the tensor-core steps that would only multiply zeros."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    U = _as_device_f32(U, 3, 'U')
    t = _as_device_f32(t, 2, 't')
    c = _as_device_f32(c, 1, 'c')
    n, d = X.shape
    k = U.shape[0]
    if tuple(U.shape) != (k, d, d) or tuple(t.shape) != (k, d) or c.shape[0] != k:
        raise ValueError("mixture_logits: inconsistent shapes")
    dev = X.device
    with torch.cuda.device(dev):
        logits = torch.empty((n, k), dtype=torch.float32, device=dev)
        lse = torch.empty(n, dtype=torch.float32, device=dev) if want_lse else None
        total = torch.empty(1, dtype=torch.float64, device=dev) if want_sum else None
        ws = _workspace(lib.bb_mixture_logits_workspace(n, d, k), dev)
        L.check(lib.bb_mixture_logits(X.data_ptr(), U.data_ptr(), t.data_ptr(), c.data_ptr(), n, d, k,
                                      1 if upper_triangular else 0, logits.data_ptr(), lse.data_ptr() if want_lse else None,
                                      total.data_ptr() if want_sum else None, ws.data_ptr(), ws.numel(),
                                      _stream(dev)), 'bb_mixture_logits')
    return logits, lse, total


def weighted_suffstats_from_logits(X, logits, lse):
    """``(N_k, sum_rx, sum_rxx)`` with ``r[n, k] = exp(logits[n, k] - lse[n])`` formed on the fly
    inside the statistics kernel (the responsibility matrix is never written)."""
    torch = _torch()
    lib = L.load()
    X = _as_device_f32(X, 2, 'X')
    logits = _as_device_f32(logits, 2, 'logits')
    lse = _as_device_f32(lse, 1, 'lse')
    n, d = X.shape
    k = logits.shape[1]
    if logits.shape[0] != n or lse.shape[0] != n:
        raise ValueError("X, logits and lse disagree on the data axis")
    dev = X.device
    with torch.cuda.device(dev):
        nk = torch.empty(k, dtype=torch.float64, device=dev)
        rx = torch.empty((k, d), dtype=torch.float64, device=dev)
        rxx = torch.empty((k, d, d), dtype=torch.float64, device=dev)
        ws = _workspace(lib.bb_suffstats_weighted_workspace(n, d, k), dev)
        L.check(lib.bb_suffstats_weighted_from_logits(X.data_ptr(), logits.data_ptr(), lse.data_ptr(), n, d, k,
                                                      nk.data_ptr(), rx.data_ptr(), rxx.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), _stream(dev)),
                'bb_suffstats_weighted_from_logits')
    return nk, rx, rxx
