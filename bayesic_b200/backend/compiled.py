"""``expr.compile()`` -> ``f(**inputs)`` on the sm_100a executor.

Mirrors ``Expression.compile`` of the reference (``bayesic/algebra.py:50-58``):
inputs are passed by var name, the result is one array per compiled expression.
What differs is only what runs underneath -- ``bb_plan_execute`` instead of a
Theano function -- and two conveniences the hot path needs:

* inputs may be CUDA ``torch.Tensor``s (data resident in HBM); results then stay
  on the device.  numpy / list inputs are copied host->device per call and the
  result comes back as numpy, exactly like the reference;
* :func:`compile_many` compiles several expressions into one plan so statistics
  that read the same data share sub-trees (the reference has one output per
  ``compile()``).

There is no CPU evaluation path: without the CUDA library or a GPU, calling the
compiled function raises.
"""
import ctypes

import numpy as np

from . import library as L
from .lowering import lower_plans

__all__ = ['compile_expressions', 'compile_many', 'CompiledPlan']

_FLOAT32 = np.dtype('float32')


def _merge_input_types(exprs):
    merged = {}
    for expr in exprs:
        for name, typ in expr.input_types.items():
            seen = merged.setdefault(name, typ)
            if seen != typ:
                raise TypeError("same input %s occurs with different types %s, %s" % (name, seen, typ))
    return merged


class _Workspace(object):
    """Grow-only device scratch owned by a compiled plan (torch is used for allocation only)."""

    def __init__(self):
        self.buffer = None

    def get(self, nbytes, device):
        import torch
        if self.buffer is None or self.buffer.numel() < nbytes or self.buffer.device != device:
            self.buffer = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buffer


class CompiledPlan(object):
    def __init__(self, exprs, fuse=True):
        self.expressions = list(exprs)
        self.input_types = _merge_input_types(self.expressions)
        trees = [e.lower() for e in self.expressions]
        self.plan_trees = trees
        self.lowered = lower_plans(trees, self.input_types, fuse=fuse)
        self._nodes = self.lowered.as_ctypes()
        self._handle = None
        self._workspace = _Workspace()
        self._bound = {}           # slot -> device tensor of an ndarray literal
        self._shape_cache = {}
        self.last_launches = 0

    # ---- native handle ---------------------------------------------------
    def _native(self):
        if self._handle is None:
            lib = L.load()
            outs = (ctypes.c_int32 * len(self.lowered.outputs))(*self.lowered.outputs)
            handle = ctypes.c_void_p()
            L.check(lib.bb_plan_create(self._nodes, len(self.lowered.nodes), outs,
                                       len(self.lowered.outputs), len(self.lowered.input_names),
                                       ctypes.byref(handle)), 'bb_plan_create')
            self._handle = handle
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                L.load().bb_plan_destroy(self._handle)
        except Exception:
            pass

    # ---- per-call ------------------------------------------------------------
    def _device(self, inputs):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("bayesic_b200: no CUDA device; compiled plans only run on the GPU "
                               "(there is no CPU fallback)")
        for value in inputs.values():
            if isinstance(value, torch.Tensor) and value.is_cuda:
                return value.device, True
        return torch.device('cuda', torch.cuda.current_device()), False

    def _prepare(self, name, value, device, keep):
        """-> (TensorArg fields) for one named input."""
        import torch
        dtype, ndim = self.input_types[name]
        if isinstance(value, torch.Tensor):
            if value.dim() != ndim:
                raise TypeError("input %s: expected ndim %d, got %d" % (name, ndim, value.dim()))
            if ndim == 0:
                return None, (), float(value.item())
            t = value
            if not t.is_cuda:
                t = t.to(device)
            if t.dtype != torch.float32:
                t = t.to(torch.float32)
            if not t.is_contiguous():
                t = t.contiguous()
            keep.append(t)
            return t.data_ptr(), tuple(t.shape), None
        arr = np.asarray(value, dtype=np.dtype(dtype))
        if arr.ndim != ndim:
            raise TypeError("input %s: expected ndim %d, got %d" % (name, ndim, arr.ndim))
        if ndim == 0:
            return None, (), float(arr)
        if arr.dtype.kind in 'iu' and arr.size and int(np.abs(arr).max()) > (1 << 24):
            # the device path computes in float32: integers beyond 2^24 would be rounded silently, where
            # Theano (the reference's evaluator, algebra.py:34-40) keeps the integer dtype
            raise ValueError("input %s: integer values beyond 2**24 are not exact in the float32 device path" % name)
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=_FLOAT32)).to(device)
        keep.append(t)
        return t.data_ptr(), tuple(t.shape), None

    def __call__(self, **inputs):
        import torch
        lib = L.load()
        handle = self._native()
        device, on_device = self._device(inputs)
        missing = [n for n in self.input_types if n not in inputs]
        if missing:
            raise KeyError(missing[0])
        keep = []
        n_slots = len(self.lowered.input_names)
        args = (L.TensorArg * max(n_slots, 1))()
        with torch.cuda.device(device):
            for slot, name in enumerate(self.lowered.input_names):
                if name == '':
                    if slot not in self._bound or self._bound[slot].device != device:
                        self._bound[slot] = torch.from_numpy(self.lowered.bound_constants[slot]).to(device)
                    t = self._bound[slot]
                    ptr, shp, host = t.data_ptr(), tuple(t.shape), None
                else:
                    ptr, shp, host = self._prepare(name, inputs[name], device, keep)
                arg = args[slot]
                arg.ndim = len(shp)
                for i, e in enumerate(shp):
                    arg.shape[i] = e
                if host is not None:
                    arg.is_host_scalar, arg.host_value, arg.data = 1, host, None
                else:
                    arg.is_host_scalar, arg.host_value, arg.data = 0, 0.0, ptr
            n_out = len(self.lowered.outputs)
            key = tuple((a.ndim, a.is_host_scalar, tuple(a.shape[:a.ndim]),
                         a.host_value if a.is_host_scalar else 0.0) for a in args[:n_slots])
            cached = self._shape_cache.get(key)
            if cached is None:
                infos = (L.ResultInfo * n_out)()
                ws_bytes = ctypes.c_int64(0)
                L.check(lib.bb_plan_infer(handle, args, n_slots, infos, ctypes.byref(ws_bytes)),
                        'bb_plan_infer')
                cached = ([(i.ndim, bool(i.is_host_scalar), tuple(i.shape[:i.ndim]), i.host_value)
                           for i in infos], ws_bytes.value)
                if len(self._shape_cache) < 64:
                    self._shape_cache[key] = cached
            infos, ws_bytes = cached
            outs, out_ptrs = [], (ctypes.c_void_p * n_out)()
            for j, (ndim, is_host, shp, _) in enumerate(infos):
                if is_host:
                    outs.append(None)
                    out_ptrs[j] = None
                else:
                    t = torch.empty(shp, dtype=torch.float32, device=device)
                    outs.append(t)
                    out_ptrs[j] = t.data_ptr()
            ws = self._workspace.get(ws_bytes, device)
            stream = torch.cuda.current_stream(device).cuda_stream
            L.check(lib.bb_plan_execute(handle, args, n_slots, out_ptrs, ws.data_ptr(), ws.numel(),
                                        ctypes.c_void_p(stream)), 'bb_plan_execute')
            count = ctypes.c_int32(0)
            lib.bb_plan_last_launch_count(handle, ctypes.byref(count))
            self.last_launches = count.value
            results = []
            for j, (ndim, is_host, shp, host_value) in enumerate(infos):
                as_int = self.lowered.integer_result[j]
                if is_host:
                    value = np.int64(round(host_value)) if as_int else np.float32(host_value)
                    results.append(np.full(shp, value) if ndim else value)
                elif on_device:
                    results.append(outs[j].round().to(torch.int64) if as_int else outs[j])
                else:
                    host = outs[j].cpu().numpy()
                    results.append(np.rint(host).astype(np.int64) if as_int else host)
        return results


def compile_expressions(exprs, single=False, fuse=True):
    """Compile into one plan; returns ``f(**inputs)`` giving one result (``single``)
    or a tuple of results."""
    plan = CompiledPlan(exprs, fuse=fuse)

    def f(**inputs):
        results = plan(**inputs)
        return results[0] if single else tuple(results)

    f.plan = plan
    f.theano_fn = plan          # reference-era attribute name (algebra.py:57)
    f.input_types = plan.input_types
    return f


def compile_many(exprs, fuse=True):
    """Several expressions, one plan, one call: ``f(**inputs) -> tuple``."""
    return compile_expressions(list(exprs), single=False, fuse=fuse)
