"""Minimal numpy-backed stand-in for the handful of Theano symbols that
``/root/reference/bayesic/algebra.py`` touches.

TEST INFRASTRUCTURE ONLY.  It exists so the *unmodified* reference module can
be imported in the authoring container (real Theano is not installable here:
no network) in order to (a) run the reference's own 36 tests and (b) generate
the golden vectors under ``tests/golden/``.  Nothing in ``bayesic_b200/``
imports it.

The shim is a lazy expression graph: every symbolic variable carries a closure
``_ev(env)`` that evaluates it with numpy given ``env: {input Variable -> ndarray}``.
"""
import numpy as _np

from . import tensor  # noqa: F401  (theano.tensor)
from . import printing  # noqa: F401
from .tensor import Variable as _Variable


def function(inputs, output):
    """theano.function(inputs, output) -> callable(*arrays) (algebra.py:54)."""
    inputs = list(inputs)

    def fn(*arrays):
        if len(arrays) != len(inputs):
            raise TypeError("expected %d inputs, got %d" % (len(inputs), len(arrays)))
        env = {}
        for var, arr in zip(inputs, arrays):
            arr = _np.asarray(arr, dtype=var.dtype)
            if arr.ndim != var.ndim:
                raise TypeError("input %s: expected ndim %d, got %d" % (var.name, var.ndim, arr.ndim))
            env[var] = arr
        if output is None:      # e.g. algebra.py:1370-1373 forgets to return
            return None
        out = output._ev(env) if isinstance(output, _Variable) else output
        return _np.asarray(out)

    return fn
