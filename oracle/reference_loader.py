"""Import the UNMODIFIED reference ``bayesic.algebra`` through the numpy Theano shim.

Source of the module, in this order: ``/root/reference`` (the authoring container), else the
byte-identical copy under ``oracle/_ref`` that ``oracle/build_ref.py`` makes at ``build()`` time (the
GPU box, where ``/root/reference`` does not exist; hashes checked against the manifest).
TEST INFRASTRUCTURE: used by ``oracle/make_golden.py``, ``tests/test_reference_crosscheck.py`` and the
reference / ``cpu_baseline`` legs of ``bench.py`` -- never by ``bayesic_b200/``."""
import os
import sys

from . import build_ref as _build_ref

REFERENCE_ROOT = '/root/reference'
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'theano_shim')


def reference_root():
    """Directory that contains the reference's ``bayesic`` package, or None."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, 'bayesic', 'algebra.py')):
        return REFERENCE_ROOT
    if _build_ref.verify_ref():
        return _build_ref.REF_DIR
    return None


def reference_available():
    return reference_root() is not None


def load_reference_algebra():
    """Returns the reference's ``bayesic.algebra`` module object."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present at %s and no valid copy under %s (run python -m oracle.build_ref "
                           "where the reference tree exists)" % (REFERENCE_ROOT, _build_ref.REF_DIR))
    sys.dont_write_bytecode = True          # /root/reference is read-only
    for path in (root, _SHIM):
        if path not in sys.path:
            sys.path.insert(0, path)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', SyntaxWarning)
        import bayesic.algebra as ref_algebra
    return ref_algebra
