"""Import the UNMODIFIED reference ``bayesic.algebra`` through the numpy Theano
shim.  Authoring container only: ``/root/reference`` does not exist on the GPU
box, so nothing reachable from ``pytest -m gpu``, ``smoke()`` or ``bench.py`` may
call this.  Used by ``oracle/make_golden.py`` and by the optional
``tests/test_reference_crosscheck.py`` (skipped when the reference is absent)."""
import os
import sys

REFERENCE_ROOT = '/root/reference'
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'theano_shim')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'bayesic', 'algebra.py'))


def load_reference_algebra():
    """Returns the reference's ``bayesic.algebra`` module object."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True          # /root/reference is read-only
    for path in (REFERENCE_ROOT, _SHIM):
        if path not in sys.path:
            sys.path.insert(0, path)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', SyntaxWarning)
        import bayesic.algebra as ref_algebra
    return ref_algebra
