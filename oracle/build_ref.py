"""Recipe for ``oracle/_ref``: an UNMODIFIED copy of the reference's importable package, so that the
reference itself can run where ``/root/reference`` does not exist (the GPU box).

    python -m oracle.build_ref

The reference is pure Python with no packaging (no setup.py / pyproject: ``pip install
/root/reference`` has nothing to install), so "installing" it is copying its package directory
byte for byte -- ``bayesic/__init__.py`` and ``bayesic/algebra.py``, the two files that import
(``bayesic/distribution/*`` does not parse) -- into ``oracle/_ref/bayesic/``.
``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the copy travels to the GPU box with the built ``.so``.  ``MANIFEST.json``
records the SHA-256 of every copied file next to the SHA-256 of its source; the loader refuses a
copy whose hashes do not match its manifest.

TEST INFRASTRUCTURE: used by ``bench.py --impl reference`` / the ``cpu_baseline`` leg and by tests.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = '/root/reference'
REF_DIR = os.path.join(HERE, '_ref')
FILES = ['bayesic/__init__.py', 'bayesic/algebra.py']


def _sha(path):
    with open(path, 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def build_ref(verbose=False):
    """Copy the reference package into oracle/_ref (no-op when the reference tree is absent and a
    valid copy already exists).  Returns the directory, or None when neither is available."""
    have_source = all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in FILES)
    if not have_source:
        return REF_DIR if verify_ref() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {'sha256': _sha(dst), 'source': src, 'source_sha256': _sha(src)}
        if manifest[rel]['sha256'] != manifest[rel]['source_sha256']:
            raise RuntimeError("copy of %s differs from its source" % rel)
    with open(os.path.join(REF_DIR, 'MANIFEST.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    if verbose:
        sys.stderr.write("oracle/_ref: copied %s from %s\n" % (', '.join(FILES), REFERENCE_ROOT))
    return REF_DIR


def verify_ref():
    """True when oracle/_ref holds every file of the manifest with the recorded hash."""
    try:
        with open(os.path.join(REF_DIR, 'MANIFEST.json')) as fh:
            manifest = json.load(fh)
        return set(manifest) == set(FILES) and all(
            _sha(os.path.join(REF_DIR, rel)) == entry['sha256'] == entry['source_sha256']
            for rel, entry in manifest.items())
    except (OSError, ValueError, KeyError):
        return False


if __name__ == '__main__':
    out = build_ref(verbose=True)
    print(out if out else 'reference unavailable')
