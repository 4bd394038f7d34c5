"""Builds ``bayesic_b200/lib/libbayesic_b200.so`` in-tree with nvcc for sm_100a.

    python -m bayesic_b200.build [--force]

The library is a plain C-ABI shared object (include/bayesic_b200.h); it links the
static CUDA runtime and gets the one driver entry point it needs
(cuTensorMapEncodeTiled) through cudaGetDriverEntryPoint, so it has no link-time
dependency on libcuda and builds on a box without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libbayesic_b200.so')
SOURCES = ['runtime.cu', 'generic_kernels.cu', 'suffstats_sm100.cu', 'weighted_sm100.cu', 'weighted_pairs_sm100.cu', 'gram_sm100.cu', 'rowproj_sm100.cu', 'colproj_sm100.cu', 'logistic_fused2_sm100.cu', 'mixture_logits_sm100.cu', 'mixture_kernels.cu',
           'stats_kernels.cu', 'update_kernels.cu', 'linalg_kernels.cu', 'p2p_reduce.cu', 'executor.cu', 'api.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--use_fast_math=false']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build the CUDA library')
    return nvcc


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'bayesic_b200.h'))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False, defines=(), lib_path=None):
    """Compile every .cu into one shared library; returns its path.  ``defines`` (e.g.
    ``['BB_CHAIN_DIV=4']``) with ``lib_path`` builds an experimental variant next to the product
    library (loaded through the BB_LIB_PATH environment variable by developer scripts only)."""
    variant = bool(defines)
    if variant and not lib_path:
        raise ValueError("a variant build needs its own lib_path")
    if not variant and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, 'build' if not variant else 'build_' + os.path.basename(lib_path).replace('.', '_'))
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')] + ['-D' + d for d in defines]
    out_path = lib_path if variant else LIB_PATH
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [nvcc] + flags + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        text = out.decode(errors='replace')
        if proc.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed on %s:\n%s\n' % (src, text))
        elif verbose or text.strip():
            sys.stderr.write('[%s]\n%s\n' % (src, text))
    if failed:
        raise RuntimeError('nvcc failed; see messages above')
    link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out_path] + objs + \
           ['-cudart', 'static', '-Xlinker', '--exclude-libs=ALL']
    subprocess.check_call(link)
    return out_path


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--out=')]
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, defines=defs,
                 lib_path=os.path.join(LIB_DIR, outs[0]) if outs else None)
    print(path)
