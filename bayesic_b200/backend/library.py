"""ctypes binding of ``libbayesic_b200.so`` (include/bayesic_b200.h).

The library is built in-tree by ``bayesic_b200.build``; if it is missing or cannot
be loaded every compute entry point raises ``RuntimeError`` -- there is no CPU
fallback (the reference's numpy/theano evaluation, ``bayesic/algebra.py:50-58``, is
exactly what this package replaces, not something it falls back to).
"""
import ctypes
import os

ABI_VERSION = 2
MAX_DIMS = 8
MAX_PARENTS = 8
MAX_IPARAMS = 40

# bb_node_kind
NODE_INPUT, NODE_SCALAR, NODE_SHAPE, NODE_EYE, NODE_SUM, NODE_MUL = 0, 1, 2, 3, 4, 5
NODE_DIMSHUFFLE, NODE_TENSORDOT, NODE_DIAGONAL, NODE_ELEMWISE = 6, 7, 8, 9
NODE_LOGSOFTMAX, NODE_SYRK, NODE_WEIGHTED_SCATTER, NODE_LOGDET = 20, 21, 22, 23
# bb_elemwise_op
OP_CODES = {'add': 0, 'mul': 1, 'log': 2, 'exp': 3, 'pow': 4, 'abs_': 5, 'lgamma': 6}

STATUS_NAMES = {0: 'BB_OK', 1: 'BB_ERR_INVALID', 2: 'BB_ERR_CUDA', 3: 'BB_ERR_UNSUPPORTED',
                4: 'BB_ERR_SHAPE', 5: 'BB_ERR_WORKSPACE'}


class NodeDesc(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('n_parents', ctypes.c_int32),
                ('parents', ctypes.c_int32 * MAX_PARENTS), ('n_iparams', ctypes.c_int32),
                ('iparams', ctypes.c_int32 * MAX_IPARAMS), ('fparam', ctypes.c_double)]


class TensorArg(ctypes.Structure):
    _fields_ = [('data', ctypes.c_void_p), ('ndim', ctypes.c_int32),
                ('is_host_scalar', ctypes.c_int32), ('shape', ctypes.c_int64 * MAX_DIMS),
                ('host_value', ctypes.c_double)]


class ResultInfo(ctypes.Structure):
    _fields_ = [('ndim', ctypes.c_int32), ('is_host_scalar', ctypes.c_int32),
                ('shape', ctypes.c_int64 * MAX_DIMS), ('host_value', ctypes.c_double)]


LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        'lib', 'libbayesic_b200.so')
if os.environ.get('BB_LIB_PATH'):          # developer knob: an experimental build variant of the same library
    LIB_PATH = os.environ['BB_LIB_PATH']

# every symbol include/bayesic_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
SIGNATURES = {
    'bb_abi_version': (ctypes.c_int, []),
    'bb_last_error': (ctypes.c_char_p, []),
    'bb_device_info': (ctypes.c_int, [ctypes.POINTER(_i32)] * 3),
    'bb_launch_count': (_i64, []),
    'bb_plan_create': (ctypes.c_int, [ctypes.POINTER(NodeDesc), _i32, ctypes.POINTER(_i32), _i32,
                                      _i32, ctypes.POINTER(_vp)]),
    'bb_plan_destroy': (ctypes.c_int, [_vp]),
    'bb_plan_infer': (ctypes.c_int, [_vp, ctypes.POINTER(TensorArg), _i32,
                                     ctypes.POINTER(ResultInfo), ctypes.POINTER(_i64)]),
    'bb_plan_execute': (ctypes.c_int, [_vp, ctypes.POINTER(TensorArg), _i32, ctypes.POINTER(_vp),
                                       _vp, _i64, _vp]),
    'bb_plan_last_launch_count': (ctypes.c_int, [_vp, ctypes.POINTER(_i32)]),
    'bb_suffstats_gaussian_workspace': (_i64, [_i64, _i32]),
    'bb_suffstats_gaussian': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp]),
    'bb_suffstats_gaussian_host': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    'bb_release_staging': (ctypes.c_int, []),
    'bb_gaussian_expected_loglik': (ctypes.c_int, [_vp, _vp, _dbl, _vp, _vp, _dbl, _dbl, _i32,
                                                   _vp, _vp]),
    'bb_suffstats_gaussian_loglik': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _dbl, _vp, _vp, _dbl, _dbl, _vp,
                                                    _vp, _i64, _vp]),
    'bb_suffstats_regression_workspace': (_i64, [_i64, _i32]),
    'bb_suffstats_regression': (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    'bb_rowproj_workspace': (_i64, [_i64, _i32, _i32]),
    'bb_rowproj': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    'bb_colproj_workspace': (_i64, [_i64, _i32, _i32]),
    'bb_colproj': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    'bb_logistic_reparam_workspace': (_i64, [_i64, _i32, _i32]),
    'bb_logistic_reparam_pass': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    'bb_mixture_logits_workspace': (_i64, [_i64, _i32, _i32]),
    'bb_mixture_logits': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    'bb_logsoftmax_rows': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    'bb_softmax_rows': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    'bb_softmax_rows_split_bytes': (ctypes.c_int64, [_i64, _i32]),
    'bb_softmax_rows_split': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    'bb_suffstats_weighted_split': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    'bb_suffstats_weighted_workspace': (_i64, [_i64, _i32, _i32]),
    'bb_suffstats_weighted_from_logits': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64,
                                                         _vp]),
    'bb_suffstats_weighted': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64,
                                             _vp]),
    'bb_gmm_global_update': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _dbl, _dbl, _dbl, _vp, _vp,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bb_comm_flag_bytes': (_i64, [_i32]),
    'bb_comm_create': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64,
                                      _dbl, ctypes.POINTER(_vp)]),
    'bb_comm_allreduce_sum': (ctypes.c_int, [_vp, _i64, _vp]),
    'bb_comm_status': (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _vp]),
    'bb_comm_destroy': (ctypes.c_int, [_vp]),
    'bb_gaussian_pass_create': (ctypes.c_int, [_i32, ctypes.POINTER(_vp)]),
    'bb_gaussian_pass_peer_bytes': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    'bb_gaussian_pass_attach_peers': (ctypes.c_int, [_vp, _i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _dbl]),
    'bb_gaussian_pass_run': (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp]),
    'bb_gaussian_pass_status': (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _vp]),
    'bb_gaussian_pass_destroy': (ctypes.c_int, [_vp]),
    'bb_gather_rows': (ctypes.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    'bb_svi_natural_blend': (ctypes.c_int, [_vp, _vp, _vp, _dbl, _dbl, _i64, _vp]),
    'bb_reparam_draws': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    'bb_reparam_gradient': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    'bb_adam_step': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _i64, _i32, _vp]),
}

_lib = None


def load():
    """The loaded library with typed entry points; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "bayesic_b200: CUDA library not built (%s missing). Run `python -m bayesic_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bb_abi_version() != ABI_VERSION:
        raise RuntimeError("bayesic_b200: ABI version mismatch")
    _lib = lib
    return lib


class BackendError(RuntimeError):
    pass


def check(status, what=''):
    """Turn a bb_status into the exception the reference's callers would see:
    shape/argument problems are ``ValueError`` (like numpy/theano at call time),
    everything else ``RuntimeError``."""
    if status == 0:
        return
    message = load().bb_last_error().decode(errors='replace')
    text = "%s%s: %s" % (what + ': ' if what else '', STATUS_NAMES.get(status, status), message)
    if status in (1, 4):
        raise ValueError(text)
    raise BackendError(text)
