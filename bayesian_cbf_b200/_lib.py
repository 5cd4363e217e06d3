"""ctypes binding of libbcbf.so (include/bcbf.h).  There is no fallback: if the library cannot be loaded
every op raises, and the ops themselves refuse non-CUDA tensors."""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libbcbf.so')

BCBF_OK = 0
BCBF_ERR_INVALID = -1
BCBF_ERR_CUDA = -2
BCBF_ERR_NOT_PD = -3
BCBF_ERR_NOT_FITTED = -4
MAX_N_DIM = 8
MAX_P_DIM = 4
BLOCK = 128


class BcbfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class NotPositiveDefiniteError(BcbfError):
    """Raised when the Cholesky meets a non-positive pivot.  Subclass of RuntimeError so that the
    reference's `except RuntimeError` jitter-retry loop (control_affine_model.py:913) keeps working."""


class Hyper(Structure):
    _fields_ = [('n', c_int), ('p', c_int), ('outputscale', c_double),
                ('lengthscale', c_double * MAX_N_DIM),
                ('A', c_double * (MAX_N_DIM * MAX_N_DIM)),
                ('B', c_double * (MAX_P_DIM * MAX_P_DIM)),
                ('C', c_double * (MAX_P_DIM * MAX_N_DIM))]


_P = c_void_p
_SIGNATURES = {
    'bcbf_last_error': (c_char_p, []),
    'bcbf_version': (c_int, []),
    'bcbf_padded': (c_int, [c_int]),
    'bcbf_launch_count': (ctypes.c_ulonglong, []),
    'bcbf_profile_enable': (c_int, [c_int]),
    'bcbf_profile_read': (c_int, [POINTER(c_double), POINTER(c_int)]),
    'bcbf_debug_counters': (c_int, [c_int, POINTER(ctypes.c_ulonglong * 8)]),
    'bcbf_dinv_elems': (c_longlong, [c_int]),
    'bcbf_gram_train': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    'bcbf_gram_train_lower': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    'bcbf_gram_resid_scratch_elems': (c_longlong, [c_int]),
    'bcbf_gram_resid': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_double, _P, c_int, _P, c_int, c_int,
                                _P, c_int, _P, c_longlong, _P]),
    'bcbf_gram_resid_stored': (c_int, [_P, c_int, c_int, _P, c_double, _P, c_int, _P, c_int, c_int, _P, c_int, _P,
                                       c_longlong, _P]),
    'bcbf_alpha_refine_ws': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_double, _P, c_int, c_int, _P,
                                     c_int, c_int, c_int, _P, _P, c_longlong, _P, c_int, _P]),
    'bcbf_alpha_refine_scratch_elems': (c_longlong, [c_int, c_int, c_int]),
    'bcbf_alpha_refine': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_double, _P, c_int, c_int, _P, c_int,
                                  c_int, c_int, _P, _P, c_longlong, _P]),
    'bcbf_cross_gram': (c_int, [_P, _P, _P, c_double, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    'bcbf_gram_ca': (c_int, [_P, _P, c_int, _P, _P, c_int, _P, _P, c_double, c_int, c_int, _P, c_int, _P]),
    'bcbf_ca_weight': (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, _P]),
    'bcbf_gemm': (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, _P, c_int, _P, c_int, c_double, _P, c_int, _P]),
    'bcbf_gram_train_backward': (c_int, [_P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, c_int,
                                         _P, c_longlong, _P, _P]),
    'bcbf_gram_backward_layout': (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    'bcbf_rbf_blocks': (c_int, [_P, _P, _P, c_double, c_int, c_int, c_int, _P, _P, _P, _P]),
    'bcbf_potrf': (c_int, [_P, c_int, c_int, c_int, _P, c_double, _P, _P, _P]),
    'bcbf_check_info': (c_int, [_P, _P]),
    'bcbf_trtri': (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    'bcbf_trmm_lower': (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, c_double, c_double, _P, c_int, _P]),
    'bcbf_posterior_blocks': (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_double, c_int, c_int, c_int,
                                      _P, _P, _P]),
    'bcbf_contract_u': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    'bcbf_posterior_fu': (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, c_double, c_int, c_int, c_int,
                                  _P, _P, _P]),
    'bcbf_cbc1_terms': (c_int, [_P, _P, _P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    'bcbf_socp_factor': (c_int, [_P, c_int, c_int, c_double, _P, _P, _P, _P]),
    'bcbf_ens_gram': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'bcbf_potrf_batched': (c_int, [_P, c_int, c_int, c_int, _P, c_double, _P, _P, c_int, _P]),
    'bcbf_trtri_batched': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    'bcbf_trmm_lower_batched': (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, c_double, c_double, _P, c_int, c_int,
                                        _P]),
    'bcbf_ens_prep': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    'bcbf_ens_w': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P]),
    'bcbf_ens_gram_backward': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P,
                                       c_longlong, _P, _P]),
    'bcbf_gemm_batched': (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, _P, c_int, c_longlong, _P, c_int, c_longlong,
                                  c_double, _P, c_int, c_longlong, c_int, _P]),
    'bcbf_ens_transpose': (c_int, [_P, _P, c_int, c_int, _P]),
    'bcbf_ens_posterior': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    'bcbf_socp_solve': (c_int, [c_int, c_int, c_int, c_int, c_double, _P, c_int, _P, _P, _P, _P, _P, c_double, _P, _P, _P, _P]),
    'bcbf_socp_solve_lin': (c_int, [c_int, c_int, c_int, c_int, c_double, _P, c_int, _P, _P, _P, _P, _P, _P, c_double, _P, _P, _P,
                                    _P]),
    'bcbf_model_create': (c_int, [POINTER(c_void_p), c_int]),
    'bcbf_model_destroy': (None, [c_void_p]),
    'bcbf_model_fit': (c_int, [c_void_p, POINTER(Hyper), _P, _P, _P, c_int, _P, c_double]),
    'bcbf_model_query': (c_int, [c_void_p, _P, _P, c_int, _P, _P, _P, _P]),
    'bcbf_model_query_device': (c_int, [c_void_p, _P, _P, c_int, _P, _P, _P, _P, _P]),
    'bcbf_model_state': (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)] + [POINTER(c_void_p)] * 6),
    'bcbf_model_alloc_state': (c_int, [c_void_p, POINTER(Hyper), c_int]),
    'bcbf_model_adopt': (c_int, [c_void_p]),
    'bcbf_packed_lower_elems': (c_longlong, [c_int]),
    'bcbf_pack_lower': (c_int, [_P, c_int, c_int, _P, _P]),
    'bcbf_unpack_lower': (c_int, [_P, c_int, _P, c_int, _P]),
    'bcbf_model_fit_timing': (c_int, [c_void_p, POINTER(c_double * 5)]),
    'bcbf_oz_factor_bytes': (c_longlong, [c_int]),
    'bcbf_oz_max_npad': (c_int, []),
    'bcbf_oz_split_factor': (c_int, [_P, c_int, c_int, _P, _P, _P]),
    'bcbf_posterior_var_i8': (c_int, [_P, _P, c_int, _P, c_int, _P, _P, c_double, c_int, c_int, _P, _P]),
    'bcbf_posterior_blocks_i8': (c_int, [_P, _P, c_int, _P, c_int, _P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, _P,
                                         _P]),
    'bcbf_oz_factor_bytes_d': (c_longlong, [c_int, c_int]),
    'bcbf_oz_split_factor_d': (c_int, [_P, c_int, c_int, _P, _P, c_int, _P]),
    'bcbf_posterior_blocks_i8_d': (c_int, [_P, _P, c_int, _P, c_int, _P, _P, _P, _P, c_double, c_int, c_int, c_int, _P, _P,
                                           c_int, _P]),
    'bcbf_model_set_oz_digits': (c_int, [c_void_p, c_int]),
    'bcbf_oz_debug_counters': (c_int, [c_int, POINTER(ctypes.c_ulonglong * 8)]),
    'bcbf_oz_gemm': (c_int, [c_int, c_int, c_int, c_double, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
    'bcbf_oz_gemm_tn': (c_int, [c_int, c_int, c_int, c_double, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
    'bcbf_oz_gemm_reserve': (c_int, [c_int, c_int, c_int]),
    'bcbf_oz_update': (c_int, [c_int, c_int, c_int, c_double, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
    'bcbf_oz_update_reserve': (c_int, [c_int, c_int, c_int]),
    'bcbf_set_trtri_i8': (c_int, [c_int]),
    'bcbf_set_potrf_i8': (c_int, [c_int]),
    'bcbf_set_potf2_variant': (c_int, [c_int]),
    'bcbf_set_gemm_tile_policy': (c_int, [c_int]),
    'bcbf_oz_set_cluster': (c_int, [c_int]),
    'bcbf_oz_set_group': (c_int, [c_int]),
    'bcbf_oz_debug_skip_loads': (c_int, [c_int]),
    'bcbf_oz_profile_enable': (c_int, [c_int]),
    'bcbf_oz_profile_read': (c_int, [POINTER(c_double), POINTER(c_int)]),
    'bcbf_model_set_var_path': (c_int, [c_void_p, c_int]),
    'bcbf_model_get_var_path': (c_int, [c_void_p]),
    'bcbf_model_oz_split_ms': (c_double, [c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load libbcbf.so (built by `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "bayesian_cbf_b200: native library %s is missing. Build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == BCBF_OK:
        return
    msg = load().bcbf_last_error().decode('utf-8', 'replace')
    if rc == BCBF_ERR_NOT_PD:
        raise NotPositiveDefiniteError(rc, "linalg.cholesky: " + msg)
    raise BcbfError(rc, "libbcbf error %d: %s" % (rc, msg))
