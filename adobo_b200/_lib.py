"""ctypes binding of libadobo_b200.so (include/adobo_b200.h).

Same FFI idiom as the reference's only native binding (adobo/preproc.py:446-459: ctypes +
numpy.ctypeslib.ndpointer, caller-allocated outputs).  The library is built in-tree by
`python -m adobo_b200.build`; importing this module without it raises -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np
from numpy.ctypeslib import ndpointer

HERE = os.path.dirname(os.path.abspath(__file__))
# ADOBO_B200_LIB_SUFFIX selects a diagnostic build of the same library (see build.py); the product is the default
LIB_PATH = os.path.join(HERE, 'libadobo_b200%s.so' % os.environ.get('ADOBO_B200_LIB_SUFFIX', ''))


class AdoboB200Error(Exception):
    """Raised for every non-zero return code of the C-ABI; carries adobo_last_error()."""

    def __init__(self, code, message):
        super().__init__('%s (adobo_b200 error %d)' % (message, code))
        self.code = code
        self.message = message


def _arr(dtype, ndim=1):
    return ndpointer(dtype=dtype, ndim=ndim, flags='C_CONTIGUOUS')


class _OptArr:
    """ndpointer that also accepts None (NULL)."""

    def __init__(self, dtype, ndim=1):
        self._base = _arr(dtype, ndim)

    def from_param(self, obj):
        if obj is None:
            return C.c_void_p(None)
        return self._base.from_param(obj)


_i32, _i64, _f64, _f32 = np.int32, np.int64, np.float64, np.float32
ctx_p = C.c_void_p

# name -> (restype, argtypes); every symbol include/adobo_b200.h declares
SIGNATURES = {
    'adobo_version': (C.c_int, []),
    'adobo_last_error': (C.c_char_p, []),
    'adobo_ctx_create': (C.c_int, [C.c_int, C.POINTER(ctx_p)]),
    'adobo_ctx_destroy': (None, [ctx_p]),
    'adobo_sync': (C.c_int, [ctx_p]),
    'adobo_comm_unique_id': (C.c_int, [_arr(np.uint8)]),
    'adobo_comm_init': (C.c_int, [ctx_p, C.c_int, C.c_int, _arr(np.uint8)]),
    'adobo_event_record': (C.c_int, [ctx_p, C.c_int]),
    'adobo_event_elapsed_ms': (C.c_int, [ctx_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    'adobo_profile': (C.c_int, [ctx_p, C.c_int]),
    'adobo_profile_read': (C.c_int, [ctx_p, C.c_int, C.c_char_p, _arr(_f32), _arr(_i64), _arr(_f64),
                                     C.POINTER(C.c_int)]),
    'adobo_host_register': (C.c_int, [C.c_void_p, C.c_size_t]),
    'adobo_host_unregister': (C.c_int, [C.c_void_p]),
    'adobo_launch_count': (C.c_int64, [ctx_p]),
    'adobo_l2_flush': (C.c_int, [ctx_p]),
    'adobo_upload_counts': (C.c_int, [ctx_p, C.c_int64, C.c_int32, _arr(_i64), _arr(_i32), _arr(_i32)]),
    'adobo_upload_normalized': (C.c_int, [ctx_p, C.c_int64, C.c_int32, _arr(_i64), _arr(_i32), _arr(_f64)]),
    'adobo_synth_counts': (C.c_int, [ctx_p, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, _arr(_f32, 2), C.c_int32,
                                     C.POINTER(C.c_int64)]),
    'adobo_download_counts': (C.c_int, [ctx_p, _OptArr(_i64), _OptArr(_i32), _OptArr(_i32)]),
    'adobo_count_stats': (C.c_int, [ctx_p, _OptArr(_i64), _OptArr(_i64), _OptArr(_i64)]),
    'adobo_normalize': (C.c_int, [ctx_p, C.c_double, C.c_int, C.c_double, _OptArr(_i64), _OptArr(_f64)]),
    'adobo_gene_moments': (C.c_int, [ctx_p, _OptArr(_f64), _OptArr(_f64)]),
    'adobo_hvg_seurat': (C.c_int, [ctx_p, C.c_int32, C.c_int32, _OptArr(_i32), C.POINTER(C.c_int32), _OptArr(_f64)]),
    'adobo_irlb': (C.c_int, [ctx_p, _arr(_i32), C.c_int32, C.c_int32, C.c_double, C.c_int32, _arr(_f64), C.c_int32,
                             C.c_int32, _OptArr(_f64, 2), _OptArr(_f64), _OptArr(_f64, 2), C.POINTER(C.c_int32),
                             C.POINTER(C.c_int32)]),
    'adobo_gather_subset': (C.c_int, [ctx_p, _arr(_i32), C.c_int32, C.c_int32]),
    'adobo_irlb_permuted': (C.c_int, [ctx_p, _OptArr(_i32), C.c_int32, C.c_uint64, C.c_int32, C.c_double, C.c_int32, _arr(_f64),
                                      C.c_int32, _OptArr(_f64, 2), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'adobo_perm_row': (C.c_int64, [C.c_int64, C.c_uint64, C.c_int32, C.c_int64]),
    'adobo_lanczos_csr': (C.c_int, [ctx_p, C.c_int64, C.c_int32, _arr(_i64), _arr(_i32), _arr(_f64), C.c_int32,
                                    C.c_double, C.c_int32, _arr(_f64), _OptArr(_f64, 2), _OptArr(_f64),
                                    _OptArr(_f64, 2), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'adobo_knn': (C.c_int, [ctx_p, _OptArr(_f64, 2), C.c_int64, C.c_int32, C.c_int32, _OptArr(_i64, 2)]),
    'adobo_knn_metric': (C.c_int, [ctx_p, _OptArr(_f64, 2), C.c_int64, C.c_int32, C.c_int32, C.c_int32, _OptArr(_i64, 2)]),
    'adobo_upload_embedding': (C.c_int, [ctx_p, _arr(_f64, 2), C.c_int64, C.c_int32]),
    'adobo_snn': (C.c_int, [ctx_p, _OptArr(_i64, 2), C.c_int64, C.c_int32, C.c_double, C.POINTER(C.c_int64),
                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'adobo_snn_fetch': (C.c_int, [ctx_p, _arr(_i32), _arr(_i32)]),
    'adobo_snn_fetch_counts': (C.c_int, [ctx_p, _arr(_i32)]),
    'adobo_result_sizes': (C.c_int, [ctx_p, _arr(_i64)]),
    'adobo_embed_path': (C.c_int, [ctx_p, C.c_double, C.c_int32, C.c_int32, _arr(_f64), C.c_int32, C.c_double,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    'adobo_fetch_pca': (C.c_int, [ctx_p, _OptArr(_f64, 2), _OptArr(_f64), _OptArr(_f64, 2), _OptArr(_i32)]),
    'adobo_fetch_knn': (C.c_int, [ctx_p, _OptArr(_i64, 2)]),
}

_lib = None


def load():
    """Loads the shared library (once).  Raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('adobo_b200: %s is missing -- build it with `python -m adobo_b200.build` '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().adobo_last_error()
        raise AdoboB200Error(code, msg.decode('utf-8', 'replace') if msg else 'unknown error')
