"""ctypes binding of libspalign_b200.so (the C ABI declared in include/spalign.h).

There is no CPU fallback: if the shared library is missing or a symbol does not resolve,
``load()`` raises.  Build it with ``python __graft_entry__.py`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SPALIGN_LIB', os.path.join(_HERE, 'libspalign_b200.so'))

I32, I64, U8 = 0, 1, 2
F32, F64 = 0, 1
F_LABEL_RANGE, F_NNZ_OVERFLOW, F_EMPTY_ROW = 1, 2, 4
KM_RUNNING, KM_CONVERGED, KM_EMPTY_CLUSTER, KM_ITER_CAP, KM_COMM_TIMEOUT = -1, 0, 1, 2, 3
COMM_HANDLE_BYTES = 64
ABI_VERSION = 1

_p, _i, _l, _d, _z = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/spalign.h
SIGNATURES = {
    'spalign_abi_version': (_i, []),
    'spalign_last_error': (C.c_char_p, []),
    'spalign_label_max': (_i, [_p, _i, _i, _i, _i, _p, _p]),
    'spalign_overlap_workspace_bytes': (_z, [_i, _i, _i, _i, _i, _l, _l]),
    'spalign_overlap_csr': (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _l, _p, _p, _l, _p, _p, _p, _p,
                                 _p, _p, _p, _p, _p, _z, _p]),
    'spalign_pool': (_i, [_p, _i, _i, _i, _i, _p, _l, _i, _p, _p, _p, _p, _p, _p, _i, _p, _l,
                          _p]),
    'spalign_pool_weighted': (_i, [_p, _i, _i, _i, _i, _p, _l, _i, _p, _p, _p, _p, _p, _p, _i, _p, _l,
                                   _p]),
    'spalign_overlap_bilinear_workspace_bytes': (_z, [_i, _i, _i, _i, _i, _l, _l]),
    'spalign_overlap_bilinear_csr': (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _l, _p, _p, _p, _p, _p, _p,
                                          _p, _p, _l, _p, _p, _p, _p, _p, _p, _z, _p]),
    'spalign_sample_anchors': (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _l, _p, _p, _p, _p, _i, C.c_uint64, _p,
                                    _p, _p]),
    'spalign_anchor_weights': (_i, [_p, _p, _l, _i, _i, _i, _i, _p, _p, _p, _p]),
    'spalign_nchw_to_cellmajor': (_i, [_p, _p, _i, _i, _i, _p]),
    'spalign_kmeans_groups_workspace_bytes': (_z, [_i, _i, _i]),
    'spalign_kmeans_groups': (_i, [_p, _i, _l, _i, _i, _l, _p, _i, _i, _i, _p, _i, _p, _p, _p,
                                   _p, _p, _z, _p]),
    'spalign_kmeans_sweep': (_i, [_p, _i, _l, _i, _i, _l, _l, _p, _i, _i, _p, _i, _p, _i, _p, _p,
                                  _p, _p]),
    'spalign_kmeans_iterate': (_i, [_p, _i, _l, _i, _i, _l, _l, _p, _i, _i, _p, _i, _p, _i, _i, _p, _p,
                                    _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'spalign_kmeans_iterate_dist': (_i, [_p, _i, _l, _i, _i, _l, _l, _p, _i, _i, _p, _i, _p, _i, _i, _p,
                                         _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'spalign_comm_create': (_i, [_i, _i, _l, C.POINTER(_p)]),
    'spalign_comm_handle': (_i, [_p, _p]),
    'spalign_comm_connect': (_i, [_p, _p]),
    'spalign_comm_destroy': (_i, [_p]),
    'spalign_kmeans_finish': (_i, [_p, _i, _l, _i, _i, _l, _l, _p, _i, _i, _p, _i, _i, _p, _p, _p, _p,
                                   _p, _p, _p, _p, _i, _i, _p]),
    'spalign_kmeans_reduce': (_i, [_p, _p, _i, _i, _i, _p, _p]),
    'spalign_kmeans_update': (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    'spalign_kmeans_debug_stats': (_i, [C.POINTER(C.c_int64), _i]),
    'spalign_kmeans_init': (_i, [_p, _p, _i, _p, _p, _p, _p, _p]),
    'spalign_slic_segments': (_i, [_i, _i, _i]),
    'spalign_slic_workspace_bytes': (_z, [_i, _i, _i, _i]),
    'spalign_slic': (_i, [_p, _i, _i, _i, _i, _d, _i, _i, _i, _d, _p, _p, _p, _z, _p]),
    'spalign_felzenszwalb_workspace_bytes': (_z, [_i, _i, _i]),
    'spalign_felzenszwalb': (_i, [_p, _i, _i, _i, _d, _d, _i, _p, _p, _p, _z, _p]),
    'spalign_paint': (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p]),
    'spalign_refine': (_i, [_p, _i, _l, _i, _i, _p, _p, _p, _p, _d, _p, _p, _p, _p]),
    'spalign_resize_nearest_u8': (_i, [_p, _i, _i, _i, _p, _i, _i, _p]),
    'spalign_confusion2': (_i, [_p, _p, _i, _l, _p, _p]),
}

_lib = None


class SpalignError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpalignError(
            'libspalign_b200.so not found at %s -- build it with `python __graft_entry__.py` '
            '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.spalign_abi_version() != ABI_VERSION:
        raise SpalignError('ABI version mismatch: library %d, binding %d'
                           % (lib.spalign_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().spalign_last_error().decode('utf-8', 'replace')
        raise SpalignError('%s failed (status %d): %s' % (what or 'spalign call', rc, msg))


def header_symbols():
    """Function names declared in include/spalign.h (used by the CPU tests)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), 'include', 'spalign.h')
    with open(hdr) as fp:
        text = fp.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(spalign_[a-z0-9_]+)\s*\(', text)))
