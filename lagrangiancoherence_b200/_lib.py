"""ctypes binding of liblcs_b200.so (the C ABI declared in include/lcs_b200.h).

There is no CPU fallback: if the shared object is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LCS_B200_LIB') or os.path.join(HERE, 'liblcs_b200.so')   # override: tuning experiments only

LCS_F64, LCS_F32 = 0, 1
LCS_X_CYCLIC, LCS_X_CLAMP_POINTWISE, LCS_X_CLAMP_OUTER = 0, 1, 2
LCS_LAYOUT_PAIR4, LCS_LAYOUT_ES = 0, 1
LCS_ARITH_F64, LCS_ARITH_F32 = 0, 1
LCS_HALO_LO, LCS_HALO_HI = 2, 3          # include/lcs_b200.h
ABI_VERSION = 3

c_void_p, c_int, c_double, c_size_t, c_int64 = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_int64


class Grid(C.Structure):
    _fields_ = [('nlat', C.c_int32), ('nlon', C.c_int32),
                ('lat_min', c_double), ('lat_max', c_double), ('lon_min', c_double), ('lon_max', c_double)]


class Particles(C.Structure):
    _fields_ = [('nrow', C.c_int32), ('ncol', C.c_int32), ('row0', C.c_int32), ('nrow_global', C.c_int32),
                ('lat', c_void_p), ('lon', c_void_p), ('kx', c_void_p), ('hx', c_void_p),
                ('ky', c_double), ('hy', c_double)]


class XRank(C.Structure):
    _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('ngroups', C.c_int32), ('reserved', C.c_int32),
                ('mailboxes', c_void_p), ('mailbox_bytes', c_size_t)]


class AdvectOpts(C.Structure):
    _fields_ = [('nsteps', C.c_int32), ('settls_order', C.c_int32), ('interp_order', C.c_int32),
                ('xmode', C.c_int32), ('strict', C.c_int32),
                ('nwindows', C.c_int32), ('level0', C.c_int32), ('level_stride', C.c_int32), ('arith', C.c_int32),
                ('round32', C.c_int32), ('xrank', C.POINTER(XRank))]


class Winds(C.Structure):
    _fields_ = [('layout', C.c_int32), ('dtype', C.c_int32),
                ('raw_a', c_void_p), ('raw_b', c_void_p), ('coef_a', c_void_p), ('coef_b', c_void_p),
                ('raw_planar', C.c_int32), ('raw_dtype', C.c_int32)]


# name -> (restype, argtypes): every symbol include/lcs_b200.h declares
SIGNATURES = {
    'lcs_abi_version': (c_int, []),
    'lcs_last_error': (C.c_char_p, []),
    'lcs_kernel_launches': (C.c_ulonglong, []),
    'lcs_prefilter_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'lcs_prefilter': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                              c_int, c_int, c_int, c_int, c_void_p]),
    'lcs_pack_pairs': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'lcs_time_lerp': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'lcs_regrid_linear_nearest': (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 10 + [c_int, c_int, c_void_p, c_void_p]),
    'lcs_spectral_truncate_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'lcs_spectral_truncate': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p, c_void_p]),
    'lcs_gaussian_filter2d': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'lcs_advect_workspace_bytes': (c_size_t, [C.POINTER(Particles), C.POINTER(AdvectOpts)]),
    'lcs_advect_check': (c_int, [c_void_p, c_void_p]),
    'lcs_xrank_mailbox_bytes': (c_size_t, [c_int, c_int, c_int]),
    'lcs_pack_es': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'lcs_advect': (c_int, [C.POINTER(Grid), C.POINTER(Particles), C.POINTER(AdvectOpts), C.POINTER(Winds),
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'lcs_ftle_epilogue': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_double, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'lcs_map_coordinates': (c_int, [C.POINTER(Grid), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_int, c_void_p, c_void_p]),
    'lcs_fourth_order_derivative': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'lcs_spectral_norm_3x3': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'lcs_ridge_classify': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'lcs_gather_peak_smem': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'lcs_gather_peak': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int,
                                c_void_p, c_void_p]),
}

_lib = None


class LcsError(RuntimeError):
    pass


def load():
    """Load the library once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LcsError(
            f'{LIB_PATH} is missing: build it with `python -m lagrangiancoherence_b200.build` '
            '(nvcc, sm_100a). lagrangiancoherence_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lcs_abi_version() != ABI_VERSION:
        raise LcsError(f'liblcs_b200.so ABI {lib.lcs_abi_version()} != binding {ABI_VERSION}: rebuild')
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().lcs_last_error().decode(errors='replace')
        raise LcsError(f'{what} failed with status {status}: {msg}')
