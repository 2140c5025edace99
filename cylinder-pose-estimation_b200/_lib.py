"""ctypes binding of liblgx.so (include/lgx.h).  No fallback: if the library or a B200 is missing,
every call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGX_LIB") or os.path.join(HERE, "liblgx.so")    # LGX_LIB: A/B experiments only

LGX_OK = 0
LGX_FLAG_HOLES, LGX_FLAG_GENERIC_FILL, LGX_FLAG_COMP_OVERFLOW, LGX_FLAG_CENT_OVERFLOW = 1, 2, 4, 8
LGX_OPT_MIXED_FROM_COLS = 1
LGX_OPT_TIMING = 2
LGX_OPT_RIDGE_PROF = 3
LGX_OPT_RIDGE_WARPS = 4
LGX_OPT_RIDGE_SMS = 5
LGX_OPT_JOINTS_GLOBAL = 10
LGX_OPT_PACKED_MASKS = 11
LGX_OPT_SAUVOLA = 6
LGX_OPT_HOST_SPLIT_FIRST = 7
LGX_OPT_FLOAT_DIV = 8
LGX_OPT_FUSED = 9

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); exactly the prototypes of include/lgx.h
PROTOTYPES = {
    "lgx_version": (_i, []),
    "lgx_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "lgx_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "lgx_destroy": (_i, [_vp]),
    "lgx_set_option": (_i, [_vp, _i, _i]),
    "lgx_set_gauss_weights": (_i, [_vp, C.POINTER(C.c_double)]),
    "lgx_strerror": (C.c_char_p, [_i]),
    "lgx_last_cuda_error": (C.c_char_p, []),
    "lgx_frontend": (_i, [_vp, _vp, _i, _i, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "lgx_frontend_host": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "lgx_bgr2gray": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "lgx_undistort": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lgx_blur5": (_i, [_vp, _vp, _i, _i, _i, _i, _sz, _sz, _vp, _vp]),
    "lgx_ridge": (_i, [_vp, _vp, _i, _i, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lgx_ridge_sauvola": (_i, [_vp, _vp, _i, _i, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lgx_last_ridge_kernel": (C.c_char_p, [_vp]),
    "lgx_last_joints_kernel": (C.c_char_p, [_vp]),
    "lgx_sauvola": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lgx_extract_joints": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "lgx_contour_centroids": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "lgx_debug_contours": (_i, [_vp, _i, _vp, _i, C.POINTER(_i)]),
    "lgx_get_stats": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _i]),
    "lgx_get_ridge_prof": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    "lgx_debug_sqrt": (_i, [C.c_ulonglong, C.c_ulonglong, _i, _vp, C.POINTER(C.c_ulonglong)]),
    "lgx_debug_fused_prof": (_i, [C.POINTER(C.c_ulonglong), _i]),
    "lgx_plane_pitch": (_i, [_i]),
    "lgx_bits_pitch": (_i, [_i]),
    "lgx_render_noisy": (_i, [_vp, _i, _i, _i, _i, C.c_float, C.c_uint64, _i, _vp, _vp]),
}

_lib = None


class LgxError(RuntimeError):
    pass


def load():
    """dlopen liblgx.so and set prototypes.  Loading needs no GPU (the CPU test-suite checks symbols)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LgxError(f"{LIB_PATH} not built: run `python __graft_entry__.py build` "
                           "(lgx has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status, what="lgx"):
    if status != LGX_OK:
        lib = load()
        msg = lib.lgx_strerror(status).decode()
        if status == -2:
            msg += ": " + lib.lgx_last_cuda_error().decode()
        raise LgxError(f"{what} failed ({status}): {msg}")
