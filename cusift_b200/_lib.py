"""ctypes binding of libcusift_b200.so — the C ABI declared in include/cusift_b200.h.

There is no fallback: if the CUDA library has not been built (``make lib`` /
``__graft_entry__.build()``) importing :func:`lib` raises, and every entry point
needs a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["CSB_LIB_PATH"]) if os.environ.get("CSB_LIB_PATH") else PKG_DIR / "libcusift_b200.so"   # override: kernel-variant experiments
HEADER_PATH = PKG_DIR.parent / "include" / "cusift_b200.h"

MAX_OCTAVES = 8


class CsbParams(C.Structure):
    """csb_params (include/cusift_b200.h) == SiftData's parameter fields (cuSIFT.h:45-51)."""

    _fields_ = [
        ("num_octaves", C.c_int),
        ("init_blur", C.c_double),
        ("peak_thresh", C.c_float),
        ("edge_thresh", C.c_float),
        ("lowest_scale", C.c_float),
        ("subsampling", C.c_float),
        ("rootsift", C.c_int),
    ]


_vp = C.c_void_p
_i = C.c_int
_ull = C.c_ulonglong
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/cusift_b200.h declares
SIGNATURES = {
    "csb_version": (_i, []),
    "csb_sizeof_sift_point": (_i, []),
    "csb_ctx_create": (_i, [_i, _i, C.POINTER(_vp)]),
    "csb_ctx_destroy": (None, [_vp]),
    "csb_ctx_device": (_i, [_vp]),
    "csb_ctx_num_slots": (_i, [_vp]),
    "csb_last_error": (C.c_char_p, [_vp]),
    "csb_host_alloc": (_i, [C.POINTER(_vp), _ull]),
    "csb_host_free": (_i, [_vp]),
    "csb_device_alloc": (_i, [_vp, C.POINTER(_vp), _ull]),
    "csb_device_free": (_i, [_vp, _vp]),
    "csb_forget_image": (_i, [_vp, _vp]),
    "csb_memcpy_h2d": (_i, [_vp, _vp, _vp, _ull]),
    "csb_memcpy_d2h": (_i, [_vp, _vp, _vp, _ull]),
    "csb_upload_image": (_i, [_vp, _vp, _i, _vp, _i, _i]),
    "csb_download_image": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "csb_extract": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(CsbParams), _vp, _i, _vp, _ip]),
    "csb_extract_host": (_i, [_vp, _vp, _i, _i, C.POINTER(CsbParams), _vp, _i, _vp, _ip]),
    "csb_extract_batch": (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i, _i, C.POINTER(CsbParams), C.POINTER(_vp),
                                C.POINTER(_vp), _i, _ip]),
    "csb_extract_batch_compact": (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i, _i, C.POINTER(CsbParams), C.POINTER(_vp),
                                        C.POINTER(_vp), _i, _ip]),
    "csb_rigid_transform": (_i, [_vp, _fp, _i, _i, _ip, _i, C.c_float, C.c_uint, _fp, _ip, C.c_char_p]),
    "csb_rigid_sample_hash": (C.c_uint, [C.c_uint, C.c_uint, C.c_uint, C.c_uint]),
    "csb_ingest_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i]),
    "csb_extract_batch_u8": (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i, _i, C.POINTER(CsbParams), C.POINTER(_vp),
                                   C.POINTER(_vp), _i, _ip]),
    "csb_scale_down": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "csb_scale_down_var": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, C.c_float]),
    "csb_rootsift": (_i, [_vp, _vp, _i]),
    "csb_match": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "csb_match_redo_blocks": (C.c_longlong, [_vp]),
    "csb_match_domain_fallbacks": (C.c_longlong, [_vp]),
    "csb_find_homography": (_i, [_vp, _vp, _i, _ip, _i, C.c_float, _fp, _ip]),
    "csb_allpairs_match_ransac": (_i, [_vp, _i, C.POINTER(_vp), _ip, _i, _ip, _ip, C.POINTER(C.c_uint), _i, _i, C.c_float,
                                       C.c_float, C.c_float, C.c_uint, _fp, _ip, _ip]),
    "csb_sample_hash": (C.c_uint, [C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]),
    "csb_allpairs_match_ransac_improve": (_i, [_vp, _i, C.POINTER(_vp), _ip, _i, _ip, _ip, C.POINTER(C.c_uint), _i, _i, C.c_float,
                                               C.c_float, C.c_float, C.c_uint, _i, C.c_float, _fp, _ip, _ip, _fp, _ip]),
    "csb_nccl_unique_id": (_i, [_vp]),
    "csb_nccl_comm_create": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp)]),
    "csb_nccl_comm_destroy": (_i, [_vp]),
    "csb_allpairs_distributed": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_vp), _ip, _i, _i, _i, C.c_float, C.c_float, C.c_float,
                                      C.c_uint, _i, C.c_float, _fp, _ip, _ip, _fp, _ip, C.POINTER(C.c_double)]),
    "csb_improve_homography": (_i, [_vp, _vp, _i, _fp, _i, C.c_float, C.c_float, C.c_float, _ip, _vp]),
    "csb_debug_octave": (_i, [_vp, _i, _fp, _fp, _ip, _ip]),
    "csb_profile_enable": (_i, [_vp, _i]),
    "csb_profile_reset": (_i, [_vp]),
    "csb_profile_count": (_i, [_vp]),
    "csb_profile_get": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "csb_launch_count": (C.c_longlong, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library (no CUDA call is made by loading it)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the sm_100a library first (make lib). "
                "cusift_b200 has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # raises AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.csb_sizeof_sift_point() != 588:
            raise RuntimeError("SiftPoint layout mismatch")
        _lib = L
    return _lib
