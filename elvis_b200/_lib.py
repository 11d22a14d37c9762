"""ctypes binding of libelvis_b200.so (the C ABI declared in include/elvis_b200.h).

There is no CPU fallback: importing this module without the built library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libelvis_b200.so")

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_SHAPE = 0, -1, -2, -3, -4
F32, F64 = 0, 1
REMOVE_HIGH, REMOVE_LOW = 0, 1
LEVELS_ROUND, LEVELS_INVERTED_ROUND, LEVELS_INVERTED_BINS = 0, 1, 2
ABI_VERSION = 7


class Plane(C.Structure):
    """struct elvis_plane"""
    _fields_ = [("data", C.c_void_p), ("frame_stride", C.c_int64), ("row_stride", C.c_int64),
                ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32), ("reserved", C.c_int32)]


class ElvisError(RuntimeError):
    def __init__(self, code: int, fn: str, detail: str = ""):
        self.code = code
        super().__init__(f"{fn} failed: {detail} (code {code})")


_vp, _i32, _i64, _f64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float
_PP = C.POINTER(Plane)

# name -> argtypes; every entry point returns int.  Kept in the order of the header.
SIGNATURES = {
    "elvis_score_sc_tc": [_PP, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp],
    "elvis_minmax": [_vp, _i32, _i64, _vp, _vp],
    "elvis_combine_removability": [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _f64, _f64,
                                   _i32, _vp, _vp, _vp],
    "elvis_normalize": [_vp, _i64, _vp, _vp],
    "elvis_importance_scores": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp],
    "elvis_select_rows": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp],
    "elvis_split_channels3": [_vp, _vp, _i32, _vp],
    "elvis_merge_channels3": [_vp, _vp, _i32, _vp],
    "elvis_normalize_select_rows": [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp],
    "elvis_shrink": [_PP, _PP, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_stretch": [_PP, _PP, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_shrink_yuv420": [_PP, _PP, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_stretch_yuv420": [_PP, _PP, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_levels_from_scores": [_vp, _i64, _i32, _i32, _vp, _vp],
    "elvis_degrade_blur": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _vp],
    "elvis_degrade_downsample": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp],
    "elvis_degrade_downsample_pow2_yuv420": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_dct_dampen": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _vp],
    "elvis_restore_unsharp": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp],
    "elvis_restore_lanczos": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "elvis_temporal_blend": [_PP, _i32, _f64, _vp],
    "elvis_pack_mask_bits": [_vp, _i64, _vp, _vp],
    "elvis_unpack_mask_bits": [_vp, _i64, _vp, _vp],
    "elvis_pack_levels_2bit": [_vp, _i64, _i32, _vp, _vp],
    "elvis_unpack_levels_2bit": [_vp, _i64, _i32, _vp, _vp],
    "elvis_rowcol_plan": [_vp, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "elvis_rowcol_expand": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp],
    "elvis_invert_block_map": [_vp, _i32, _i64, _vp, _i64, _vp],
    "elvis_gather_blocks": [_PP, _PP, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp],
    "elvis_roi_kvazaar": [_vp, _i64, _i32, _i32, _vp, _vp],
    "elvis_roi_prepare_f32": [_vp, _i64, _i32, _vp, _vp],
    "elvis_resize_area_f32": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "elvis_roi_svtav1_offsets": [_vp, _i64, _i32, _i32, _vp, _vp],
    "elvis_rgb_to_i420": [_PP, _PP, _PP, _PP, _i32, _vp],
    "elvis_area_downscale": [_PP, _PP, _i32, _i32, _vp],
    "elvis_merge_blocks": [_PP, _PP, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "elvis_levels_to_gray": [_vp, _i64, _i32, _i32, _vp, _vp],
    "elvis_gray_to_levels": [_vp, _i64, _f32, _f32, _vp, _vp],
    "elvis_resize_nearest": [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp],
    "elvis_resize_linear_float": [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "elvis_rgb_to_gray": [_PP, _PP, _i32, _vp],
    "elvis_refill_map": [_vp, _i32, _i64, _i64, _vp, _vp],
    "elvis_peer_alloc": [_i64, C.POINTER(_vp), _vp],
    "elvis_peer_open": [_vp, C.POINTER(_vp)],
    "elvis_peer_close": [_vp],
    "elvis_peer_free": [_vp],
    "elvis_peer_put": [_vp, _vp, _i64, _vp, C.c_uint32, _vp],
    "elvis_peer_signal": [_vp, C.c_uint32, _vp],
    "elvis_peer_wait": [_vp, C.c_uint32, _vp, _vp],
    "elvis_peer_allreduce_minmax": [_vp, _i32, _i32, _i32, _i32, C.POINTER(_vp), _i32, C.c_uint32, _vp, _vp],
}
EXPORTS = ["elvis_abi_version", "elvis_error_string", "elvis_last_cuda_error", "elvis_peer_mailbox_bytes", *SIGNATURES]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. elvis_b200 has no CPU fallback: build the CUDA library with "
            "`python -m elvis_b200.build` (needs nvcc; cross-compiles for sm_100a without a GPU).")
    lib = C.CDLL(LIB_PATH)
    lib.elvis_abi_version.restype = C.c_int
    lib.elvis_error_string.restype = C.c_char_p
    lib.elvis_error_string.argtypes = [C.c_int]
    lib.elvis_last_cuda_error.restype = C.c_int
    lib.elvis_peer_mailbox_bytes.restype = C.c_int64
    if lib.elvis_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.elvis_abi_version()} != {ABI_VERSION}; rebuild")
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    return lib


lib = _load()


def call(name: str, *args) -> None:
    """Invoke an entry point and raise on a non-zero return code (ValueError for the
    reference's own dimension check, elvis.py:1376)."""
    rc = getattr(lib, name)(*args)
    if rc != OK:
        raise_for(name, rc)


def raise_for(name: str, rc: int) -> None:
    detail = lib.elvis_error_string(rc).decode()
    if rc == ERR_CUDA:
        detail += f" (cudaError {lib.elvis_last_cuda_error()})"
    if rc == ERR_SHAPE:
        raise ValueError("Image dimensions must be divisible by block_size.")
    raise ElvisError(rc, name, detail)
