"""ctypes binding of ``libsdg.so`` (C ABI declared in ``include/sdg.h``).

This is the whole FFI surface: plain pointers and sizes, no torch types cross the boundary.  The
library is hand-written CUDA for sm_100a; there is no CPU or PyTorch fallback -- if the shared object
is missing or the device is not a B200 the import / first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdg.so")

# constants mirrored from include/sdg.h
ABI_VERSION = 1
ARCH_DCGAN32, ARCH_STYLEGAN2, ARCH_SNGAN32, ARCH_SNGAN64 = 1, 2, 32, 64
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
LAYOUT_U8_NHWC, LAYOUT_F32_NCHW = 0, 1
RANGE_ACT, RANGE_WEIGHT = 1, 2

_vp, _i, _i64, _d, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float, C.c_size_t
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); one entry per SDG_API declaration in include/sdg.h
SIGNATURES = {
    "sdg_last_error": (C.c_char_p, []),
    "sdg_abi_version": (_i, []),
    "sdg_ctx_create": (_i, [_i, _pp]),
    "sdg_ctx_destroy": (_i, [_vp]),
    "sdg_ctx_set_range_flag": (_i, [_vp, _vp]),
    "sdg_ctx_set_chunk": (_i, [_vp, _i64]),
    "sdg_sngan_load": (_i, [_vp, _i, _i, _pp, _pp, _pp, _i, _i, _vp]),
    "sdg_sngan_sigmas": (_i, [_vp, _vp, _vp]),
    "sdg_dcgan_load": (_i, [_vp, _pp, _pp, _pp, _pp, _pp, _vp, _vp, _i, _vp]),
    "sdg_stylegan2_load": (_i, [_vp, _i, _i, _pp, _i, _vp]),
    "sdg_ctx_set_batch": (_i, [_vp, _i]),
    "sdg_d_forward": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "sdg_conv2d_h16": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i,
                             _vp]),
    "sdg_conv2d_sg2_h16": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _f, _vp, _vp, _i,
                                 _vp]),
    "sdg_blur_h16": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "sdg_set_conv_pair": (_i, [_i]),
    "sdg_first_conv_h16": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "sdg_sngan32_block1_fused_h16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "sdg_sngan64_block1_fused_h16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "sdg_resize_center_crop_u8": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "sdg_stats_update": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "sdg_window_moments_f32": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sdg_window_moments_f64": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sdg_score_floor_min": (_i, [_vp, _vp, _i64, C.POINTER(_d), _i, _d, _d, _vp, _vp, _vp]),
    "sdg_score_clip": (_i, [_vp, _i64, _i, _vp, _d, _d, _vp]),
    "sdg_score_clip_gathered": (_i, [_vp, _i, _i64, _i64, _d, _d, _vp, _vp]),
    "sdg_topk_workspace_bytes": (_sz, [_i64]),
    "sdg_topk_indices": (_i, [_vp, _i64, _i, _i, _vp, _vp, _sz, _vp]),
    "sdg_drs_update_max": (_i, [_vp, _i, _vp, _vp]),
    "sdg_drs_accept": (_i, [_vp, _i, _vp, _f, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sdg_ctx_profile": (_i, [_vp, _i]),
    "sdg_ctx_profile_read": (_i, [_vp, C.POINTER(_d), C.POINTER(_i64), C.POINTER(_d)]),
    "sdg_launch_count": (_i64, []),
    "sdg_launch_count_reset": (None, []),
}

_lib = None


class SdgError(RuntimeError):
    pass


def load():
    """Load libsdg.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdgError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C self-diagnosing-gan_b200/csrc).  diagan_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.sdg_abi_version() != ABI_VERSION:
        raise SdgError(f"libsdg.so ABI {lib.sdg_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sdg_last_error()
        raise SdgError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or None)."""
    return C.c_void_p(None if t is None else t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def stream_ptr(device=None) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
