"""ctypes binding of libvstab.so (include/vstab.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present
when a context is created, this raises.  Nothing under oracle/ is ever imported here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libvstab.so")

VS_MEM_HOST, VS_MEM_DEVICE = 0, 1
VS_WARP_CV_EXACT_BILINEAR, VS_WARP_FLOAT_BILINEAR, VS_WARP_LANCZOS2 = 0, 1, 2
VS_BORDER_CONSTANT0, VS_BORDER_REPEAT_EDGE = 0, 1
VS_CLIP_DEBUG_TAPS = 1
VS_CLIP_NV12 = 2
VS_KERNEL_COUNT = 12


class VsImg(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("stride", C.c_int64), ("batch", C.c_int32), ("batch_stride", C.c_int64)]


class VsAlignParams(C.Structure):
    _fields_ = [("phase_correlate", C.c_int32), ("phase_correlate_threshold", C.c_double),
                ("threshold", C.c_double), ("smallest_fraction", C.c_float), ("max_iters", C.c_int32),
                ("pyramid_min_width", C.c_int32), ("pyramid_min_height", C.c_int32),
                ("max_displacement", C.c_double)]


class VsSweepParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("max_displacement", C.c_double), ("smallest_fraction", C.c_float),
                ("max_iters", C.c_int32), ("phase_correlate", C.c_int32)]


class VsPair(C.Structure):
    _fields_ = [("template_slot", C.c_int32), ("keyframe_slot", C.c_int32), ("invert", C.c_int32)]


# every symbol include/vstab.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_IMG = C.POINTER(VsImg)
SYMBOLS = {
    "vs_abi_version": (C.c_int, []),
    "vs_device_count": (C.c_int, []),
    "vs_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vs_ctx_destroy": (C.c_int, [_P]),
    "vs_ctx_set_stream": (C.c_int, [_P, _P]),
    "vs_ctx_synchronize": (C.c_int, [_P]),
    "vs_last_error": (C.c_char_p, [_P]),
    "vs_ctx_launch_count": (C.c_int64, [_P]),
    "vs_ctx_profile_enable": (C.c_int, [_P, C.c_int]),
    "vs_ctx_profile_reset": (C.c_int, [_P]),
    "vs_ctx_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "vs_ctx_profile_timeline": (C.c_int, [_P, _P, C.c_int]),
    "vs_kernel_name": (C.c_char_p, [C.c_int]),
    "vs_dev_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "vs_dev_free": (C.c_int, [_P, _P]),
    "vs_host_alloc_pinned": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "vs_host_free_pinned": (C.c_int, [_P, _P]),
    "vs_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "vs_pinned_free": (C.c_int, [_P]),
    "vs_host_is_pinned": (C.c_int, [_P]),
    "vs_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "vs_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "vs_bgr2gray_u8": (C.c_int, [_P, _IMG, _IMG, C.c_int]),
    "vs_ingest_bgr_u8": (C.c_int, [_P, _IMG, _IMG, _IMG, C.c_int]),
    "vs_pyr_down_u8": (C.c_int, [_P, _IMG, _IMG, C.c_int]),
    "vs_grad_xy_u8_f32": (C.c_int, [_P, _IMG, _IMG, _IMG, C.c_int]),
    "vs_grad_argmax_tile_size": (C.c_int, [C.c_int, C.c_int]),
    "vs_grad_argmax_f32_u16": (C.c_int, [_P, _IMG, _IMG, C.c_int, _P, _P, C.c_int]),
    "vs_sparse_jac_f32": (C.c_int, [_P, _IMG, _IMG, _P, _P, C.c_int, C.c_int, _P, _P, C.c_int]),
    "vs_sparse_warpdiff_u8_u16": (C.c_int, [_P, _IMG, _IMG, _P, C.c_int, C.c_int,
                                            C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int]),
    "vs_sparse_ica_f64": (C.c_int, [_P, _IMG, _IMG, _P, C.c_int, _P, C.c_int, _P, _P,
                                    C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int]),
    "vs_image_warp_u8_f32": (C.c_int, [_P, _IMG, _P, _IMG, C.c_int]),
    "vs_bgr_warp_u8": (C.c_int, [_P, _IMG, _P, _IMG, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vs_plane_warp_u8": (C.c_int, [_P, _IMG, C.c_int, _P, _IMG, C.c_int, C.c_int, C.c_int]),
    "vs_align_params_default": (None, [C.POINTER(VsAlignParams)]),
    "vs_clip_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(VsAlignParams), C.c_int, C.POINTER(_P)]),
    "vs_clip_destroy": (C.c_int, [_P]),
    "vs_clip_set_params": (C.c_int, [_P, C.POINTER(VsAlignParams)]),
    "vs_clip_levels": (C.c_int, [_P]),
    "vs_clip_level_info": (C.c_int, [_P, C.c_int] + [C.POINTER(C.c_int)] * 5),
    "vs_clip_upload": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64, C.c_int64, C.c_int]),
    "vs_clip_build_pyramids": (C.c_int, [_P, C.c_int, C.c_int]),
    "vs_clip_build_keyframes": (C.c_int, [_P, _P, C.c_int]),
    "vs_clip_align": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, C.c_int]),
    "vs_clip_align_sweep": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int]),
    "vs_clip_align_async": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "vs_clip_align_wait": (C.c_int, [_P, C.c_int, _P, _P]),
    "vs_clip_warp": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int64, C.c_int]),
    "vs_clip_upload_async": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64, C.c_int64]),
    "vs_clip_wait_uploads": (C.c_int, [_P]),
    "vs_clip_warp_to_host_async": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int64]),
    "vs_clip_sync_transfers": (C.c_int, [_P]),
    "vs_clip_get_bgr": (C.c_int, [_P, C.c_int, _P]),
    "vs_clip_get_gray": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vs_clip_get_keypoints": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "vs_clip_get_jacobians": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "vs_clip_get_warpdiff": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "vs_clip_get_solver_cycles": (C.c_int, [_P, C.c_int, _P]),
    "vs_clip_get_solver_level_cycles": (C.c_int, [_P, C.c_int, _P]),
    "vs_clip_get_phase": (C.c_int, [_P, C.c_int, _P]),
    "vs_phase_correlate_u8": (C.c_int, [_P, _IMG, _IMG, _P, C.c_int]),
    "vs_debug_invert4": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "vs_clip_get_selected": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
}

_lib = None


def load():
    """dlopen libvstab.so and bind every symbol of the header.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libvstab.so is not built (%s). Run `python -m video_stabilizer_b200.build`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class VsError(RuntimeError):
    pass


def check(ctx_handle, code: int, what: str) -> None:
    if code != 0:
        msg = load().vs_last_error(ctx_handle)
        raise VsError("%s failed (%d): %s" % (what, code, (msg or b"").decode()))


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def img_of(a: np.ndarray, width: int | None = None) -> VsImg:
    """Descriptor of a host numpy image: (h,w) gray / f32, (h,w,3) BGR, or a leading batch axis."""
    if a.ndim == 2:
        h, w = a.shape
        return VsImg(a.ctypes.data, w, h, a.strides[0] // a.itemsize, 1, 0)
    if a.ndim == 3 and a.shape[2] == 3 and a.dtype == np.uint8 and width is None:
        h, w, _ = a.shape
        return VsImg(a.ctypes.data, w, h, a.strides[0], 1, 0)
    raise ValueError("unsupported image array shape %r" % (a.shape,))


def img_batch(a: np.ndarray, bgr: bool = False) -> VsImg:
    if bgr:
        n, h, w, _ = a.shape
        return VsImg(a.ctypes.data, w, h, a.strides[1], n, a.strides[0])
    n, h, w = a.shape
    return VsImg(a.ctypes.data, w, h, a.strides[1] // a.itemsize, n, a.strides[0] // a.itemsize)
