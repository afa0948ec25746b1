"""ctypes binding of libvstab_host.so (include/vstab_host.h): the C++ host layer that keeps
the reference's classes — VideoAligner, VideoStabilizer, L1SmootherCenter,
SimilarityTransform, the imgproc.hpp operators — on top of the CUDA C ABI.

Used by the tests and bench.py; C++ callers include the .hpp files directly.  The transform
algebra, smoother and trajectory are host-only C++ and work without a GPU; everything that
touches an image raises without one (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libvstab_host.so")


class VshStabParams(C.Structure):
    _fields_ = [("aligner", capi.VsAlignParams), ("lag", C.c_int32), ("smoother_memory", C.c_int32), ("lambda_", C.c_double),
                ("enable_smoother", C.c_int32), ("crop_pixels", C.c_int32), ("min_disp", C.c_double), ("max_disp", C.c_double),
                ("min_decay", C.c_double), ("max_decay", C.c_double)]


_P, _I, _D, _I64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
_PI = C.POINTER(C.c_int)
_AP = C.POINTER(capi.VsAlignParams)
_SP = C.POINTER(VshStabParams)
SYMBOLS = {
    "vsh_stab_params_default": (None, [_SP]),
    "vsh_last_error": (C.c_char_p, []),
    "vsh_tf_inverse": (None, [_P, _P]),
    "vsh_tf_compose": (None, [_P, _P, _P]),
    "vsh_tf_warp": (None, [_P, _D, _D, _P]),
    "vsh_tf_warp_center": (None, [_P, _D, _D, _D, _D, _P]),
    "vsh_tf_max_corner_displacement": (_D, [_P, _D, _D]),
    "vsh_PyrDown": (_I, [_P, _I, _I, _P, _I, _I]),
    "vsh_GradXY": (_I, [_P, _I, _I, _P, _P]),
    "vsh_GradArgMax": (_I, [_P, _P, _I, _I, _PI, _P, _P, _I]),
    "vsh_SparseJacobian": (_I, [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P]),
    "vsh_SparseWarpDiff": (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P]),
    "vsh_SparseICA": (_I, [_P, _P, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P]),
    "vsh_ImageWarp": (_I, [_P, _I, _I, _P, _P, _I, _I]),
    "vsh_warpBySimilarityTransform": (_I, [_P, _I, _I, _I64, _P, _P]),
    "vsh_smoother_create": (_P, [_I, _I, _D]),
    "vsh_smoother_destroy": (None, [_P]),
    "vsh_smoother_update": (_I, [_P, _P, _P]),
    "vsh_tvl1_relax": (None, [_P, _I, _D, _I, _P]),
    "vsh_trajectory_create": (_P, [_SP]),
    "vsh_trajectory_destroy": (None, [_P]),
    "vsh_trajectory_push": (_I, [_P, _P, _I, _I, _I, _P]),
    "vsh_aligner_create": (_P, [_I]),
    "vsh_aligner_destroy": (None, [_P]),
    "vsh_aligner_align": (_I, [_P, _P, _I, _I, _I64, _AP, _P]),
    "vsh_stabilizer_create": (_P, [_SP, _I]),
    "vsh_stabilizer_destroy": (None, [_P]),
    "vsh_stabilizer_process": (_I, [_P, _P, _I, _I, _I64, _P, _PI, _PI]),
    "vsh_clipstab_create": (_P, [_I, _I, _I, _I, _SP]),
    "vsh_clipstab_create_nv12": (_P, [_I, _I, _I, _I, _SP]),
    "vsh_clipstab_destroy": (None, [_P]),
    "vsh_clipstab_reset": (_I, [_P]),
    "vsh_clipstab_set_pipeline_frames": (_I, [_P, _I]),
    "vsh_clipstab_set_solver_lanes": (_I, [_P, _I]),
    "vsh_clipstab_feed": (_I, [_P, _P, _I, _I64, _I64, _I, _P, _I64, _I]),
    "vsh_clipstab_upload_only": (_I, [_P, _I64, _P, _I, _I64, _I64, _I]),
    "vsh_clipstab_feed_resident": (_I, [_P, _I, _P, _I64, _I]),
    "vsh_clipstab_last_records": (_I, [_P, _P, _P, _P]),
    "vsh_clipstab_out_size": (_I, [_P, _PI, _PI]),
    "vsh_clipstab_context": (_P, [_P]),
    "vsh_multigpu_create": (_P, [_P, _I, _I, _I, _I, _SP]),
    "vsh_multigpu_destroy": (None, [_P]),
    "vsh_multigpu_stabilize": (_I, [_P, _P, _I, _I64, _I64, _P, _I64, _P, _P]),
    "vsh_clipstab_clip": (_P, [_P]),
    "vsh_gridsearch_create": (_P, [_I, _I, _I, _I, _I, _I]),
    "vsh_gridsearch_destroy": (None, [_P]),
    "vsh_gridsearch_reference_grid": (_I, [_P, _I]),
    "vsh_flow_median_px": (_D, [_P, _I, _I]),
    "vsh_gridsearch_jitter": (_I, [_P, _P, _I, _I64, _I64, _P]),
    "vsh_gridsearch_run": (_I, [_P, _P, _I, _I64, _I64, _P, _I, _P, _P, _P, _P]),
    "vsh_parttraj_create": (_P, [_I, _I, _I, _I, _I64, _I, _I, _SP, C.c_char_p, _I]),
    "vsh_parttraj_destroy": (None, [_P]),
    "vsh_parttraj_output_count": (_I, [_P]),
    "vsh_parttraj_run": (_I, [_P, _P, _P, _P, _P]),
    "vsh_partstab_create": (_P, [_I, _I, _I, _I, _I, _I64, _I, _I, _SP, C.c_char_p, _I, _I, _I]),
    "vsh_partstab_create_nv12": (_P, [_I, _I, _I, _I, _I, _I64, _I, _I, _SP, C.c_char_p, _I, _I, _I]),
    "vsh_partstab_destroy": (None, [_P]),
    "vsh_partstab_local_count": (_I, [_P]),
    "vsh_partstab_local_frame": (_I64, [_P, _I]),
    "vsh_partstab_local_is_halo": (_I, [_P, _I]),
    "vsh_partstab_output_count": (_I, [_P]),
    "vsh_partstab_output_frame": (_I, [_P, _I]),
    "vsh_partstab_upload_resident": (_I, [_P, _P, _I64, _I64, _I]),
    "vsh_partstab_stabilize": (_I, [_P, _P, _I64, _I64, _P, _I64, _I]),
    "vsh_partstab_records": (_I64, [_P, _P, _P, _P]),
    "vsh_partstab_out_size": (_I, [_P, _PI, _PI]),
    "vsh_partstab_set_lanes": (_I, [_P, _I]),
    "vsh_partstab_context": (_P, [_P]),
}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    capi.load()   # libvstab.so first, so the host library's dependency resolves to the same image
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libvstab_host.so is not built (%s). Run `python -m video_stabilizer_b200.build`." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HostError(RuntimeError):
    pass


def _raise(what):
    raise HostError("%s: %s" % (what, (load().vsh_last_error() or b"").decode()))


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _t(T):
    return np.ascontiguousarray(T, np.float64)


def stab_params_default() -> VshStabParams:
    p = VshStabParams()
    load().vsh_stab_params_default(C.byref(p))
    return p


# ---- SimilarityTransform
def tf_inverse(T):
    a, out = _t(T), np.zeros(4)
    load().vsh_tf_inverse(_p(a), _p(out))
    return out


def tf_compose(T1, T2):
    a, b, out = _t(T1), _t(T2), np.zeros(4)
    load().vsh_tf_compose(_p(a), _p(b), _p(out))
    return out


def tf_warp(T, x, y, center=None):
    a, out = _t(T), np.zeros(2)
    if center is None:
        load().vsh_tf_warp(_p(a), float(x), float(y), _p(out))
    else:
        load().vsh_tf_warp_center(_p(a), float(x), float(y), float(center[0]), float(center[1]), _p(out))
    return out


def tf_max_corner_displacement(T, w, h):
    a = _t(T)
    return load().vsh_tf_max_corner_displacement(_p(a), float(w), float(h))


# ---- imgproc.hpp operators through the C++ wrappers (dense host arrays)
def PyrDown(img, ow=None, oh=None):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ow, oh = (w // 2 if ow is None else ow), (h // 2 if oh is None else oh)
    out = np.empty((oh, ow), np.uint8)
    r = load().vsh_PyrDown(_p(img), w, h, _p(out), ow, oh)
    if r < 0:
        _raise("PyrDown")
    return bool(r), out


def GradXY(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    gx, gy = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    r = load().vsh_GradXY(_p(img), w, h, _p(gx), _p(gy))
    if r < 0:
        _raise("GradXY")
    return bool(r), gx, gy


def GradArgMax(gx, gy):
    gx, gy = np.ascontiguousarray(gx, np.float32), np.ascontiguousarray(gy, np.float32)
    h, w = gx.shape
    cap = (w // 2) * (h // 2) * 2
    lmx, lmy = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16)
    tile = C.c_int()
    r = load().vsh_GradArgMax(_p(gx), _p(gy), w, h, C.byref(tile), _p(lmx), _p(lmy), cap)
    if r < 0:
        _raise("GradArgMax")
    t = tile.value
    tw, th = w // t, h // t
    return bool(r), t, lmx[:2 * tw * th].reshape(2, th, tw).copy(), lmy[:2 * tw * th].reshape(2, th, tw).copy()


def SparseJacobian(gx, gy, lmx, lmy):
    gx, gy = np.ascontiguousarray(gx, np.float32), np.ascontiguousarray(gy, np.float32)
    lmx, lmy = np.ascontiguousarray(lmx, np.uint16), np.ascontiguousarray(lmy, np.uint16)
    h, w = gx.shape
    _, th, tw = lmx.shape
    jx, jy = np.zeros((4, th, tw), np.float32), np.zeros((4, th, tw), np.float32)
    r = load().vsh_SparseJacobian(_p(gx), _p(gy), w, h, _p(lmx), _p(lmy), tw, th, _p(jx), _p(jy))
    if r < 0:
        _raise("SparseJacobian")
    return bool(r), jx, jy


def SparseWarpDiff(tmpl, key, lm, T):
    tmpl, key = np.ascontiguousarray(tmpl, np.uint8), np.ascontiguousarray(key, np.uint8)
    lm, T = np.ascontiguousarray(lm, np.uint16), _t(T)
    h, w = key.shape
    _, th, tw = lm.shape
    out = np.zeros((th, tw), np.uint16)
    r = load().vsh_SparseWarpDiff(_p(tmpl), _p(key), w, h, _p(lm), tw, th, _p(T), _p(out))
    if r < 0:
        _raise("SparseWarpDiff")
    return bool(r), out


def SparseICA(tmpl, key, selx, sely, jx, jy, T):
    tmpl, key = np.ascontiguousarray(tmpl, np.uint8), np.ascontiguousarray(key, np.uint8)
    selx, sely = np.ascontiguousarray(selx, np.uint16), np.ascontiguousarray(sely, np.uint16)
    jx, jy, T = np.ascontiguousarray(jx, np.float32), np.ascontiguousarray(jy, np.float32), _t(T)
    h, w = key.shape
    out = np.zeros(4)
    r = load().vsh_SparseICA(_p(tmpl), _p(key), w, h, _p(selx), selx.shape[1], _p(sely), sely.shape[1], _p(jx), _p(jy), _p(T), _p(out))
    if r < 0:
        _raise("SparseICA")
    return bool(r), out


def ImageWarp(img, T, ow=None, oh=None):
    img, T = np.ascontiguousarray(img, np.uint8), _t(T)
    h, w = img.shape
    ow, oh = (w if ow is None else ow), (h if oh is None else oh)
    out = np.empty((oh, ow), np.float32)
    r = load().vsh_ImageWarp(_p(img), w, h, _p(T), _p(out), ow, oh)
    if r < 0:
        _raise("ImageWarp")
    return bool(r), out


def warpBySimilarityTransform(bgr, T):
    bgr, T = np.ascontiguousarray(bgr, np.uint8), _t(T)
    h, w, _ = bgr.shape
    out = np.empty((h, w, 3), np.uint8)
    if load().vsh_warpBySimilarityTransform(_p(bgr), w, h, bgr.strides[0], _p(T), _p(out)) < 0:
        _raise("warpBySimilarityTransform")
    return out


# ---- host-only classes
def tvl1_relax(data, lam, iterations=100):
    d = np.ascontiguousarray(data, np.float64)
    out = np.zeros_like(d)
    load().vsh_tvl1_relax(_p(d), d.size, float(lam), iterations, _p(out))
    return out


class _Handle:
    _destroy = None

    def close(self):
        if getattr(self, "h", None):
            getattr(load(), self._destroy)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class L1SmootherCenter(_Handle):
    _destroy = "vsh_smoother_destroy"

    def __init__(self, lag_behind, lag_ahead, lam=1.0):
        self.h = C.c_void_p(load().vsh_smoother_create(lag_behind, lag_ahead, float(lam)))

    def update(self, meas):
        m, out = _t(meas), np.zeros(4)
        ok = load().vsh_smoother_update(self.h, _p(m), _p(out))
        return bool(ok), out


class StabilizerTrajectory(_Handle):
    _destroy = "vsh_trajectory_destroy"

    def __init__(self, params: VshStabParams | None = None):
        self.params = params or stab_params_default()
        self.h = C.c_void_p(load().vsh_trajectory_create(C.byref(self.params)))

    def push(self, meas, success, w, h):
        m, out = _t(meas), np.zeros(4)
        due = load().vsh_trajectory_push(self.h, _p(m), int(bool(success)), w, h, _p(out))
        return bool(due), out


# ---- GPU classes
class VideoAligner(_Handle):
    _destroy = "vsh_aligner_destroy"

    def __init__(self, device: int = -1):
        self.h = C.c_void_p(load().vsh_aligner_create(device))
        if not self.h:
            _raise("VideoAligner")

    def AlignNextFrame(self, frame, params: capi.VsAlignParams | None = None):
        frame = np.ascontiguousarray(frame, np.uint8)
        h, w, _ = frame.shape
        T = np.zeros(4)
        r = load().vsh_aligner_align(self.h, _p(frame), w, h, frame.strides[0], C.byref(params) if params is not None else None, _p(T))
        if r < 0:
            _raise("AlignNextFrame")
        return bool(r), T


class VideoStabilizer(_Handle):
    _destroy = "vsh_stabilizer_destroy"

    def __init__(self, params: VshStabParams | None = None, device: int = -1):
        self.params = params or stab_params_default()
        self.h = C.c_void_p(load().vsh_stabilizer_create(C.byref(self.params), device))
        if not self.h:
            _raise("VideoStabilizer")

    def processFrame(self, frame):
        frame = np.ascontiguousarray(frame, np.uint8)
        h, w, _ = frame.shape
        # the frame that comes back is `lag` frames old and may have another (earlier) size
        self._max_bytes = max(getattr(self, "_max_bytes", 0), h * w * 3)
        out = np.empty(self._max_bytes, np.uint8)
        ow, oh = C.c_int(), C.c_int()
        r = load().vsh_stabilizer_process(self.h, _p(frame), w, h, frame.strides[0], _p(out), C.byref(ow), C.byref(oh))
        if r < 0:
            _raise("processFrame")
        if r == 0:
            return None
        return out.ravel()[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()


class ClipStabilizer(_Handle):
    """Batched VideoStabilizer (clip_stabilizer.hpp)."""
    _destroy = "vsh_clipstab_destroy"

    def __init__(self, width, height, chunk_frames, params: VshStabParams | None = None, device: int = 0, nv12: bool = False):
        """nv12: frames in and out are NV12 ((h * 3 / 2, w) u8 arrays) instead of (h, w, 3) BGR."""
        self.params = params or stab_params_default()
        self.width, self.height, self.chunk, self.nv12 = width, height, chunk_frames, nv12
        create = load().vsh_clipstab_create_nv12 if nv12 else load().vsh_clipstab_create
        self.h = C.c_void_p(create(device, width, height, chunk_frames, C.byref(self.params)))
        if not self.h:
            _raise("ClipStabilizer")
        ow, oh = C.c_int(), C.c_int()
        load().vsh_clipstab_out_size(self.h, C.byref(ow), C.byref(oh))
        self.out_w, self.out_h = ow.value, oh.value
        self.out_frame_bytes = self.out_w * self.out_h * 3 // (2 if nv12 else 1)
        self.out_shape = (self.out_h * 3 // 2, self.out_w) if nv12 else (self.out_h, self.out_w, 3)
        self.ctx_handle = C.c_void_p(load().vsh_clipstab_context(self.h))

    def reset(self):
        if load().vsh_clipstab_reset(self.h) < 0:
            _raise("reset")

    def set_pipeline_frames(self, frames: int):
        load().vsh_clipstab_set_pipeline_frames(self.h, int(frames))

    def set_solver_lanes(self, lanes: int):
        load().vsh_clipstab_set_solver_lanes(self.h, int(lanes))

    def feed(self, frames: np.ndarray) -> np.ndarray:
        """frames: (n,h,w,3) u8 host array; returns the (k,oh,ow,3) stabilized frames that became due."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = frames.shape[0]
        out = np.empty((n,) + self.out_shape, np.uint8)
        k = load().vsh_clipstab_feed(self.h, _p(frames), n, frames.strides[1], frames.strides[0], capi.VS_MEM_HOST,
                                     _p(out), self.out_frame_bytes, capi.VS_MEM_HOST)
        if k < 0:
            _raise("feed")
        return out[:k]

    def feed_ptr(self, ptr: int, n: int, row_stride: int, frame_stride: int, mem: int, out_ptr: int, out_mem: int) -> int:
        k = load().vsh_clipstab_feed(self.h, C.c_void_p(ptr), n, row_stride, frame_stride, mem, C.c_void_p(out_ptr),
                                     self.out_frame_bytes, out_mem)
        if k < 0:
            _raise("feed")
        return k

    def upload_only(self, first_frame: int, ptr: int, n: int, row_stride: int, frame_stride: int, mem: int):
        if load().vsh_clipstab_upload_only(self.h, first_frame, C.c_void_p(ptr), n, row_stride, frame_stride, mem) < 0:
            _raise("upload_only")

    def feed_resident(self, n: int, out_ptr: int, out_mem: int) -> int:
        k = load().vsh_clipstab_feed_resident(self.h, n, C.c_void_p(out_ptr), self.out_frame_bytes, out_mem)
        if k < 0:
            _raise("feed_resident")
        return k

    def last_records(self, n: int):
        meas, ok, corr = np.zeros((n, 4)), np.zeros(n, np.uint8), np.zeros((n, 4))
        k = load().vsh_clipstab_last_records(self.h, _p(meas), _p(ok), _p(corr))
        return meas, ok.astype(bool), corr[:k]

    @property
    def launches(self) -> int:
        return int(capi.load().vs_ctx_launch_count(self.ctx_handle))

    def set_stream(self, stream: int | None):
        capi.check(self.ctx_handle, capi.load().vs_ctx_set_stream(self.ctx_handle, C.c_void_p(stream or 0)), "vs_ctx_set_stream")

    def synchronize(self):
        capi.check(self.ctx_handle, capi.load().vs_ctx_synchronize(self.ctx_handle), "vs_ctx_synchronize")


class MultiGpuStabilizer(_Handle):
    """One video partitioned by frame chunk over several GPUs of one box (multi_gpu.hpp)."""
    _destroy = "vsh_multigpu_destroy"

    def __init__(self, devices, width, height, max_frames, params: VshStabParams | None = None):
        self.params = params or stab_params_default()
        self.width, self.height = width, height
        dev = np.asarray(list(devices), np.int32)
        self.h = C.c_void_p(load().vsh_multigpu_create(_p(dev), len(dev), width, height, max_frames, C.byref(self.params)))
        if not self.h:
            _raise("MultiGpuStabilizer")
        crop = max(0, self.params.crop_pixels)
        self.out_w, self.out_h = width - 2 * crop, height - 2 * crop

    def stabilize(self, frames: np.ndarray, out: np.ndarray | None = None):
        """frames (n,h,w,3) u8 -> (stabilized (n-lag,oh,ow,3), meas (n,4), ok (n,)).  The copies run at the PCIe rate
        when `frames` and `out` (optional, (n,oh,ow,3) u8, C-contiguous) are pinned host memory."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = frames.shape[0]
        if out is None:
            out = np.empty((n, self.out_h, self.out_w, 3), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (n, self.out_h, self.out_w, 3)
        meas, ok = np.zeros((n, 4)), np.zeros(n, np.uint8)
        k = load().vsh_multigpu_stabilize(self.h, _p(frames), n, frames.strides[1], frames.strides[0], _p(out),
                                          self.out_w * self.out_h * 3, _p(meas), _p(ok))
        if k < 0:
            _raise("stabilize")
        return out[:k], meas, ok.astype(bool)


class PartitionedTrajectory(_Handle):
    """Host half of a worker of a partitioned video (partitioned.hpp): no GPU involved."""
    _destroy = "vsh_parttraj_destroy"

    def __init__(self, rank, world, width, height, total_frames, sub_frames, block, params: VshStabParams | None = None,
                 exchange_name: str = "", host_threads: int = 2):
        self.params = params or stab_params_default()
        self.h = C.c_void_p(load().vsh_parttraj_create(rank, world, width, height, total_frames, sub_frames, block,
                                                       C.byref(self.params), exchange_name.encode(), host_threads))
        if not self.h:
            _raise("PartitionedTrajectory")
        self.outputs = load().vsh_parttraj_output_count(self.h)

    def run(self, meas_all, ok_all):
        """One video from its full measurement table: returns (frame indices, corrections) of this worker's outputs."""
        m = np.ascontiguousarray(meas_all, np.float64)
        o = np.ascontiguousarray(ok_all, np.uint8)
        corr, frames = np.zeros((self.outputs, 4)), np.zeros(self.outputs, np.int64)
        k = load().vsh_parttraj_run(self.h, _p(m), _p(o), _p(corr), _p(frames))
        if k < 0:
            _raise("PartitionedTrajectory.run")
        assert k == self.outputs, (k, self.outputs)
        return frames, corr


class PartitionedStabilizer(_Handle):
    """One worker (GPU) of ONE video partitioned by frame chunk over several workers (partitioned.hpp)."""
    _destroy = "vsh_partstab_destroy"

    def __init__(self, rank, world, width, height, total_frames, sub_frames, block, params: VshStabParams | None = None,
                 exchange_name: str = "", resident: bool = True, device: int = 0, host_threads: int = 4, lanes: int = 0,
                 nv12: bool = False):
        self.params = params or stab_params_default()
        self.width, self.height, self.nv12 = width, height, nv12
        create = load().vsh_partstab_create_nv12 if nv12 else load().vsh_partstab_create
        self.h = C.c_void_p(create(device, rank, world, width, height, total_frames, sub_frames, block,
                                                       C.byref(self.params), exchange_name.encode(), int(resident), host_threads, lanes))
        if not self.h:
            _raise("PartitionedStabilizer")
        lib = load()
        ow, oh = C.c_int(), C.c_int()
        lib.vsh_partstab_out_size(self.h, C.byref(ow), C.byref(oh))
        self.out_w, self.out_h = ow.value, oh.value
        self.out_frame_bytes = self.out_w * self.out_h * 3 // (2 if nv12 else 1)
        self.out_shape = (self.out_h * 3 // 2, self.out_w) if nv12 else (self.out_h, self.out_w, 3)
        self.ctx_handle = C.c_void_p(lib.vsh_partstab_context(self.h))
        n = lib.vsh_partstab_local_count(self.h)
        self.local_frames = np.array([lib.vsh_partstab_local_frame(self.h, i) for i in range(n)], np.int64)
        self.local_is_halo = np.array([lib.vsh_partstab_local_is_halo(self.h, i) for i in range(n)], bool)
        self.outputs = lib.vsh_partstab_output_count(self.h)
        self.output_frames = np.array([lib.vsh_partstab_output_frame(self.h, k) for k in range(self.outputs)], np.int64)

    def upload_resident(self, ptr: int, row_stride: int, frame_stride: int, mem: int = capi.VS_MEM_HOST):
        if load().vsh_partstab_upload_resident(self.h, C.c_void_p(ptr), row_stride, frame_stride, mem) < 0:
            _raise("upload_resident")

    def stabilize_ptr(self, frames_ptr: int | None, row_stride: int, frame_stride: int, out_ptr: int, out_mem: int) -> int:
        k = load().vsh_partstab_stabilize(self.h, C.c_void_p(frames_ptr or 0), row_stride, frame_stride, C.c_void_p(out_ptr),
                                          self.out_frame_bytes, out_mem)
        if k < 0:
            _raise("stabilize")
        return k

    def stabilize(self, local_frames: np.ndarray | None) -> np.ndarray:
        """local_frames: (local_count,h,w,3) u8 host array (None when resident); returns this worker's stabilized frames."""
        out = np.empty((max(self.outputs, 1),) + self.out_shape, np.uint8)
        if local_frames is None:
            k = self.stabilize_ptr(None, 0, 0, out.ctypes.data, capi.VS_MEM_HOST)
        else:
            f = np.ascontiguousarray(local_frames, np.uint8)
            k = self.stabilize_ptr(f.ctypes.data, f.strides[1], f.strides[0], out.ctypes.data, capi.VS_MEM_HOST)
        return out[:k]

    def set_lanes(self, lanes: int):
        load().vsh_partstab_set_lanes(self.h, int(lanes))

    def records(self, total_frames: int):
        corr, meas, ok = np.zeros((max(self.outputs, 1), 4)), np.zeros((total_frames, 4)), np.zeros(total_frames, np.uint8)
        seen = load().vsh_partstab_records(self.h, _p(corr), _p(meas), _p(ok))
        return corr[: self.outputs], meas[:seen], ok[:seen].astype(bool)

    @property
    def launches(self) -> int:
        return int(capi.load().vs_ctx_launch_count(self.ctx_handle))

    def set_stream(self, stream: int | None):
        capi.check(self.ctx_handle, capi.load().vs_ctx_set_stream(self.ctx_handle, C.c_void_p(stream or 0)), "vs_ctx_set_stream")

    def synchronize(self):
        capi.check(self.ctx_handle, capi.load().vs_ctx_synchronize(self.ctx_handle), "vs_ctx_synchronize")


def reference_grid() -> np.ndarray:
    """The 54 VideoAlignerParams combinations of grid_search_align.cpp:134-146 as rows {phase_correlate, threshold,
    smallest_fraction, max_displacement}."""
    g = np.zeros((64, 4))
    n = load().vsh_gridsearch_reference_grid(_p(g), 64)
    return g[:n].copy()


def flow_median_px(T, w, h) -> float:
    a = _t(T)
    return load().vsh_flow_median_px(_p(a), w, h)


class AlignerGridSearch(_Handle):
    """Jitter score and the batched parameter sweep (grid_search.hpp)."""
    _destroy = "vsh_gridsearch_destroy"

    def __init__(self, width, height, max_frames, max_combos=54, crop_pixels=32, device=0):
        self.width, self.height = width, height
        self.h = C.c_void_p(load().vsh_gridsearch_create(device, width, height, max_frames, max_combos, crop_pixels))
        if not self.h:
            _raise("AlignerGridSearch")

    def measure_jitter(self, frames):
        f = np.ascontiguousarray(frames, np.uint8)
        out = np.zeros(3)
        if load().vsh_gridsearch_jitter(self.h, _p(f), f.shape[0], f.strides[1], f.strides[0], _p(out)) < 0:
            _raise("measure_jitter")
        return {"median_px": float(out[0]), "pairs": int(out[1]), "failed": int(out[2])}

    def run(self, frames, combos):
        """Returns (input jitter dict, results (n_combos, 5), T (n_combos, n-1, 4), status (n_combos, n-1), launches)."""
        f = np.ascontiguousarray(frames, np.uint8)
        c = np.ascontiguousarray(combos, np.float64)
        n, S = f.shape[0], c.shape[0]
        inp, res = np.zeros(3), np.zeros((S, 5))
        T, st = np.zeros((S, n - 1, 4)), np.zeros((S, n - 1), np.int32)
        k = load().vsh_gridsearch_run(self.h, _p(f), n, f.strides[1], f.strides[0], _p(c), S, _p(inp), _p(res), _p(T), _p(st))
        if k < 0:
            _raise("AlignerGridSearch.run")
        return {"median_px": float(inp[0]), "pairs": int(inp[1]), "failed": int(inp[2])}, res, T, st, k
