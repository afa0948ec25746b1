"""ctypes binding of the CPU oracle (oracle/libvs_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class AlignParams(C.Structure):
    _fields_ = [("phase_correlate", C.c_int), ("phase_correlate_threshold", C.c_double), ("threshold", C.c_double),
                ("smallest_fraction", C.c_float), ("max_iters", C.c_int), ("pyramid_min_width", C.c_int),
                ("pyramid_min_height", C.c_int), ("max_displacement", C.c_double)]


class StabParams(C.Structure):
    _fields_ = [("aligner", AlignParams), ("lag", C.c_int), ("smoother_memory", C.c_int), ("lambda_", C.c_double),
                ("enable_smoother", C.c_int), ("crop_pixels", C.c_int), ("min_disp", C.c_double), ("max_disp", C.c_double),
                ("min_decay", C.c_double), ("max_decay", C.c_double)]


_libs = {}


def load(fast: bool = False):
    name = "libvs_oracle_fast.so" if fast else "libvs_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(path)
    P, I, D = C.c_void_p, C.c_int, C.c_double
    lib.vo_tile_size.restype = I
    lib.vo_tile_size.argtypes = [I, I]
    lib.vo_tf_max_corner_displacement.restype = D
    lib.vo_tf_max_corner_displacement.argtypes = [P, D, D]
    lib.vo_tf_warp.argtypes = [P, D, D, P]
    lib.vo_tf_warp_center.argtypes = [P, D, D, D, D, P]
    lib.vo_select_smallest.restype = I
    lib.vo_select_smallest.argtypes = [P, I, C.c_float, P]
    lib.vo_introselect_depth.restype = I
    lib.vo_introselect_depth.argtypes = [P, I, I, I, P]
    lib.vo_aligner_create.restype = P
    lib.vo_aligner_destroy.argtypes = [P]
    lib.vo_aligner_align.restype = I
    lib.vo_aligner_align.argtypes = [P, P, I, I, C.POINTER(AlignParams), P]
    lib.vo_aligner_levels.argtypes = [P]
    lib.vo_aligner_curr_index.argtypes = [P]
    lib.vo_aligner_level_info.argtypes = [P, I] + [C.POINTER(I)] * 5
    for f in ("vo_aligner_pyramid", "vo_aligner_keypoints", "vo_aligner_jacobians", "vo_aligner_warpdiff"):
        getattr(lib, f).restype = P
    lib.vo_aligner_pyramid.argtypes = [P, I, I]
    lib.vo_aligner_keypoints.argtypes = [P, I, I]
    lib.vo_aligner_jacobians.argtypes = [P, I, I]
    lib.vo_aligner_warpdiff.argtypes = [P, I, I]
    lib.vo_aligner_selected.argtypes = [P, I, I, C.POINTER(P)]
    lib.vo_aligner_iterations.argtypes = [P, I]
    lib.vo_tvl1_smooth.argtypes = [P, I, D, I, P]
    lib.vo_smoother_create.restype = P
    lib.vo_smoother_create.argtypes = [I, I, D]
    lib.vo_smoother_destroy.argtypes = [P]
    lib.vo_smoother_update.argtypes = [P, P, P]
    lib.vo_stabilizer_create.restype = P
    lib.vo_stabilizer_create.argtypes = [C.POINTER(StabParams)]
    lib.vo_stabilizer_destroy.argtypes = [P]
    lib.vo_stabilizer_process.argtypes = [P, P, I, I, P, C.POINTER(I), C.POINTER(I), C.POINTER(I), P, P]
    lib.vo_bgr2gray.argtypes = [P, I, I, P]
    lib.vo_pyr_down.argtypes = [P, I, I, P, I, I]
    lib.vo_grad_xy.argtypes = [P, I, I, P, P, I, I]
    lib.vo_grad_argmax.argtypes = [P, P, I, I, I, P, P]
    lib.vo_sparse_jac.argtypes = [P, P, I, I, P, P, I, I, P, P]
    lib.vo_sparse_warpdiff.argtypes = [P, P, I, I, P, I, I, P, P]
    lib.vo_sparse_ica.argtypes = [P, P, I, I, P, I, P, I, P, P, P, P]
    lib.vo_image_warp.argtypes = [P, I, I, P, P, I, I]
    lib.vo_warp_bgr.argtypes = [P, I, I, P, P, I, I, I]
    lib.vo_warp_bgr_matrix.argtypes = [P, I, I, P, P, I, I, I, I, I, I]
    lib.vo_warp_plane_matrix.argtypes = [P, I, I, I, P, P, I, I, I, I]
    lib.vo_warp_nv12.argtypes = [P, I, I, P, P, I]
    lib.vo_tf_inverse.argtypes = [P, P]
    lib.vo_optimal_dft_size.argtypes = [I]
    lib.vo_phase_correlate_u8.argtypes = [P, P, I, I, P]
    lib.vo_aligner_phase.argtypes = [P, P]
    lib.vo_tf_compose.argtypes = [P, P, P]
    lib.vo_svd4.argtypes = [P, P, P, P]
    lib.vo_inv4_svd.argtypes = [P, P]
    _libs[name] = lib
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _t(T):
    return np.ascontiguousarray(T, np.float64)


def align_params_default() -> AlignParams:
    p = AlignParams()
    load().vo_align_params_default(C.byref(p))
    return p


def stab_params_default() -> StabParams:
    p = StabParams()
    load().vo_stab_params_default(C.byref(p))
    return p


# ---- element ops on numpy arrays
def bgr2gray(bgr, fast=False):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    load(fast).vo_bgr2gray(_p(bgr), w, h, _p(out))
    return out


def phase_correlate_u8(a, b, fast=False):
    """cv::phaseCorrelate (alignment.cpp:374) of two u8 images taken as f32: (shift x, shift y, response)."""
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    assert a.shape == b.shape and a.ndim == 2
    out = np.zeros(3)
    load(fast).vo_phase_correlate_u8(_p(a), _p(b), a.shape[1], a.shape[0], _p(out))
    return out


def pyr_down(img, ow=None, oh=None, fast=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ow = w // 2 if ow is None else ow
    oh = h // 2 if oh is None else oh
    out = np.empty((oh, ow), np.uint8)
    load(fast).vo_pyr_down(_p(img), w, h, _p(out), ow, oh)
    return out


def grad_xy(img, ow=None, oh=None, fast=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ow = w if ow is None else ow
    oh = h if oh is None else oh
    gx = np.empty((oh, ow), np.float32)
    gy = np.empty((oh, ow), np.float32)
    load(fast).vo_grad_xy(_p(img), w, h, _p(gx), _p(gy), ow, oh)
    return gx, gy


def tile_size(w, h):
    return load().vo_tile_size(w, h)


def grad_argmax(gx, gy, tile):
    gx = np.ascontiguousarray(gx, np.float32)
    gy = np.ascontiguousarray(gy, np.float32)
    h, w = gx.shape
    tw, th = w // tile, h // tile
    lmx = np.zeros((2, th, tw), np.uint16)
    lmy = np.zeros((2, th, tw), np.uint16)
    load().vo_grad_argmax(_p(gx), _p(gy), w, h, tile, _p(lmx), _p(lmy))
    return lmx, lmy


def sparse_jac(gx, gy, lmx, lmy):
    gx = np.ascontiguousarray(gx, np.float32)
    gy = np.ascontiguousarray(gy, np.float32)
    h, w = gx.shape
    _, th, tw = lmx.shape
    jx = np.zeros((4, th, tw), np.float32)
    jy = np.zeros((4, th, tw), np.float32)
    lmx, lmy = np.ascontiguousarray(lmx, np.uint16), np.ascontiguousarray(lmy, np.uint16)
    load().vo_sparse_jac(_p(gx), _p(gy), w, h, _p(lmx), _p(lmy), tw, th, _p(jx), _p(jy))
    return jx, jy


def sparse_warpdiff(tmpl, key, lm, T):
    tmpl = np.ascontiguousarray(tmpl, np.uint8)
    key = np.ascontiguousarray(key, np.uint8)
    h, w = key.shape
    _, th, tw = lm.shape
    out = np.zeros((th, tw), np.uint16)
    T = _t(T)
    lm = np.ascontiguousarray(lm, np.uint16)
    load().vo_sparse_warpdiff(_p(tmpl), _p(key), w, h, _p(lm), tw, th, _p(T), _p(out))
    return out


def sparse_ica(tmpl, key, selx, sely, jx, jy, T):
    tmpl = np.ascontiguousarray(tmpl, np.uint8)
    key = np.ascontiguousarray(key, np.uint8)
    h, w = key.shape
    selx, sely = np.ascontiguousarray(selx, np.uint16), np.ascontiguousarray(sely, np.uint16)
    jx, jy = np.ascontiguousarray(jx, np.float32), np.ascontiguousarray(jy, np.float32)
    out = np.zeros(4, np.float64)
    T = _t(T)
    load().vo_sparse_ica(_p(tmpl), _p(key), w, h, _p(selx), selx.shape[1], _p(sely), sely.shape[1], _p(jx), _p(jy), _p(T), _p(out))
    return out


def image_warp(img, T, ow=None, oh=None):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ow = w if ow is None else ow
    oh = h if oh is None else oh
    out = np.empty((oh, ow), np.float32)
    T = _t(T)
    load().vo_image_warp(_p(img), w, h, _p(T), _p(out), ow, oh)
    return out


def warp_bgr(src, T, mode=0, border=0, crop=0, fast=False):
    src = np.ascontiguousarray(src, np.uint8)
    h, w, _ = src.shape
    out = np.empty((h - 2 * crop, w - 2 * crop, 3), np.uint8)
    T = _t(T)
    load(fast).vo_warp_bgr(_p(src), w, h, _p(T), _p(out), mode, border, crop)
    return out


def warp_bgr_matrix(src, M6, out_w, out_h, dx0=0, dy0=0, mode=0, border=0, fast=False, out=None):
    """cv::warpAffine(src, M (forward 2x3), Size(out_w + dx0.., ..)) restated: window (dx0, dy0, out_w, out_h) of the output."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w, _ = src.shape
    if out is None:
        out = np.empty((out_h, out_w, 3), np.uint8)
    assert out.flags.c_contiguous and out.shape == (out_h, out_w, 3)
    M = np.ascontiguousarray(M6, np.float64)
    load(fast).vo_warp_bgr_matrix(_p(src), w, h, _p(M), _p(out), out_w, out_h, dx0, dy0, mode, border)
    return out


def warp_plane_matrix(src, M6, out_w, out_h, dx0=0, dy0=0):
    """cv::warpAffine of a (h, w) or (h, w, 2) u8 image (the Y / UV plane of an NV12 frame), window (dx0, dy0, out_w, out_h)."""
    src = np.ascontiguousarray(src, np.uint8)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    out = np.empty((out_h, out_w) if src.ndim == 2 else (out_h, out_w, ch), np.uint8)
    M = np.ascontiguousarray(M6, np.float64)
    load().vo_warp_plane_matrix(_p(src), w, h, ch, _p(M), _p(out), out_w, out_h, dx0, dy0)
    return out


def warp_nv12(frame, w, h, T, crop=0):
    """A dense NV12 frame (h * 3 / 2 rows of w bytes) warped by the centre-based correction T."""
    frame = np.ascontiguousarray(frame, np.uint8).reshape(h * 3 // 2, w)
    out = np.empty(((h - 2 * crop) * 3 // 2, w - 2 * crop), np.uint8)
    t = _t(T)
    load().vo_warp_nv12(_p(frame), w, h, _p(t), _p(out), crop)
    return out


# NOTE: every array handed to ctypes is bound to a local first; `_p(_t(x))` alone would let
# the temporary be collected before the call.
def tf_inverse(T):
    out = np.zeros(4)
    a = _t(T)
    load().vo_tf_inverse(_p(a), _p(out))
    return out


def tf_compose(T1, T2):
    out = np.zeros(4)
    a, b = _t(T1), _t(T2)
    load().vo_tf_compose(_p(a), _p(b), _p(out))
    return out


def tf_warp(T, x, y, center=None):
    out = np.zeros(2)
    a = _t(T)
    if center is None:
        load().vo_tf_warp(_p(a), float(x), float(y), _p(out))
    else:
        load().vo_tf_warp_center(_p(a), float(x), float(y), float(center[0]), float(center[1]), _p(out))
    return out


def tf_max_corner_displacement(T, w, h):
    a = _t(T)
    return load().vo_tf_max_corner_displacement(_p(a), float(w), float(h))


def svd4(H):
    H = np.ascontiguousarray(H, np.float64)
    w, u, vt = np.zeros(4), np.zeros((4, 4)), np.zeros((4, 4))
    load().vo_svd4(_p(H), _p(w), _p(u), _p(vt))
    return w, u, vt


def inv4_svd(H):
    H = np.ascontiguousarray(H, np.float64)
    out = np.zeros((4, 4))
    load().vo_inv4_svd(_p(H), _p(out))
    return out


def select_smallest(abs_delta, fraction=0.8):
    a = np.ascontiguousarray(abs_delta, np.uint16).ravel()
    order = np.zeros(a.size, np.uint32)
    k = load().vo_select_smallest(_p(a), a.size, C.c_float(fraction), _p(order))
    return order[:k].copy()


def introselect_depth(abs_delta, nth, depth):
    a = np.ascontiguousarray(abs_delta, np.uint16).ravel()
    order = np.zeros(a.size, np.uint32)
    load().vo_introselect_depth(_p(a), a.size, nth, depth, _p(order))
    return order


def tvl1_smooth(data, lam, iterations=100):
    d = np.ascontiguousarray(data, np.float64)
    out = np.zeros_like(d)
    load().vo_tvl1_smooth(_p(d), d.size, lam, iterations, _p(out))
    return out


class Aligner:
    """vo_aligner: the restated VideoAligner (alignment.cpp:149-704)."""

    def __init__(self, params: AlignParams | None = None, fast: bool = False):
        self.lib = load(fast)
        self.h = C.c_void_p(self.lib.vo_aligner_create())
        self.params = params or align_params_default()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vo_aligner_destroy(self.h)
            self.h = None

    def align(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        T = np.zeros(4)
        ok = self.lib.vo_aligner_align(self.h, _p(bgr), w, h, C.byref(self.params), _p(T))
        return bool(ok), T

    @property
    def levels(self):
        return self.lib.vo_aligner_levels(self.h)

    def phase(self):
        """cv::phaseCorrelate result (shift x, shift y, response) of the last align (params.phase_correlate on)."""
        out = np.zeros(3)
        self.lib.vo_aligner_phase(self.h, _p(out))
        return out

    def level_info(self, level):
        v = [C.c_int() for _ in range(5)]
        self.lib.vo_aligner_level_info(self.h, level, *[C.byref(x) for x in v])
        w, h, tile, tw, th = [x.value for x in v]
        return dict(w=w, h=h, tile=tile, tw=tw, th=th)

    def _arr(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    def pyramid(self, slot, level):
        li = self.level_info(level)
        return self._arr(self.lib.vo_aligner_pyramid(self.h, slot, level), (li["h"], li["w"]), np.uint8)

    def keypoints(self, level, axis):
        li = self.level_info(level)
        return self._arr(self.lib.vo_aligner_keypoints(self.h, level, axis), (2, li["th"], li["tw"]), np.uint16)

    def jacobians(self, level, axis):
        li = self.level_info(level)
        return self._arr(self.lib.vo_aligner_jacobians(self.h, level, axis), (4, li["th"], li["tw"]), np.float32)

    def warpdiff(self, level, axis):
        li = self.level_info(level)
        return self._arr(self.lib.vo_aligner_warpdiff(self.h, level, axis), (li["th"], li["tw"]), np.uint16)

    def selected(self, level, axis):
        p = C.c_void_p()
        k = self.lib.vo_aligner_selected(self.h, level, axis, C.byref(p))
        if k == 0:
            return np.zeros(0, np.uint32)
        return self._arr(p.value, (k,), np.uint32)

    def iterations(self, level):
        return self.lib.vo_aligner_iterations(self.h, level)


class Smoother:
    def __init__(self, lag_behind, lag_ahead, lam):
        self.lib = load()
        self.h = C.c_void_p(self.lib.vo_smoother_create(lag_behind, lag_ahead, lam))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vo_smoother_destroy(self.h)
            self.h = None

    def update(self, meas):
        out = np.zeros(4)
        m = _t(meas)
        ok = self.lib.vo_smoother_update(self.h, _p(m), _p(out))
        return bool(ok), out


class Stabilizer:
    """vo_stabilizer: the restated VideoStabilizer (stabilizer.cpp:3-117)."""

    def __init__(self, params: StabParams | None = None, fast: bool = False):
        self.lib = load(fast)
        self.params = params or stab_params_default()
        self.h = C.c_void_p(self.lib.vo_stabilizer_create(C.byref(self.params)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vo_stabilizer_destroy(self.h)
            self.h = None

    def process(self, bgr):
        """Returns (frame or None, meas_ok, meas[4], correction[4])."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        self._max_bytes = max(getattr(self, "_max_bytes", 0), h * w * 3)   # the delayed frame may be larger
        out = np.empty(self._max_bytes, np.uint8)
        ow, oh, ok = C.c_int(), C.c_int(), C.c_int()
        meas, corr = np.zeros(4), np.zeros(4)
        has = self.lib.vo_stabilizer_process(self.h, _p(bgr), w, h, _p(out), C.byref(ow), C.byref(oh), C.byref(ok), _p(meas), _p(corr))
        frame = None
        if has:
            frame = out.ravel()[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()
        return frame, bool(ok.value), meas, corr


# ===================================================================== oracle/_ref
# The reference's OWN alignment.cpp / imgproc.cpp / smoother.cpp / stabilizer.cpp compiled
# unmodified against shim headers (oracle/Makefile `ref`).  Built here when /root/reference
# is present; on the GPU box only the prebuilt oracle/_ref/*.so exist.
REFERENCE_DIR = "/root/reference"
_ref_libs = {}


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libvs_ref.so")) or os.path.isdir(REFERENCE_DIR)


def load_ref(fast: bool = False):
    name = "libvs_ref_fast.so" if fast else "libvs_ref.so"
    if name in _ref_libs:
        return _ref_libs[name]
    path = os.path.join(_HERE, "_ref", name)
    if os.path.isdir(REFERENCE_DIR):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)   # no-op when up to date
    if not os.path.exists(path):
        raise RuntimeError("oracle/_ref is not built and %s is absent" % REFERENCE_DIR)
    lib = C.CDLL(path)
    P, I, D = C.c_void_p, C.c_int, C.c_double
    PI = C.POINTER(I)
    lib.vr_tf_inverse.argtypes = [P, P]
    lib.vr_tf_compose.argtypes = [P, P, P]
    lib.vr_tf_warp.argtypes = [P, D, D, P]
    lib.vr_tf_warp_center.argtypes = [P, D, D, D, D, P]
    lib.vr_tf_max_corner_displacement.restype = D
    lib.vr_tf_max_corner_displacement.argtypes = [P, D, D]
    lib.vr_aligner_create.restype = P
    lib.vr_aligner_destroy.argtypes = [P]
    lib.vr_aligner_align.argtypes = [P, P, I, I, C.POINTER(AlignParams), P]
    lib.vr_aligner_levels.argtypes = [P]
    lib.vr_aligner_curr_index.argtypes = [P]
    lib.vr_aligner_tile_size.argtypes = [P, I]
    for f in ("vr_aligner_pyramid", "vr_aligner_keypoints", "vr_aligner_jacobians", "vr_aligner_warpdiff",
              "vr_aligner_selected_pixels"):
        getattr(lib, f).restype = P
    lib.vr_aligner_pyramid.argtypes = [P, I, I, PI, PI]
    lib.vr_aligner_keypoints.argtypes = [P, I, I, PI, PI]
    lib.vr_aligner_jacobians.argtypes = [P, I, I]
    lib.vr_aligner_warpdiff.argtypes = [P, I, I]
    lib.vr_aligner_selected_pixels.argtypes = [P, I, I, PI]
    lib.vr_smoother_create.restype = P
    lib.vr_smoother_create.argtypes = [I, I, D]
    lib.vr_smoother_destroy.argtypes = [P]
    lib.vr_smoother_update.argtypes = [P, P, P]
    lib.vr_stabilizer_create.restype = P
    lib.vr_stabilizer_create.argtypes = [C.POINTER(StabParams)]
    lib.vr_stabilizer_destroy.argtypes = [P]
    lib.vr_stabilizer_process.argtypes = [P, P, I, I, P, PI, PI, P]
    _ref_libs[name] = lib
    return lib


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class RefAligner:
    """The reference's own VideoAligner (alignment.cpp), through oracle/_ref."""

    def __init__(self, params: AlignParams | None = None, fast: bool = False):
        self.lib = load_ref(fast)
        self.h = C.c_void_p(self.lib.vr_aligner_create())
        self.params = params or align_params_default()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vr_aligner_destroy(self.h)
            self.h = None

    def align(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        T = np.zeros(4)
        ok = self.lib.vr_aligner_align(self.h, _p(bgr), w, h, C.byref(self.params), _p(T))
        return bool(ok), T

    @property
    def levels(self):
        return self.lib.vr_aligner_levels(self.h)

    @property
    def curr(self):
        return self.lib.vr_aligner_curr_index(self.h)

    def tile_size(self, level):
        return self.lib.vr_aligner_tile_size(self.h, level)

    def pyramid(self, slot, level):
        w, h = C.c_int(), C.c_int()
        p = self.lib.vr_aligner_pyramid(self.h, slot, level, C.byref(w), C.byref(h))
        return _view(p, (h.value, w.value), np.uint8)

    def keypoints(self, level, axis):
        tw, th = C.c_int(), C.c_int()
        p = self.lib.vr_aligner_keypoints(self.h, level, axis, C.byref(tw), C.byref(th))
        return _view(p, (2, th.value, tw.value), np.uint16)

    def jacobians(self, level, axis):
        kp = self.keypoints(level, axis)
        return _view(self.lib.vr_aligner_jacobians(self.h, level, axis), (4,) + kp.shape[1:], np.float32)

    def warpdiff(self, level, axis):
        kp = self.keypoints(level, axis)
        return _view(self.lib.vr_aligner_warpdiff(self.h, level, axis), kp.shape[1:], np.uint16)

    def selected_pixels(self, level, axis):
        k = C.c_int()
        p = self.lib.vr_aligner_selected_pixels(self.h, level, axis, C.byref(k))
        return _view(p, (2, k.value), np.uint16)


class RefSmoother:
    def __init__(self, lag_behind, lag_ahead, lam):
        self.lib = load_ref()
        self.h = C.c_void_p(self.lib.vr_smoother_create(lag_behind, lag_ahead, lam))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vr_smoother_destroy(self.h)
            self.h = None

    def update(self, meas):
        out = np.zeros(4)
        m = _t(meas)
        ok = self.lib.vr_smoother_update(self.h, _p(m), _p(out))
        return bool(ok), out


class RefStabilizer:
    """The reference's own VideoStabilizer (stabilizer.cpp), through oracle/_ref."""

    def __init__(self, params: StabParams | None = None, fast: bool = False):
        self.lib = load_ref(fast)
        self.params = params or stab_params_default()
        self.h = C.c_void_p(self.lib.vr_stabilizer_create(C.byref(self.params)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vr_stabilizer_destroy(self.h)
            self.h = None

    def process(self, bgr):
        """Returns (frame or None, accum[4])."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        self._max_bytes = max(getattr(self, "_max_bytes", 0), h * w * 3)
        out = np.empty(self._max_bytes, np.uint8)
        ow, oh = C.c_int(), C.c_int()
        accum = np.zeros(4)
        has = self.lib.vr_stabilizer_process(self.h, _p(bgr), w, h, _p(out), C.byref(ow), C.byref(oh), _p(accum))
        frame = None
        if has:
            frame = out.ravel()[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()
        return frame, accum


def ref_tf(name, *args):
    """vr_tf_* through the reference's SimilarityTransform: name in inverse/compose/max_corner_displacement."""
    lib = load_ref()
    if name == "inverse":
        out, a = np.zeros(4), _t(args[0])
        lib.vr_tf_inverse(_p(a), _p(out))
        return out
    if name == "compose":
        out, a, b = np.zeros(4), _t(args[0]), _t(args[1])
        lib.vr_tf_compose(_p(a), _p(b), _p(out))
        return out
    if name == "max_corner_displacement":
        a = _t(args[0])
        return lib.vr_tf_max_corner_displacement(_p(a), float(args[1]), float(args[2]))
    if name == "warp_center":
        out, a = np.zeros(2), _t(args[0])
        lib.vr_tf_warp_center(_p(a), float(args[1]), float(args[2]), float(args[3]), float(args[4]), _p(out))
        return out
    raise ValueError(name)
