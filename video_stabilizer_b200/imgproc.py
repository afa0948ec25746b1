"""Python mirror of the reference's operator API (imgproc.hpp) on top of the C ABI.

Same names, argument meaning and error behaviour as imgproc.hpp:8-97 — each wrapper returns
a bool like the reference (True = OK) and, where the reference's wrapper (re)allocates its
output, returns the new array.  Arrays are host numpy arrays (the reference's
Halide::Runtime::Buffer are host buffers); every call goes through libvstab.so
(VS_MEM_HOST) and runs on the GPU.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _capi as capi


class Context:
    """One GPU context (vs_ctx).  `stream` may be a raw cudaStream_t to borrow."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(None, self.lib.vs_ctx_create(device, C.byref(h)), "vs_ctx_create")
        self.handle = h
        self.device = device
        if stream:
            self.set_stream(stream)

    def set_stream(self, stream: int | None) -> None:
        capi.check(self.handle, self.lib.vs_ctx_set_stream(self.handle, C.c_void_p(stream or 0)), "vs_ctx_set_stream")

    def synchronize(self) -> None:
        capi.check(self.handle, self.lib.vs_ctx_synchronize(self.handle), "vs_ctx_synchronize")

    @property
    def launches(self) -> int:
        return int(self.lib.vs_ctx_launch_count(self.handle))

    def close(self) -> None:
        if self.handle:
            self.lib.vs_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ----------------------------------------------------------------- transform algebra
@dataclass
class Point:
    x: float = 0.0
    y: float = 0.0

    def distance(self, p: "Point") -> float:   # imgproc.cpp:413-417
        dx, dy = self.x - p.x, self.y - p.y
        return math.sqrt(dx * dx + dy * dy)


@dataclass
class SimilarityTransform:
    """imgproc.hpp:40-65.  W(p) = [(1+A)x - By + TX, Bx + (1+A)y + TY]."""
    A: float = 0.0
    B: float = 0.0
    TX: float = 0.0
    TY: float = 0.0

    def toString(self) -> str:
        return "A=%g, B=%g, TX=%g, TY=%g" % (self.A, self.B, self.TX, self.TY)

    def inverse(self) -> "SimilarityTransform":   # imgproc.cpp:333-359
        p, q = 1.0 + self.A, self.B
        denom = p * p + q * q
        return SimilarityTransform((p / denom) - 1.0, -q / denom,
                                   (-p * self.TX - q * self.TY) / denom, (q * self.TX - p * self.TY) / denom)

    def compose(self, w2: "SimilarityTransform") -> "SimilarityTransform":   # imgproc.cpp:361-387
        p1, q1, p2, q2 = 1.0 + self.A, self.B, 1.0 + w2.A, w2.B
        return SimilarityTransform((p2 * p1 - q2 * q1) - 1.0, (p2 * q1 + q2 * p1),
                                   p2 * self.TX - q2 * self.TY + w2.TX, q2 * self.TX + p2 * self.TY + w2.TY)

    def warp(self, p: Point, cx: float | None = None, cy: float | None = None) -> Point:
        if cx is None:   # imgproc.cpp:389-394
            return Point((1 + self.A) * p.x - self.B * p.y + self.TX, self.B * p.x + (1 + self.A) * p.y + self.TY)
        px, py = p.x - cx, p.y - cy   # imgproc.cpp:401-411
        return Point((1 + self.A) * px - self.B * py + cx + self.TX, self.B * px + (1 + self.A) * py + cy + self.TY)

    def maxCornerDisplacement(self, width: float, height: float) -> float:   # imgproc.cpp:419-437
        cx, cy = width * 0.5, height * 0.5
        d = 0.0
        for c in (Point(0.0, 0.0), Point(width, 0.0), Point(0.0, height), Point(width, height)):
            d = max(d, self.warp(c, cx, cy).distance(c))
        return d

    def as_array(self) -> np.ndarray:
        return np.array([self.A, self.B, self.TX, self.TY], dtype=np.float64)


def _ul_params_half(t: SimilarityTransform, w: int, h: int):
    """imgproc.cpp:69-75 / :98-103 — (w*0.5f) is an f32 product promoted to f64."""
    hw, hh = float(np.float32(w) * np.float32(0.5)), float(np.float32(h) * np.float32(0.5))
    return (np.float32(t.A), np.float32(t.B),
            np.float32(t.TX - t.A * hw + t.B * hh), np.float32(t.TY - t.B * hw - t.A * hh))


def _dense_rows(a, dtype):
    """Row-strided views are passed through as they are (the ABI takes a row stride); only
    arrays whose rows are not dense get copied."""
    a = np.asarray(a)
    if a.dtype != dtype:
        raise TypeError("expected a %s image" % np.dtype(dtype).name)
    inner = a.strides[1:] == tuple(int(np.prod(a.shape[i + 1:])) * a.itemsize for i in range(1, a.ndim))
    if not inner or a.strides[0] < 0 or a.strides[0] % a.itemsize:
        a = np.ascontiguousarray(a)
    return a


def _u8(a):
    return _dense_rows(a, np.uint8)


def _f32(a):
    return _dense_rows(a, np.float32)


# ----------------------------------------------------------------- operator wrappers
def BGR2Gray(bgr: np.ndarray, ctx: Context | None = None) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) as used at alignment.cpp:212."""
    ctx = ctx or default_context()
    bgr = _u8(bgr)
    out = np.empty(bgr.shape[:2], np.uint8)
    capi.check(ctx.handle, ctx.lib.vs_bgr2gray_u8(ctx.handle, C.byref(capi.img_of(bgr)), C.byref(capi.img_of(out)),
                                                  capi.VS_MEM_HOST), "vs_bgr2gray_u8")
    return out


def IngestBGR(bgr: np.ndarray, ctx: Context | None = None):
    """Fused ingest (alignment.cpp:212 + the first PyrDown of :220-223): returns (gray level 0, gray level 1) with
    level 1 = (w // 2) x (h // 2), computed in one pass over the BGR frame when its geometry allows."""
    ctx = ctx or default_context()
    bgr = _u8(bgr)
    h, w = bgr.shape[:2]
    g0 = np.empty((h, w), np.uint8)
    g1 = np.empty((max(h // 2, 0), max(w // 2, 0)), np.uint8)
    capi.check(ctx.handle, ctx.lib.vs_ingest_bgr_u8(ctx.handle, C.byref(capi.img_of(bgr)), C.byref(capi.img_of(g0)),
                                                    C.byref(capi.img_of(g1)), capi.VS_MEM_HOST), "vs_ingest_bgr_u8")
    return g0, g1


def PhaseCorrelate(src1: np.ndarray, src2: np.ndarray, ctx: Context | None = None):
    """cv::phaseCorrelate(src1, src2, noArray(), &response) as called at alignment.cpp:374 (u8-valued images taken
    as CV_32F): returns ((shift_x, shift_y), response)."""
    ctx = ctx or default_context()
    a = _u8(src1)
    b = _u8(src2)
    out = np.zeros(3, np.float64)
    capi.check(ctx.handle, ctx.lib.vs_phase_correlate_u8(ctx.handle, C.byref(capi.img_of(a)), C.byref(capi.img_of(b)),
                                                         capi.ptr(out), capi.VS_MEM_HOST), "vs_phase_correlate_u8")
    return (float(out[0]), float(out[1])), float(out[2])


# NOTE: arrays are bound to locals before their address is taken: a descriptor only holds
# the raw pointer, so a temporary contiguous copy must outlive the call.
def PyrDown(input: np.ndarray, output: np.ndarray, ctx: Context | None = None) -> bool:
    """imgproc.hpp:16-18.  Caller allocates `output`; its extent defines the work."""
    ctx = ctx or default_context()
    src = _u8(input)
    r = ctx.lib.vs_pyr_down_u8(ctx.handle, C.byref(capi.img_of(src)), C.byref(capi.img_of(output)), capi.VS_MEM_HOST)
    return r == 0


def GradXY(input: np.ndarray, output_x: np.ndarray, output_y: np.ndarray, ctx: Context | None = None) -> bool:
    """imgproc.hpp:20-23."""
    ctx = ctx or default_context()
    src = _u8(input)
    r = ctx.lib.vs_grad_xy_u8_f32(ctx.handle, C.byref(capi.img_of(src)), C.byref(capi.img_of(output_x)),
                                  C.byref(capi.img_of(output_y)), capi.VS_MEM_HOST)
    return r == 0


def GradArgMax(grad_x: np.ndarray, grad_y: np.ndarray, ctx: Context | None = None):
    """imgproc.hpp:25-32.  Returns (ok, tile_size, local_max_x, local_max_y); the local_max
    arrays are shaped (2, th, tw) = Halide planar (tw,th,2)."""
    ctx = ctx or default_context()
    h, w = grad_x.shape
    tile = ctx.lib.vs_grad_argmax_tile_size(w, grad_y.shape[0])
    tw, th = w // tile, grad_y.shape[0] // tile
    lmx = np.zeros((2, th, tw), np.uint16)
    lmy = np.zeros((2, th, tw), np.uint16)
    grad_x, grad_y = _f32(grad_x), _f32(grad_y)
    r = ctx.lib.vs_grad_argmax_f32_u16(ctx.handle, C.byref(capi.img_of(grad_x)), C.byref(capi.img_of(grad_y)), tile,
                                       capi.ptr(lmx), capi.ptr(lmy), capi.VS_MEM_HOST)
    return r == 0, tile, lmx, lmy


def SparseJacobian(grad_x, grad_y, local_max_x, local_max_y, ctx: Context | None = None):
    """imgproc.hpp:8-14.  Returns (ok, output_x, output_y) shaped (4, th, tw)."""
    ctx = ctx or default_context()
    _, th, tw = local_max_x.shape
    ox = np.zeros((4, th, tw), np.float32)
    oy = np.zeros((4, th, tw), np.float32)
    grad_x, grad_y = _f32(grad_x), _f32(grad_y)
    lx, ly = np.ascontiguousarray(local_max_x, np.uint16), np.ascontiguousarray(local_max_y, np.uint16)
    r = ctx.lib.vs_sparse_jac_f32(ctx.handle, C.byref(capi.img_of(grad_x)), C.byref(capi.img_of(grad_y)),
                                  capi.ptr(lx), capi.ptr(ly), tw, th, capi.ptr(ox), capi.ptr(oy), capi.VS_MEM_HOST)
    return r == 0, ox, oy


def SparseWarpDiff(input_template, input_keyframe, local_max, transform: SimilarityTransform, ctx: Context | None = None):
    """imgproc.hpp:90-95.  Returns (ok, output) with output shaped (th, tw) u16."""
    ctx = ctx or default_context()
    _, th, tw = local_max.shape
    h, w = input_template.shape
    A, B, TX, TY = _ul_params_half(transform, w, h)
    out = np.zeros((th, tw), np.uint16)
    tm, kf, lm = _u8(input_template), _u8(input_keyframe), np.ascontiguousarray(local_max, np.uint16)
    r = ctx.lib.vs_sparse_warpdiff_u8_u16(ctx.handle, C.byref(capi.img_of(tm)), C.byref(capi.img_of(kf)),
                                          capi.ptr(lm), tw, th, A, B, TX, TY, capi.ptr(out), capi.VS_MEM_HOST)
    return r == 0, out


def SparseICA(input_template, input_keyframe, selected_pixels_x, selected_pixels_y,
              selected_jacobians_x, selected_jacobians_y, transform: SimilarityTransform, ctx: Context | None = None):
    """imgproc.hpp:78-88.  selected_pixels_*: (2,k) u16; selected_jacobians_*: (4,k) f32.
    Returns (ok, output[4] f64)."""
    ctx = ctx or default_context()
    h, w = input_template.shape
    A, B, TX, TY = _ul_params_half(transform, w, h)
    out = np.zeros(4, np.float64)
    sx, sy = np.ascontiguousarray(selected_pixels_x, np.uint16), np.ascontiguousarray(selected_pixels_y, np.uint16)
    jx, jy = np.ascontiguousarray(selected_jacobians_x, np.float32), np.ascontiguousarray(selected_jacobians_y, np.float32)
    tm, kf = _u8(input_template), _u8(input_keyframe)
    r = ctx.lib.vs_sparse_ica_f64(ctx.handle, C.byref(capi.img_of(tm)), C.byref(capi.img_of(kf)),
                                  capi.ptr(sx), sx.shape[1], capi.ptr(sy), sy.shape[1], capi.ptr(jx), capi.ptr(jy),
                                  A, B, TX, TY, capi.ptr(out), capi.VS_MEM_HOST)
    return r == 0, out


def ImageWarp(input: np.ndarray, transform: SimilarityTransform, output: np.ndarray, ctx: Context | None = None) -> bool:
    """imgproc.hpp:67-70 / imgproc.cpp:116-133 (centre = (w-1)/2, parameters passed as f32)."""
    ctx = ctx or default_context()
    h, w = input.shape
    cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
    p = np.array([transform.A, transform.B, transform.TX - transform.A * cx + transform.B * cy,
                  transform.TY - transform.B * cx - transform.A * cy], dtype=np.float64).astype(np.float32)
    src = _u8(input)
    r = ctx.lib.vs_image_warp_u8_f32(ctx.handle, C.byref(capi.img_of(src)), capi.ptr(p),
                                     C.byref(capi.img_of(output)), capi.VS_MEM_HOST)
    return r == 0


def forward_matrix(transform: SimilarityTransform, cols: int, rows: int) -> np.ndarray:
    """imgproc.cpp:458-466."""
    cx, cy = (cols - 1) * 0.5, (rows - 1) * 0.5
    tx = transform.TX - transform.A * cx + transform.B * cy
    ty = transform.TY - transform.B * cx - transform.A * cy
    return np.array([1.0 + transform.A, -transform.B, tx, transform.B, 1.0 + transform.A, ty], np.float64)


def warpBySimilarityTransform(src: np.ndarray, transform: SimilarityTransform, ctx: Context | None = None,
                              mode: int = capi.VS_WARP_CV_EXACT_BILINEAR, border: int = capi.VS_BORDER_CONSTANT0,
                              crop: int = 0) -> np.ndarray:
    """imgproc.hpp:97 / imgproc.cpp:446-484: returns a freshly allocated warped BGR frame."""
    ctx = ctx or default_context()
    src = _u8(src)
    h, w, _ = src.shape
    M = forward_matrix(transform, w, h)
    dst = np.empty((h - 2 * crop, w - 2 * crop, 3), np.uint8)
    capi.check(ctx.handle, ctx.lib.vs_bgr_warp_u8(ctx.handle, C.byref(capi.img_of(src)), capi.ptr(M), C.byref(capi.img_of(dst)),
                                                  crop, crop, mode, border, capi.VS_MEM_HOST), "vs_bgr_warp_u8")
    return dst


def PlaneWarp(src: np.ndarray, M6, out_w: int | None = None, out_h: int | None = None, dx0: int = 0, dy0: int = 0,
              ctx: Context | None = None) -> np.ndarray:
    """cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of a (h, w) or (h, w, 2) u8 image (the Y / UV plane of an NV12 frame;
    vs_plane_warp_u8): window (dx0, dy0, out_w, out_h) of the output under the forward 2x3 matrix M6."""
    ctx = ctx or default_context()
    src = _u8(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    out_w = w if out_w is None else out_w
    out_h = h if out_h is None else out_h
    dst = np.empty((out_h, out_w) if src.ndim == 2 else (out_h, out_w, ch), np.uint8)
    M = np.ascontiguousarray(M6, np.float64)
    simg = capi.VsImg(src.ctypes.data, w, h, src.strides[0], 1, 0)
    dimg = capi.VsImg(dst.ctypes.data, out_w, out_h, dst.strides[0], 1, 0)
    capi.check(ctx.handle, ctx.lib.vs_plane_warp_u8(ctx.handle, C.byref(simg), ch, capi.ptr(M), C.byref(dimg), dx0, dy0,
                                                    capi.VS_MEM_HOST), "vs_plane_warp_u8")
    return dst
