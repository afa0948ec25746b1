"""Synthetic procedurally-jittered clips (BASELINE.json configs).

Texture: multi-octave Gaussian-blurred uniform noise on an oversized canvas.  Frame t is the
canvas seen through a similarity about the frame centre (random-walk translation, small
random zoom/rotation), rendered with the bit-exact fixed-point bilinear warp — integer
arithmetic, so the numpy renderer here and the GPU renderer (vs_bgr_warp_u8 on the canvas)
produce identical bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MARGIN = 64


def make_canvas(width: int, height: int, seed: int, tint=(0, 0, 0)) -> np.ndarray:
    """(height+2*MARGIN, width+2*MARGIN, 3) uint8 BGR texture; gray replicated (+ optional tint)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    H, W = height + 2 * MARGIN, width + 2 * MARGIN
    acc = np.zeros((H, W), np.float32)
    for weight, sigma in ((1.0, 1.5), (2.0, 4.0), (3.0, 12.0)):
        n = rng.integers(0, 256, size=(H, W), dtype=np.uint8).astype(np.float32)
        g = gaussian_filter(n, sigma, mode="nearest")
        g = (g - g.mean()) / (g.std() + 1e-6)
        acc += weight * g
    lo, hi = np.percentile(acc, 0.5), np.percentile(acc, 99.5)
    gray = np.clip((acc - lo) / (hi - lo) * 255.0, 0, 255).astype(np.uint8)
    bgr = np.repeat(gray[:, :, None], 3, axis=2)
    if any(tint):
        bgr = np.clip(bgr.astype(np.int16) + np.array(tint, np.int16)[None, None, :], 0, 255).astype(np.uint8)
    return np.ascontiguousarray(bgr)


def jitter_path(n_frames: int, seed: int, step: float = 2.0, limit: float = 16.0, ab: float = 0.001) -> np.ndarray:
    """(n,4) f64 camera poses {A,B,TX,TY} about the frame centre: frame t shows canvas(W_t(p))."""
    rng = np.random.default_rng(seed)
    T = np.zeros((n_frames, 4), np.float64)
    tx = ty = 0.0
    for t in range(n_frames):
        if t > 0:
            tx = float(np.clip(tx + rng.uniform(-step, step), -limit, limit))
            ty = float(np.clip(ty + rng.uniform(-step, step), -limit, limit))
        T[t] = (rng.uniform(-ab, ab), rng.uniform(-ab, ab), tx, ty)
    return T


def pose_to_inverse_map(pose, width: int, height: int):
    """Inverse-map coefficients (dst pixel -> canvas pixel) of one pose, canvas coordinates."""
    A, B, TX, TY = [float(v) for v in pose]
    cx, cy = (width - 1) * 0.5, (height - 1) * 0.5
    # frame pixel p -> canvas pixel: rotate/scale about the frame centre, translate, add margin
    i00, i01 = 1.0 + A, -B
    i10, i11 = B, 1.0 + A
    i02 = TX - A * cx + B * cy + MARGIN
    i12 = TY - B * cx - A * cy + MARGIN
    return i00, i01, i02, i10, i11, i12


def forward_matrix_for_pose(pose, width: int, height: int) -> np.ndarray:
    """2x3 forward matrix M (as cv::warpAffine takes it without WARP_INVERSE_MAP) whose f64
    inverse is the pose's inverse map; used to render through vs_bgr_warp_u8."""
    i00, i01, i02, i10, i11, i12 = pose_to_inverse_map(pose, width, height)
    D = 1.0 / (i00 * i11 - i01 * i10)
    m00, m01, m10, m11 = i11 * D, -i01 * D, -i10 * D, i00 * D
    m02 = -m00 * i02 - m01 * i12
    m12 = -m10 * i02 - m11 * i12
    return np.array([m00, m01, m02, m10, m11, m12], np.float64)


def render_frame_numpy(canvas: np.ndarray, M6: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT) in integer numpy arithmetic (forward matrix M6)."""
    m00, m01, m02, m10, m11, m12 = [float(v) for v in M6]
    D = m00 * m11 - m01 * m10
    D = 1.0 / D if D != 0 else 0.0
    i00, i01, i10, i11 = m11 * D, m01 * (-D), m10 * (-D), m00 * D
    i02 = -i00 * m02 - i01 * m12
    i12 = -i10 * m02 - i11 * m12
    ch, cw, _ = canvas.shape
    x = np.arange(width, dtype=np.float64)
    y = np.arange(height, dtype=np.float64)
    adelta = np.rint(i00 * x * 1024).astype(np.int64)
    bdelta = np.rint(i10 * x * 1024).astype(np.int64)
    X0 = np.rint((i01 * y + i02) * 1024).astype(np.int64) + 16
    Y0 = np.rint((i11 * y + i12) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31

    def tap(xx, yy):
        inside = (xx >= 0) & (xx < cw) & (yy >= 0) & (yy < ch)
        v = canvas[np.clip(yy, 0, ch - 1), np.clip(xx, 0, cw - 1)].astype(np.int64)
        return v * inside[:, :, None]

    w00 = ((32 - fx) * (32 - fy))[:, :, None]
    w10 = (fx * (32 - fy))[:, :, None]
    w01 = ((32 - fx) * fy)[:, :, None]
    w11 = (fx * fy)[:, :, None]
    v = w00 * tap(sx, sy) + w10 * tap(sx + 1, sy) + w01 * tap(sx, sy + 1) + w11 * tap(sx + 1, sy + 1)
    return ((v + 512) >> 10).astype(np.uint8)


def make_clip_numpy(width: int, height: int, n_frames: int, seed: int, **jitter) -> tuple[np.ndarray, np.ndarray]:
    """(frames (n,h,w,3) u8, poses (n,4)).  Pure numpy: for test sizes."""
    canvas = make_canvas(width, height, 1000 + seed)
    poses = jitter_path(n_frames, 1001 + seed, **jitter)
    frames = np.empty((n_frames, height, width, 3), np.uint8)
    for t in range(n_frames):
        frames[t] = render_frame_numpy(canvas, forward_matrix_for_pose(poses[t], width, height), width, height)
    return frames, poses


def make_clip_gpu(ctx, width: int, height: int, n_frames: int, seed: int, out: np.ndarray | None = None,
                  chunk: int = 32, **jitter) -> tuple[np.ndarray, np.ndarray]:
    """Same clip rendered on the GPU through vs_bgr_warp_u8 (bit-identical to the numpy path)."""
    from . import _capi as capi
    canvas = make_canvas(width, height, 1000 + seed)
    poses = jitter_path(n_frames, 1001 + seed, **jitter)
    frames = out if out is not None else np.empty((n_frames, height, width, 3), np.uint8)
    ch, cw, _ = canvas.shape
    for t0 in range(0, n_frames, chunk):
        n = min(chunk, n_frames - t0)
        M = np.stack([forward_matrix_for_pose(poses[t0 + i], width, height) for i in range(n)])
        src = capi.VsImg(canvas.ctypes.data, cw, ch, canvas.strides[0], n, 0)   # batch_stride 0: same canvas
        dst_arr = frames[t0:t0 + n]
        dst = capi.VsImg(dst_arr.ctypes.data, width, height, dst_arr.strides[1], n, dst_arr.strides[0])
        capi.check(ctx.handle, ctx.lib.vs_bgr_warp_u8(ctx.handle, C.byref(src), capi.ptr(M), C.byref(dst), 0, 0,
                                                      capi.VS_WARP_CV_EXACT_BILINEAR, capi.VS_BORDER_CONSTANT0,
                                                      capi.VS_MEM_HOST), "vs_bgr_warp_u8")
    return frames, poses


def render_frames_gpu(ctx, canvas: np.ndarray, poses: np.ndarray, frame_indices, width: int, height: int,
                      out: np.ndarray, chunk: int = 32) -> np.ndarray:
    """Frames `frame_indices` of the clip (canvas, poses) into out[i] through vs_bgr_warp_u8: what make_clip_gpu does for a
    whole clip, for any subset of it (a worker's share of a partitioned video)."""
    from . import _capi as capi
    ch, cw, _ = canvas.shape
    idx = list(frame_indices)
    for t0 in range(0, len(idx), chunk):
        n = min(chunk, len(idx) - t0)
        M = np.stack([forward_matrix_for_pose(poses[idx[t0 + i]], width, height) for i in range(n)])
        src = capi.VsImg(canvas.ctypes.data, cw, ch, canvas.strides[0], n, 0)
        dst_arr = out[t0:t0 + n]
        dst = capi.VsImg(dst_arr.ctypes.data, width, height, dst_arr.strides[1], n, dst_arr.strides[0])
        capi.check(ctx.handle, ctx.lib.vs_bgr_warp_u8(ctx.handle, C.byref(src), capi.ptr(M), C.byref(dst), 0, 0,
                                                      capi.VS_WARP_CV_EXACT_BILINEAR, capi.VS_BORDER_CONSTANT0,
                                                      capi.VS_MEM_HOST), "vs_bgr_warp_u8")
    return out
