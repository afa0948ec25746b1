"""Device-resident clip: the batched half of the C ABI (vs_clip_*).

A Clip holds `capacity` frames of one size on the GPU (BGR, gray pyramid, keyframe
features) and runs the per-frame hot path of the reference in a handful of launches:
upload -> build_pyramids -> build_keyframes -> align -> warp.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .imgproc import Context, default_context


def pairs_for_frames(first_frame: int, n_frames: int, slot_of=lambda f: f):
    """Alignment jobs of frames first_frame+1 .. first_frame+n_frames-1 against their
    predecessor, with the reference's keyframe alternation (alignment.cpp:357,396-397,690-693):
    odd frames (0-based) are keyframes.  Pair i aligns frame f=first_frame+1+i:
      f odd  -> keyframe = f,   template = f-1, result as is
      f even -> keyframe = f-1, template = f,   result inverted
    Returns (pairs array, sorted list of keyframe frame indices)."""
    pairs = (capi.VsPair * max(n_frames - 1, 0))()
    keyframes = set()
    for i in range(n_frames - 1):
        f = first_frame + 1 + i
        if f % 2 == 1:
            pairs[i] = capi.VsPair(slot_of(f - 1), slot_of(f), 0)
            keyframes.add(f)
        else:
            pairs[i] = capi.VsPair(slot_of(f), slot_of(f - 1), 1)
            keyframes.add(f - 1)
    return pairs, sorted(keyframes)


class Clip:
    def __init__(self, width: int, height: int, capacity: int, max_pairs: int | None = None,
                 params: capi.VsAlignParams | None = None, debug: bool = False, ctx: Context | None = None,
                 nv12: bool = False):
        """nv12: the clip's frames are NV12 (h * 3 / 2 rows of w bytes: Y, then interleaved UV) instead of BGR."""
        self.nv12 = nv12
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.width, self.height, self.capacity = width, height, capacity
        self.max_pairs = capacity if max_pairs is None else max_pairs
        if params is None:
            params = capi.VsAlignParams()
            self.lib.vs_align_params_default(C.byref(params))
        self.params = params
        h = C.c_void_p()
        capi.check(self.ctx.handle, self.lib.vs_clip_create(self.ctx.handle, width, height, capacity, self.max_pairs,
                                                            C.byref(params),
                                                            (capi.VS_CLIP_DEBUG_TAPS if debug else 0) | (capi.VS_CLIP_NV12 if nv12 else 0),
                                                            C.byref(h)), "vs_clip_create")
        self.handle = h
        self.levels = self.lib.vs_clip_levels(h)

    def _chk(self, code, what):
        capi.check(self.ctx.handle, code, what)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vs_clip_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_info(self, level: int):
        v = [C.c_int() for _ in range(5)]
        self._chk(self.lib.vs_clip_level_info(self.handle, level, *[C.byref(x) for x in v]), "vs_clip_level_info")
        w, h, tile, tw, th = [x.value for x in v]
        return dict(w=w, h=h, tile=tile, tw=tw, th=th)

    # ---- pipeline stages
    def upload(self, slot0: int, frames: np.ndarray):
        """frames: (n,h,w,3) uint8 host array; NV12 clips: (n, h*3/2, w)."""
        frames = np.ascontiguousarray(frames)
        n = frames.shape[0]
        assert frames.shape[1:] == ((self.height * 3 // 2, self.width) if self.nv12 else (self.height, self.width, 3))
        self._chk(self.lib.vs_clip_upload(self.handle, slot0, n, capi.ptr(frames), frames.strides[1], frames.strides[0],
                                          capi.VS_MEM_HOST), "vs_clip_upload")

    def upload_ptr(self, slot0: int, n: int, data_ptr: int, row_stride: int, frame_stride: int, mem: int):
        self._chk(self.lib.vs_clip_upload(self.handle, slot0, n, C.c_void_p(data_ptr), row_stride, frame_stride, mem),
                  "vs_clip_upload")

    def build_pyramids(self, slot0: int, n: int):
        self._chk(self.lib.vs_clip_build_pyramids(self.handle, slot0, n), "vs_clip_build_pyramids")

    def build_keyframes(self, slots):
        a = np.asarray(list(slots), np.int32)
        self._chk(self.lib.vs_clip_build_keyframes(self.handle, capi.ptr(a), len(a)), "vs_clip_build_keyframes")

    def align(self, pairs, n: int | None = None):
        """Returns (transforms (n,4) f64, status (n,) i32, iters (n,levels) i32)."""
        n = len(pairs) if n is None else n
        T = np.zeros((n, 4), np.float64)
        st = np.zeros(n, np.int32)
        it = np.zeros((n, self.levels), np.int32)
        self._chk(self.lib.vs_clip_align(self.handle, C.cast(pairs, C.c_void_p), n, capi.ptr(T), capi.ptr(st), capi.ptr(it),
                                         capi.VS_MEM_HOST), "vs_clip_align")
        return T, st, it

    def align_device(self, pairs, n: int, d_T: int, d_status: int, d_iters: int = 0):
        self._chk(self.lib.vs_clip_align(self.handle, C.cast(pairs, C.c_void_p), n, C.c_void_p(d_T), C.c_void_p(d_status),
                                         C.c_void_p(d_iters), capi.VS_MEM_DEVICE), "vs_clip_align")

    def warp(self, slots, transforms: np.ndarray, mode=capi.VS_WARP_CV_EXACT_BILINEAR, border=capi.VS_BORDER_CONSTANT0,
             crop: int = 0) -> np.ndarray:
        s = np.asarray(list(slots), np.int32)
        T = np.ascontiguousarray(transforms, np.float64).reshape(len(s), 4)
        ow, oh = self.width - 2 * crop, self.height - 2 * crop
        out = np.empty((len(s), oh * 3 // 2, ow) if self.nv12 else (len(s), oh, ow, 3), np.uint8)
        self._chk(self.lib.vs_clip_warp(self.handle, capi.ptr(s), len(s), capi.ptr(T), mode, border, crop, capi.ptr(out),
                                        out[0].nbytes if len(s) else 0, capi.VS_MEM_HOST), "vs_clip_warp")
        return out

    def warp_device(self, slots, transforms: np.ndarray, d_out: int, out_frame_stride: int,
                    mode=capi.VS_WARP_CV_EXACT_BILINEAR, border=capi.VS_BORDER_CONSTANT0, crop: int = 0):
        s = np.asarray(list(slots), np.int32)
        T = np.ascontiguousarray(transforms, np.float64).reshape(len(s), 4)
        self._chk(self.lib.vs_clip_warp(self.handle, capi.ptr(s), len(s), capi.ptr(T), mode, border, crop, C.c_void_p(d_out),
                                        out_frame_stride, capi.VS_MEM_DEVICE), "vs_clip_warp")

    # ---- inspection taps
    def get_bgr(self, slot: int) -> np.ndarray:
        out = np.empty((self.height * 3 // 2, self.width) if self.nv12 else (self.height, self.width, 3), np.uint8)
        self._chk(self.lib.vs_clip_get_bgr(self.handle, slot, capi.ptr(out)), "vs_clip_get_bgr")
        return out

    def get_gray(self, slot: int, level: int) -> np.ndarray:
        li = self.level_info(level)
        out = np.empty((li["h"], li["w"]), np.uint8)
        self._chk(self.lib.vs_clip_get_gray(self.handle, slot, level, capi.ptr(out)), "vs_clip_get_gray")
        return out

    def get_keypoints(self, slot: int, level: int, axis: int) -> np.ndarray:
        li = self.level_info(level)
        out = np.empty((2, li["th"], li["tw"]), np.uint16)
        self._chk(self.lib.vs_clip_get_keypoints(self.handle, slot, level, axis, capi.ptr(out)), "vs_clip_get_keypoints")
        return out

    def get_jacobians(self, slot: int, level: int, axis: int) -> np.ndarray:
        li = self.level_info(level)
        out = np.empty((4, li["th"], li["tw"]), np.float32)
        self._chk(self.lib.vs_clip_get_jacobians(self.handle, slot, level, axis, capi.ptr(out)), "vs_clip_get_jacobians")
        return out

    def get_warpdiff(self, pair: int, level: int, axis: int) -> np.ndarray:
        li = self.level_info(level)
        out = np.empty((li["th"], li["tw"]), np.uint16)
        self._chk(self.lib.vs_clip_get_warpdiff(self.handle, pair, level, axis, capi.ptr(out)), "vs_clip_get_warpdiff")
        return out

    def get_phase(self, pair: int) -> np.ndarray:
        """cv::phaseCorrelate result (shift x, shift y, response) of a pair of the last align call (phase_correlate on)."""
        out = np.zeros(3, np.float64)
        self._chk(self.lib.vs_clip_get_phase(self.handle, pair, capi.ptr(out)), "vs_clip_get_phase")
        return out

    def get_selected(self, pair: int, level: int, axis: int) -> np.ndarray:
        li = self.level_info(level)
        out = np.empty(li["th"] * li["tw"], np.uint32)
        k = C.c_int()
        self._chk(self.lib.vs_clip_get_selected(self.handle, pair, level, axis, capi.ptr(out), C.byref(k)), "vs_clip_get_selected")
        return out[:k.value].copy()
