#!/usr/bin/env python
"""Per-kernel micro-benchmarks on device-resident data (CUDA events, footprint >> L2).

  python tools/kernel_bench.py warp      # BASELINE.json configs[4]: BGR warp, 3 modes x 1080p..8K, GB/s vs HBM roofline
  python tools/kernel_bench.py pyramid   # BGR->gray + pyr_down chain at 1080p / 4K
  python tools/kernel_bench.py all
Prints one JSON object per measurement.  Timing: 3 warm-up launches, then `--iters` launches
between two CUDA events on the launching stream; batches are sized to >= 512 MB per launch.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


_STREAM = None


def stream(torch):
    """A real (non-default) torch stream: the library borrows it, so the CUDA events below
    are recorded on the stream the kernels are launched on."""
    global _STREAM
    if _STREAM is None:
        _STREAM = torch.cuda.Stream()
    return _STREAM


def time_launches(torch, fn, iters):
    st = stream(torch)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_warp(args):
    """Returns the list of records (and prints each as a JSON line unless args.quiet)."""
    import torch
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.imgproc import Context
    ctx = Context(0, stream=stream(torch).cuda_stream)
    lib = ctx.lib
    sizes = [(1920, 1080), (2560, 1440), (3840, 2160), (5120, 2880), (7680, 4320)]
    if args.size:
        sizes = [tuple(int(v) for v in args.size.split("x"))]
    modes = [(0, "cv_exact_bilinear"), (1, "float_bilinear"), (2, "lanczos2")]
    if args.mode is not None:
        modes = [m for m in modes if m[0] == args.mode]
    records = []
    T = np.array([0.0013, -0.0021, 6.37, -3.81])
    if args.transform:
        T = np.array([float(v) for v in args.transform.split(",")])   # e.g. 0,0,0,0 (a camera standing still) or 0,0,5,-3
    for (w, h) in sizes:
        batch = max(1, int(np.ceil(512e6 / (6 * w * h))))
        pitch = (3 * w + 127) // 128 * 128
        src = torch.randint(0, 256, (batch, h, pitch), dtype=torch.uint8, device="cuda")
        dst = torch.empty((batch, h, w * 3), dtype=torch.uint8, device="cuda")
        cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
        M = np.tile(np.array([1 + T[0], -T[1], T[2] - T[0] * cx + T[1] * cy, T[1], 1 + T[0], T[3] - T[1] * cx - T[0] * cy]), (batch, 1))
        s = capi.VsImg(src.data_ptr(), w, h, pitch, batch, pitch * h)
        d = capi.VsImg(dst.data_ptr(), w, h, w * 3, batch, w * h * 3)
        for mode, name in modes:
            def fn():
                capi.check(ctx.handle, lib.vs_bgr_warp_u8(ctx.handle, C.byref(s), capi.ptr(M), C.byref(d), 0, 0, mode,
                                                          capi.VS_BORDER_CONSTANT0, capi.VS_MEM_DEVICE), "vs_bgr_warp_u8")
            ms = time_launches(torch, fn, args.iters)
            gbs = 6.0 * w * h * batch / (ms / 1e3) / 1e9
            rec = {"kernel": "bgr_warp", "mode": name, "size": "%dx%d" % (w, h), "batch": batch, "transform": T.tolist(),
                   "bytes_per_launch": 6 * w * h * batch, "ms_per_launch": ms, "algorithmic_gbs": gbs,
                   "frac_of_hbm_peak": gbs / peak(), "frames_per_s": batch / (ms / 1e3)}
            records.append(rec)
            if not getattr(args, "quiet", False):
                print(json.dumps(rec), flush=True)
        del src, dst
    return records


def bench_pyramid(args):
    import torch
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.imgproc import Context
    ctx = Context(0, stream=stream(torch).cuda_stream)
    lib = ctx.lib
    sizes = [(1920, 1080), (3840, 2160)]
    if args.size:
        sizes = [tuple(int(v) for v in args.size.split("x"))]
    for (w, h) in sizes:
        batch = max(1, int(np.ceil(768e6 / (4 * w * h))))
        bgr = torch.randint(0, 256, (batch, h, w * 3), dtype=torch.uint8, device="cuda")
        levels = []
        ww, hh = w, h
        while ww >= 40 and hh >= 40 or not levels:
            levels.append(torch.empty((batch, hh, ww), dtype=torch.uint8, device="cuda"))
            ww //= 2
            hh //= 2
            if len(levels) > 7:
                break
        bi = capi.VsImg(bgr.data_ptr(), w, h, w * 3, batch, w * h * 3)
        imgs = [capi.VsImg(t.data_ptr(), t.shape[2], t.shape[1], t.shape[2], batch, t.shape[1] * t.shape[2]) for t in levels]

        def gray():
            capi.check(ctx.handle, lib.vs_bgr2gray_u8(ctx.handle, C.byref(bi), C.byref(imgs[0]), capi.VS_MEM_DEVICE), "bgr2gray")
        ms = time_launches(torch, gray, args.iters)
        gbs = 4.0 * w * h * batch / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "bgr2gray", "size": "%dx%d" % (w, h), "batch": batch, "ms_per_launch": ms,
                          "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / peak()}), flush=True)
        for l in range(min(3, len(levels) - 1)):
            def down(l=l):
                capi.check(ctx.handle, lib.vs_pyr_down_u8(ctx.handle, C.byref(imgs[l]), C.byref(imgs[l + 1]), capi.VS_MEM_DEVICE), "pyr_down")
            ms = time_launches(torch, down, args.iters)
            b = (levels[l].shape[1] * levels[l].shape[2] + levels[l + 1].shape[1] * levels[l + 1].shape[2]) * batch
            gbs = b / (ms / 1e3) / 1e9
            print(json.dumps({"kernel": "pyr_down", "level": "%d->%d" % (l, l + 1), "size": "%dx%d" % (levels[l].shape[2], levels[l].shape[1]),
                              "batch": batch, "ms_per_launch": ms, "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / peak()}), flush=True)
        del bgr, levels


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["warp", "pyramid", "all"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--size", default=None)
    ap.add_argument("--mode", type=int, default=None)
    ap.add_argument("--transform", default=None, help="A,B,TX,TY of the similarity transform of the warp sweep")
    a = ap.parse_args()
    if a.what in ("warp", "all"):
        bench_warp(a)
    if a.what in ("pyramid", "all"):
        bench_pyramid(a)
