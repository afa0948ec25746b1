#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that are not the bench.py headline.

  python tools/config_bench.py pairs4k    # configs[2]: 3840x2160 batched frame-pair alignment, device-resident
  python tools/config_bench.py clips720   # configs[3]: 64 independent 720p clips aligned and warped in shared launches
One JSON object per measurement; CUDA events on the launching stream, 3 warm-ups.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _setup():
    import torch
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.imgproc import Context
    st = torch.cuda.Stream()
    ctx = Context(0, stream=st.cuda_stream)
    return torch, capi, ctx, st


def _time(torch, st, fn, iters):
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


def pairs4k(args):
    torch, capi, ctx, st = _setup()
    from video_stabilizer_b200 import synth
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    W, H, n = 3840, 2160, args.frames
    frames, _ = synth.make_clip_gpu(ctx, W, H, n, 31, chunk=16)
    clip = Clip(W, H, n, ctx=ctx)
    clip.upload(0, frames)
    pairs, keys = pairs_for_frames(0, n)
    d_T = torch.empty((n - 1, 4), dtype=torch.float64, device="cuda")
    d_st = torch.empty(n - 1, dtype=torch.int32, device="cuda")

    def step():
        clip.build_pyramids(0, n)
        clip.build_keyframes(keys)
        clip.align_device(pairs, n - 1, d_T.data_ptr(), d_st.data_ptr())
    ms = _time(torch, st, step, args.iters)
    ok = int(d_st.sum().item())
    print(json.dumps({"config": "4K batched frame-pair alignment (BGR->gray pyramid, keyframe features, sparse LK solve), device-resident",
                      "size": "3840x2160", "pairs_per_launch": n - 1, "ms_per_batch": ms, "align_ms_per_pair": ms / (n - 1),
                      "pairs_per_s": (n - 1) / (ms / 1e3), "pairs_converged": ok}), flush=True)


def clips720(args):
    torch, capi, ctx, st = _setup()
    from video_stabilizer_b200 import host, synth
    from video_stabilizer_b200.clip import Clip
    W, H, nclips, nf = 1280, 720, args.clips, args.clip_frames
    base, _ = synth.make_clip_gpu(ctx, W, H, nf, 41, chunk=nf)
    clip = Clip(W, H, nclips * nf, max_pairs=nclips * (nf - 1), ctx=ctx)
    for c in range(nclips):                       # distinct content per clip: shifted copies
        clip.upload(c * nf, np.roll(base, 3 * c, axis=2))
    pairs = (capi.VsPair * (nclips * (nf - 1)))()
    keys = []
    i = 0
    for c in range(nclips):
        for f in range(1, nf):
            s, p = c * nf + f, c * nf + f - 1
            if f % 2 == 1:
                pairs[i] = capi.VsPair(p, s, 0)
                keys.append(s)
            else:
                pairs[i] = capi.VsPair(s, p, 1)
            i += 1
    lag = 10
    n_out = nclips * (nf - lag)
    out = torch.empty((n_out, H, W, 3), dtype=torch.uint8, device="cuda")
    slots = [c * nf + f for c in range(nclips) for f in range(nf - lag)]

    def step():
        clip.build_pyramids(0, nclips * nf)
        clip.build_keyframes(keys)
        T, status, _ = clip.align(pairs)          # 40 B per pair to the host
        corr = np.zeros((n_out, 4))
        k = 0
        j = 0
        for c in range(nclips):                   # one sequential trajectory per clip
            traj = host.StabilizerTrajectory()
            for f in range(nf):
                if f == 0:
                    due, cr = traj.push(np.zeros(4), False, W, H)
                else:
                    due, cr = traj.push(T[j], bool(status[j]), W, H)
                    j += 1
                if due:
                    corr[k] = cr
                    k += 1
        clip.warp_device(slots, corr, out.data_ptr(), W * H * 3)
        return status
    ms = _time(torch, st, step, args.iters)
    status = step()
    rec = {"config": "%d independent 720p clips x %d frames aligned and warped in shared launches, device-resident" % (nclips, nf),
           "size": "1280x720", "frames_per_step": nclips * nf, "ms_per_step": ms,
           "frames_per_s": nclips * nf / (ms / 1e3), "pairs_converged": int(status.sum()), "pairs": len(status)}
    if not getattr(args, "quiet", False):
        print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["pairs4k", "clips720", "all"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--frames", type=int, default=65)
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--clip-frames", type=int, default=32)
    a = ap.parse_args()
    if a.what in ("pairs4k", "all"):
        pairs4k(a)
    if a.what in ("clips720", "all"):
        clips720(a)
