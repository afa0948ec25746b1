#!/usr/bin/env python
"""Cost of the phase-correlation initialiser (VideoAlignerParams::phase_correlate) on device-resident frames:
one vs_clip_align call over n-1 pairs with and without the seed, CUDA events on the launching stream."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import torch
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200 import synth
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    from video_stabilizer_b200.imgproc import Context
    st = torch.cuda.Stream()
    ctx = Context(0, stream=st.cuda_stream)
    frames, _ = synth.make_clip_gpu(ctx, a.width, a.height, a.frames, 5)
    for pc in (0, 1):
        p = capi.VsAlignParams()
        capi.load().vs_align_params_default(p)
        p.phase_correlate = pc
        clip = Clip(a.width, a.height, a.frames, params=p, ctx=ctx)
        clip.upload(0, frames)
        pairs, keys = pairs_for_frames(0, a.frames)
        times = []
        for it in range(a.iters + 2):
            clip.build_pyramids(0, a.frames)      # invalidates the cached spectra: every call pays the forward transforms
            clip.build_keyframes(keys)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            T, status, iters = clip.align(pairs)
            e1.record(st)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[2:]))
        print(json.dumps({"phase_correlate": pc, "size": "%dx%d" % (a.width, a.height), "pairs": len(pairs), "align_ms": ms,
                          "ms_per_pair": ms / len(pairs), "converged": int(status.sum())}), flush=True)
        clip.close()


if __name__ == "__main__":
    main()
