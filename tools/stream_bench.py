#!/usr/bin/env python
"""Throughput of the drop-in per-frame API: T host threads, one VideoStabilizer each (the reference's
own scale-out recipe, grid_search_align.cpp:159-210), all on one GPU; frames come from and return to
ordinary (pageable) host memory exactly as a cv::Mat caller would hand them over."""
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from video_stabilizer_b200 import host, synth  # noqa: E402
from video_stabilizer_b200.imgproc import Context  # noqa: E402

W, H, N = 1920, 1080, 64
ctx = Context(0)
frames, _ = synth.make_clip_gpu(ctx, W, H, N, 9, chunk=32)
ctx.close()
p = host.stab_params_default()
p.crop_pixels = 0
for T in (1, 2, 4, 8, 16):
    stabs = [host.VideoStabilizer(p, 0) for _ in range(T)]
    for s in stabs:
        s.processFrame(frames[0])

    def work(s):
        for f in frames[1:]:
            s.processFrame(f)
    ths = [threading.Thread(target=work, args=(s,)) for s in stabs]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    print(json.dumps({"threads": T, "frames_per_s": T * (N - 1) / dt, "ms_per_frame_per_stream": 1e3 * dt / (N - 1)}), flush=True)
    for s in stabs:
        s.close()
