#!/usr/bin/env python
"""One long 1080p video stabilized by frame-chunk partition over the GPUs of the box, in one process
(vstab::MultiGpuStabilizer): host frames in, host frames out, wall clock (pinned host buffers)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from video_stabilizer_b200 import _capi as capi, host, synth  # noqa: E402
from video_stabilizer_b200.imgproc import Context  # noqa: E402

W, H, N = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 600
ndev = capi.load().vs_device_count()
ctx = Context(0)
frames, _ = synth.make_clip_gpu(ctx, W, H, N, 5, chunk=50)
ctx.close()
p = host.stab_params_default()
p.crop_pixels = 0
ref = None
# pinned host buffers (as a capture / decode pipeline would provide): the copies then run at the PCIe rate
import torch  # noqa: E402
pin_in = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
pin_in.numpy()[...] = frames
frames = pin_in.numpy()
pin_out = torch.empty((N, H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()
for g in sorted(set([1, min(2, ndev), min(4, ndev), ndev])):
    mg = host.MultiGpuStabilizer(list(range(g)), W, H, N, p)
    mg.stabilize(frames[:64])                      # warm-up
    t0 = time.perf_counter()
    out, meas, ok = mg.stabilize(frames, pin_out)
    dt = time.perf_counter() - t0
    same = None if ref is None else bool(np.array_equal(out, ref))
    if ref is None:
        ref = out.copy()             # `out` is a view of the pinned buffer the next run overwrites
    print(json.dumps({"gpus": g, "frames": N, "seconds": dt, "frames_per_s": N / dt, "pairs_converged": int(ok.sum()),
                      "identical_to_1_gpu": same}), flush=True)
    mg.close()
