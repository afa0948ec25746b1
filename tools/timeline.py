#!/usr/bin/env python
"""Timeline of one device-resident video through PartitionedStabilizer (CUDA events around every launch, all streams):
where the GPU idles and which kernels overlap.   python tools/timeline.py [sub_frames block lanes [width height frames]]"""
import ctypes as C
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from video_stabilizer_b200 import _capi as capi, host, synth  # noqa: E402
from video_stabilizer_b200.imgproc import Context  # noqa: E402

W, H, F = 1920, 1080, 300
if len(sys.argv) > 6:
    W, H, F = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 100
block = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = host.stab_params_default()
p.crop_pixels = 0
ps = host.PartitionedStabilizer(0, 1, W, H, F, sub, block, p, "", True, device=0, host_threads=8, lanes=lanes)
ctx = Context(0)
frames = torch.empty((F, H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()
synth.make_clip_gpu(ctx, W, H, F, 100, out=frames, chunk=50 if W <= 1920 else 8)
ctx.close()
ps.upload_resident(frames.ctypes.data, W * 3, W * H * 3)
out = torch.empty((ps.outputs, ps.out_h, ps.out_w, 3), dtype=torch.uint8, device="cuda")
lib = capi.load()
for _ in range(3):
    ps.stabilize_ptr(None, 0, 0, out.data_ptr(), capi.VS_MEM_DEVICE)
ps.synchronize()
# plain timing first (no per-launch events)
st = torch.cuda.Stream()
ps.set_stream(st.cuda_stream)
for _ in range(2):
    ps.stabilize_ptr(None, 0, 0, out.data_ptr(), capi.VS_MEM_DEVICE)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(20):
    ps.stabilize_ptr(None, 0, 0, out.data_ptr(), capi.VS_MEM_DEVICE)
e1.record(st)
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 20
lib.vs_ctx_profile_enable(ps.ctx_handle, 1)
lib.vs_ctx_profile_reset(ps.ctx_handle)
ps.stabilize_ptr(None, 0, 0, out.data_ptr(), capi.VS_MEM_DEVICE)
ps.synchronize()
buf = np.zeros((4096, 3))
n = lib.vs_ctx_profile_timeline(ps.ctx_handle, buf.ctypes.data_as(C.c_void_p), 4096)
rows = [(lib.vs_kernel_name(int(buf[i, 0])).decode(), round(float(buf[i, 1]), 3), round(float(buf[i, 2]), 3)) for i in range(n)]
for r in sorted(rows, key=lambda r: r[1]):
    print("%-22s %8.3f -> %8.3f  (%.3f ms)" % (r[0], r[1], r[2], r[2] - r[1]))
print(json.dumps({"sub": sub, "block": block, "lanes": lanes, "ms_per_video": plain_ms, "span_ms": max(r[2] for r in rows), "launches": n}))
