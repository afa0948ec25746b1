#!/usr/bin/env python
"""Several independent clips stabilized concurrently on ONE GPU: T host threads, one ClipStabilizer (context, stream,
ring) each — the reference's own scale-out recipe (grid_search_align.cpp:159-210: one instance per thread) on a device.
The latency-bound solver of one clip then runs beside the bandwidth-bound stages of another.
Device-resident frames; timed with CUDA events: e0 on a master stream every worker stream waits for, e1 after the
master stream has waited for every worker's last kernel."""
import argparse
import json
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--steps", type=int, default=40, help="clips per worker inside the timed region")
    ap.add_argument("--threads", default="1,2,3")
    ap.add_argument("--lanes", type=int, default=3)
    a = ap.parse_args()
    import torch
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200 import host, synth
    from video_stabilizer_b200.imgproc import Context
    W, H, F = a.width, a.height, a.frames
    pinned = torch.empty((F, H, W, 3), dtype=torch.uint8, pin_memory=True)
    frames = pinned.numpy()
    g = Context(0)
    synth.make_clip_gpu(g, W, H, F, seed=100, out=frames, chunk=50)
    g.close()
    p = host.stab_params_default()
    p.crop_pixels = 0
    n_out = F - p.lag
    for T in [int(v) for v in a.threads.split(",")]:
        workers = []
        for t in range(T):
            cs = host.ClipStabilizer(W, H, F, p, device=0)
            st = torch.cuda.Stream()
            cs.set_stream(st.cuda_stream)
            cs.set_solver_lanes(a.lanes)
            out = torch.empty((n_out, cs.out_h, cs.out_w, 3), dtype=torch.uint8, device="cuda")
            cs.upload_only(0, frames.ctypes.data, F, W * 3, W * H * 3, capi.VS_MEM_HOST)
            cs.synchronize()
            workers.append((cs, st, out))

        def run(i, steps):
            cs, st, out = workers[i]
            for _ in range(steps):
                cs.reset()
                k = cs.feed_resident(F, out.data_ptr(), capi.VS_MEM_DEVICE)
                assert k == n_out

        def all_workers(steps):
            th = [threading.Thread(target=run, args=(i, steps)) for i in range(T)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        all_workers(3)
        torch.cuda.synchronize()
        master = torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(master)
        for (_, st, _) in workers:
            st.wait_event(e0)
        all_workers(a.steps)
        for (_, st, _) in workers:
            ev = torch.cuda.Event()
            ev.record(st)
            master.wait_event(ev)
        e1.record(master)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"concurrent_clips": T, "size": "%dx%d" % (W, H), "frames_per_clip": F, "clips": T * a.steps,
                          "ms_per_clip_amortised": ms / (T * a.steps), "frames_per_s": T * a.steps * F / (ms / 1e3),
                          "solver_lanes": a.lanes}), flush=True)
        for (cs, _, _) in workers:
            cs.close() if hasattr(cs, "close") else None
        del workers


if __name__ == "__main__":
    main()
