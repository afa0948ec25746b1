#!/usr/bin/env python
"""bench.py — stabilized frames/s of the alignment-and-warp hot path on B200.

Workload (BASELINE.json configs[1], "video_test"): ONE 1920x1080 synthetic, procedurally jittered BGR video of
300 frames PER GPU through the full pipeline — fused BGR->gray + pyramid ingest, keyframe features, batched sparse
Lucas-Kanade solve, the host L1 smoother, cv-exact BGR warp — with VideoStabilizerParams as video_test.cpp:53-54 sets
them (crop_pixels = 0).  With N GPUs the video is N x 300 frames long and is partitioned by frame chunk (north_star):
rank r owns the contiguous frames [300 r, 300 r + 300) plus one halo frame, the 40 bytes per pair go to a table in host
shared memory, every rank runs its share of the smoothing and the (cheap, sequential) accumulate chain and warps its own
frames.  No collective on the data path (NCCL only carries the barrier and the MAX of the step time): weak scaling,
value = all frames / max step time over ranks.  One step = `--passes` videos back to back (so that the timed region of
the default --steps is about a second).

  value  frames/s with the frames already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e    frames/s through PartitionedStabilizer::stabilize() with HOST buffers: pinned H2D of every frame and D2H of
         every stabilized frame inside the timed region; here the video is interleaved over the ranks in 32-frame
         sub-chunks so that every GPU uploads, computes and downloads all the time.  `host_ceiling` is the box's own
         pinned-copy rate with all N GPUs copying both ways at once, measured in the same run.
  roofline      dominant kernel of the timed region: algorithmic bytes / its CUDA-event time
  cpu_baseline  the reference's own host sources (oracle/_ref) or the oracle port, timed on this box's cores on a
                bounded sample of the same workload (rank 0, N=1)
  extra         4K (120 frames per GPU) through the same partitioned pipeline at every N; at N=1 also configs[3]
                (64 x 720p clips in shared launches) and the configs[4] BGR warp sweep

`--impl reference` times the reference CPU implementation instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "stabilized_frames_per_sec_1080p"
UNIT = "frames/s"
CLIP_SEED = 100


def level_table(w, h, min_w=20, min_h=20):
    """Pyramid levels and tile grids exactly as the library lays them out
    (alignment.cpp:164-169, imgproc.cpp:151-162)."""
    levels = []
    n, ww, hh = 0, w, h
    while True:
        n += 1
        ww //= 2
        hh //= 2
        if not (ww >= min_w and hh >= min_h):
            break
    ww, hh = w, h
    for l in range(n):
        if l > 0:
            ww //= 2
            hh //= 2
        tile = 2
        for i in range(4, 21, 2):
            if (ww // i) * (hh // i) < 1000:
                break
            tile = i
        levels.append(dict(w=ww, h=hh, tile=tile, tiles=(ww // tile) * (hh // tile)))
    return levels


def algorithmic_bytes(w, h, crop, n_frames, n_keyframes, n_pairs, n_warped):
    """Compulsory HBM traffic per video of each kernel (each input byte read once, each output byte written once),
    BASELINE.md section 4 / DESIGN.md.  pyr_down: the levels below level 1 when the ingest is fused (16-aligned widths)."""
    lv = level_table(w, h)
    px = [L["w"] * L["h"] for L in lv]
    tiles = sum(L["tiles"] for L in lv)
    ow, oh = w - 2 * crop, h - 2 * crop
    fused = w % 16 == 0 and len(px) > 1
    return {
        "bgr2gray": 4 * w * h * n_frames,
        "ingest_bgr_gray_l1": (3 * w * h + px[0] + px[1]) * n_frames,
        "pyr_down": sum(px[i] + px[i + 1] for i in range(1 if fused else 0, len(px) - 1)) * n_frames,
        "keyframe_features": (sum(px) + 40 * tiles) * n_keyframes,
        # both pyramids of a pair read once + keypoints/Jacobians of the keyframe + 40 B out
        "solve_pairs": (2 * sum(px) + 2 * tiles * 20 + 40) * n_pairs,
        "bgr_warp": (3 * w * h + 3 * ow * oh) * n_warped,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, world=None):
    """The workload both arms are quoted on (the CPU arm times a bounded sample of it and says which)."""
    world = world or args.gpus
    return {
        "workload": "video_test: one %dx%d synthetic procedurally-jittered BGR video of %d frames per GPU, full pipeline "
                    "(BGR->gray pyramid, keyframe features, sparse LK solve, host L1 smoother, cv-exact BGR warp), "
                    "VideoStabilizerParams of video_test.cpp:53-54 (crop_pixels=%d, lag=10)" % (args.width, args.height, args.frames, args.crop),
        "frames_per_video_per_gpu": args.frames, "width": args.width, "height": args.height,
        "partition": "frame-chunk of one video: rank r owns frames [%d r, %d r + %d) + 1 halo frame; 40 B per pair to a table "
                     "in host shared memory; no collective on the data path" % (args.frames, args.frames, args.frames),
        "l2": "inputs larger than L2: %.2f GB of frames read per video per GPU vs 126 MB L2" % (args.frames * args.width * args.height * 3 / 1e9),
    }


# ----------------------------------------------------------------------------- CPU reference
def _cpu_clip(width, height, n, seed):
    """The first n frames of the very clip rank 0 of the GPU arm stabilizes (same canvas, same poses, same bytes: the
    renderer is the cv-exact integer warp), rendered on the host by the oracle's C restatement of cv::warpAffine."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import binding as ob
    from video_stabilizer_b200 import synth
    ob.load(fast=True)
    canvas = synth.make_canvas(width, height, 1000 + seed)
    poses = synth.jitter_path(max(n, 1), 1001 + seed)
    frames = np.empty((n, height, width, 3), np.uint8)

    def one(t):
        ob.warp_bgr_matrix(canvas, synth.forward_matrix_for_pose(poses[t], width, height), width, height, fast=True, out=frames[t])
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(n)))
    return frames


def cpu_reference(clip, threads, crop):
    """Frames/s of the reference CPU implementation: `threads` workers, one VideoStabilizer each (the reference's own
    scale-out recipe, grid_search_align.cpp:105-118,159-210), every worker stabilizing `clip` from its first frame.
    Returns (frames_per_s, kind, seconds)."""
    from oracle import binding as ob
    use_ref = ob.ref_available()
    if use_ref:
        try:
            ob.load_ref(fast=True)
        except Exception:
            use_ref = False
    kind = "reference" if use_ref else "port"
    p = ob.stab_params_default()
    p.crop_pixels = crop

    def work(_):
        st = ob.RefStabilizer(p, fast=True) if use_ref else ob.Stabilizer(p, fast=True)
        for f in clip:      # read-only input shared by the workers
            st.process(f)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return threads * len(clip) / dt, kind, dt


def cpu_sample_frames(width, height, threads, crop, target_seconds, max_frames):
    """How many leading frames of the clip make one sample last about target_seconds (calibrated on 12 frames)."""
    clip = _cpu_clip(width, height, 12, CLIP_SEED)
    fps, _, _ = cpu_reference(clip, threads, crop)
    per_thread = fps / threads
    return int(max(12, min(max_frames, round(target_seconds * per_thread))))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.cpu_threads or cores))
    # each step is a bounded sample of the workload, sized so that the whole --steps/--warmup run stays around 2-3 minutes
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    n = cpu_sample_frames(args.width, args.height, threads, args.crop, budget, args.frames)
    clip = _cpu_clip(args.width, args.height, n, CLIP_SEED)
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        fps, kind, dt = cpu_reference(clip, threads, args.crop)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = threads * n / (ms / 1000.0)
    sample = ("%d threads, one VideoStabilizer each, every thread stabilizes the first %d frames of the %d-frame %dx%d video of the "
              "GPU arm (same bytes) per step" % (threads, n, args.frames, args.width, args.height))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- GPU arm
class Ranks:
    """torch.distributed plumbing: barrier and MAX only."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the one JSON line
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        if not self.dist:
            return float(v)
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def exchange_name(tag):
    """POSIX shared-memory name of one video's table: the same on every rank of this job, distinct between jobs."""
    job = os.environ.get("TORCHELASTIC_RUN_ID", "") + "_" + os.environ.get("MASTER_PORT", "0")
    return "/vstab_%s_%s_%d" % (tag, "".join(c for c in job if c.isalnum() or c == "_")[-24:], os.getuid())


class PartitionedVideo:
    """One rank's share of a synthetic N x F-frame video in a given partition, with its pinned host frames."""

    def __init__(self, R, args, W, H, F, sub, block, resident, tag, seed=CLIP_SEED, share=None, nv12=False):
        import torch
        from video_stabilizer_b200 import host, synth
        from video_stabilizer_b200.imgproc import Context
        self.R, self.W, self.H, self.F = R, W, H, F
        self.total = R.world * F
        p = host.stab_params_default()
        p.crop_pixels = args.crop
        self.params = p
        threads = max(1, min(8, (os.cpu_count() or 1) // max(1, R.world * max(1, args.inflight))))
        name = exchange_name(tag) if R.world > 1 else ""
        # rank 0 creates the shared table; the others attach once it exists
        if R.rank == 0:
            self.ps = host.PartitionedStabilizer(R.rank, R.world, W, H, self.total, sub, block, p, name, resident,
                                                 device=R.local, host_threads=threads, nv12=nv12)
        R.barrier()
        if R.rank != 0:
            self.ps = host.PartitionedStabilizer(R.rank, R.world, W, H, self.total, sub, block, p, name, resident,
                                                 device=R.local, host_threads=threads, nv12=nv12)
        R.barrier()
        self.host_threads = threads
        ps = self.ps
        n_local = len(ps.local_frames)
        if share is not None:      # a second instance over the same video: the same pinned frames
            self.pinned, self.frames = share.pinned, share.frames
        else:
            self.pinned = torch.empty((n_local, H * 3 // 2, W) if nv12 else (n_local, H, W, 3), dtype=torch.uint8, pin_memory=True)
            self.frames = self.pinned.numpy()
            ctx = Context(R.local)
            canvas = synth.make_canvas(W, H, 1000 + seed)
            poses = synth.jitter_path(self.total, 1001 + seed)
            if nv12:
                # the same video as NV12 frames: Y = the gray value (the canvas is gray replicated, so BGR2GRAY returns it),
                # U / V derived from it (any chroma does: it is only warped)
                bgr = np.empty((min(n_local, 32), H, W, 3), np.uint8)
                for i0 in range(0, n_local, len(bgr)):
                    k = min(len(bgr), n_local - i0)
                    synth.render_frames_gpu(ctx, canvas, poses, ps.local_frames[i0:i0 + k], W, H, bgr[:k], chunk=32 if W <= 1920 else 8)
                    self.frames[i0:i0 + k, :H] = bgr[:k, :, :, 0]
                    uv = self.frames[i0:i0 + k, H:].reshape(k, H // 2, W // 2, 2)
                    uv[..., 0] = bgr[:k, ::2, ::2, 0] // 2 + 64
                    uv[..., 1] = 191 - bgr[:k, 1::2, 1::2, 0] // 2
                del bgr
            else:
                synth.render_frames_gpu(ctx, canvas, poses, ps.local_frames, W, H, self.frames, chunk=32 if W <= 1920 else 8)
            ctx.close()
        self.frame_bytes = W * H * 3 // 2 if nv12 else W * H * 3
        self.row_stride = W if nv12 else W * 3
        self.stream = torch.cuda.Stream()
        assert self.stream.cuda_stream != 0
        # a real (non-default) torch stream, borrowed by the library: the CUDA events of the timed region are recorded on
        # the stream the kernels are launched on
        ps.set_stream(self.stream.cuda_stream)
        self.h2d_bytes = n_local * self.frame_bytes
        self.d2h_bytes = ps.outputs * ps.out_frame_bytes


def timed(R, stream, fn, steps, warmup, launches_of=None):
    torch = R.torch
    for _ in range(warmup):
        fn()
    R.barrier()
    n0 = launches_of() if launches_of else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    R.barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / steps
    timed.launches = (launches_of() - n0) if launches_of else 0
    return R.reduce(ms, "max"), t0, t1


def timed_inflight(R, streams, fns, steps, warmup, launches_of=None):
    """`steps` steps over len(fns) independent instances, one host thread each (step s runs on instance s % n): the next
    video's uploads and first stages fill the drain of the previous one.  CUDA events on every instance's stream; the
    time is from the earliest start event to the latest end event."""
    torch = R.torch
    n = len(fns)
    if n == 1:
        return timed(R, streams[0], fns[0], steps, warmup, launches_of)

    def run(count):
        errs = []

        def work(i):
            try:
                torch.cuda.set_device(R.local)
                for _ in range(i, count, n):
                    fns[i]()
            except BaseException as e:   # noqa: BLE001 - re-raised on the main thread
                errs.append(e)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    run(max(warmup, 1) * n)
    R.barrier()
    n0 = launches_of() if launches_of else 0
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    t0 = time.time()
    for e, st in zip(e0, streams):
        e.record(st)
    run(steps)
    for e, st in zip(e1, streams):
        e.record(st)
    R.barrier()
    t1 = time.time()
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b) for a in e0 for b in e1) / steps
    timed.launches = (launches_of() - n0) if launches_of else 0
    return R.reduce(ms, "max"), t0, t1


def host_copy_ceiling(R, seconds=0.4, mbytes=256):
    """Pinned host <-> device copy rate of THIS box with every rank copying both ways at once: the denominator of the
    end-to-end number.  Returns GB/s summed over ranks: (h2d, d2h) together, and h2d alone."""
    torch = R.torch
    n = mbytes << 20
    hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(n, dtype=torch.uint8, device="cuda")
    dout = torch.empty(n, dtype=torch.uint8, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(both):
        reps = 0
        for phase in (0, 1):
            R.barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            k = 2 if phase == 0 else max(2, reps)
            e[0].record(s_up)
            e[2].record(s_down)
            for _ in range(k):
                with torch.cuda.stream(s_up):
                    din.copy_(hin, non_blocking=True)
                if both:
                    with torch.cuda.stream(s_down):
                        hout.copy_(dout, non_blocking=True)
            e[1].record(s_up)
            e[3].record(s_down)
            R.barrier()
            ms_up, ms_down = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
            if phase == 0:      # the same repeat count on every rank: they must all be copying for the whole measurement
                reps = int(R.reduce(max(2, seconds * 1e3 / max(ms_up / k, 1e-3)), "max"))
        up = n * k / (ms_up / 1e3) / 1e9
        down = n * k / (ms_down / 1e3) / 1e9 if both else 0.0
        return up, down

    up1, _ = run(False)
    up2, down2 = run(True)
    return {"h2d_alone_gbs": R.reduce(up1, "sum"), "h2d_duplex_gbs": R.reduce(up2, "sum"), "d2h_duplex_gbs": R.reduce(down2, "sum")}


def measure_partitioned(R, args, W, H, F, steps, warmup, passes, e2e_steps, e2e_sub, profile, tag, seed=CLIP_SEED, nv12=False):
    """value (resident, contiguous chunk per rank) and e2e (host-streamed, interleaved sub-chunks) of one video size."""
    import torch
    from video_stabilizer_b200 import _capi as capi
    lib = capi.load()
    res = {}
    # ---- device-resident: each rank's contiguous chunk as three sub-chunks (the solve of one beside the pyramids of the next)
    nsub = max(1, getattr(args, "sub_chunks", 3))
    sub = F // nsub if (F // nsub) % 2 == 0 and F % nsub == 0 else F
    block = F // sub
    if sub % 2 or F % sub or sub < 10:
        raise SystemExit("bench.py: --frames must be even (and at least 10)")
    # `inflight` videos are in flight at a time (independent instances, one host thread each): the pyramids of the next
    # video run beside the last solves and warps of the previous one
    inflight = max(1, args.inflight)
    pvs = []
    for i in range(inflight):
        pvs.append(PartitionedVideo(R, args, W, H, F, sub, block, True, tag + "r%d" % i, seed, share=pvs[0] if i else None, nv12=nv12))
    pv = pvs[0]
    ps = pv.ps
    outs_dev = []
    for q in pvs:
        q.ps.upload_resident(q.frames.ctypes.data, q.row_stride, q.frame_bytes)
        q.ps.synchronize()
        outs_dev.append(torch.empty((max(ps.outputs, 1),) + ps.out_shape, dtype=torch.uint8, device="cuda"))

    def make_step(q, out):
        def step_resident():
            k = q.ps.stabilize_ptr(None, 0, 0, out.data_ptr(), capi.VS_MEM_DEVICE)
            assert k == q.ps.outputs, k
        return step_resident

    sampler = ClockSampler(R.local) if (R.rank == 0 and profile) else None
    # a step is `passes` videos; they are handed out one by one to the instances
    ms, t0, t1 = timed_inflight(R, [q.stream for q in pvs], [make_step(q, o) for q, o in zip(pvs, outs_dev)], steps * passes,
                                warmup * passes, lambda: sum(q.ps.launches for q in pvs))
    ms *= passes
    res["ms_per_step"] = ms
    res["launches"] = timed.launches
    res["clocks"] = sampler.stop(t0, t1) if sampler else None
    res["frames_per_step"] = R.world * F * passes
    res["value"] = res["frames_per_step"] / (ms / 1e3)
    res["sub"], res["block"], res["host_threads"] = sub, block, pv.host_threads
    corr, meas, ok = ps.records(pv.total)
    res["pairs_seen"] = int(len(ok) - 1)
    res["pairs_converged"] = int(ok[1:].sum())
    res["outputs_per_rank"] = int(ps.outputs)

    n_own, n_local, n_out = int((~ps.local_is_halo).sum()), len(ps.local_frames), int(ps.outputs)
    res["per_rank"] = {"frames": n_local, "keyframes": int((ps.local_frames % 2 == 1).sum()),
                       "pairs": n_own - (1 if R.rank == 0 else 0), "warped": n_out}
    res["videos_in_flight"] = inflight
    del outs_dev
    for q in pvs:
        q.ps.close()
    del pvs, pv, ps
    torch.cuda.empty_cache()

    # ---- per-kernel CUDA-event times over a second timed region: the rank's chunk as ONE sub-chunk, every stage one
    #      launch, back to back on one stream (with the lanes of the timed run a kernel's events also span the kernels
    #      beside it, and a solve lasts as long as its slowest pair however few pairs it has)
    if profile:
        pp = PartitionedVideo(R, args, W, H, F, F, 1, True, tag + "p", seed)
        pp.ps.upload_resident(pp.frames.ctypes.data, W * 3, pp.frame_bytes)
        pp.ps.set_lanes(1)
        pp.ps.synchronize()
        out_dev = torch.empty((max(pp.ps.outputs, 1), pp.ps.out_h, pp.ps.out_w, 3), dtype=torch.uint8, device="cuda")

        def step_profile():
            for _ in range(passes):
                pp.ps.stabilize_ptr(None, 0, 0, out_dev.data_ptr(), capi.VS_MEM_DEVICE)

        step_profile()
        lib.vs_ctx_profile_enable(pp.ps.ctx_handle, 1)
        lib.vs_ctx_profile_reset(pp.ps.ctx_handle)
        psteps = max(1, min(steps, 4))
        prof_ms, _, _ = timed(R, pp.stream, step_profile, psteps, 1)
        kernels = {}
        for k in range(capi.VS_KERNEL_COUNT):
            n, tot = C.c_int64(), C.c_double()
            lib.vs_ctx_profile_read(pp.ps.ctx_handle, k, C.byref(n), C.byref(tot))
            if n.value:
                kernels[lib.vs_kernel_name(k).decode()] = (n.value, tot.value)
        lib.vs_ctx_profile_enable(pp.ps.ctx_handle, 0)
        res["kernels_raw"] = kernels
        res["prof_videos"] = (psteps + 1) * passes
        res["prof_ms_per_video"] = prof_ms / passes
        del out_dev
        pp.ps.close()
        del pp
        torch.cuda.empty_cache()

    # ---- end to end: host frames in, host frames out, copies inside the timed region; sub-chunks interleaved over the ranks
    if e2e_steps > 0:
        pes = []
        # end to end a third video in flight fills more of the pipeline's fill and drain (0.86 -> 0.89 of the copy ceiling on
        # one GPU); with many ranks the host memory system is the limit and more instances only add contention
        inflight = max(1, args.e2e_inflight) if getattr(args, "e2e_inflight", 0) > 0 else (max(inflight, 3) if R.world <= 2 else inflight)
        for i in range(inflight):
            pes.append(PartitionedVideo(R, args, W, H, F, e2e_sub, 1, False, tag + "s%d" % i, seed, share=pes[0] if i else None, nv12=nv12))
        pe = pes[0]
        outs_host = [torch.empty((max(pe.ps.outputs, 1),) + pe.ps.out_shape, dtype=torch.uint8, pin_memory=True)
                     for _ in pes]

        def make_e2e(q, out):
            def step_e2e():
                k = q.ps.stabilize_ptr(q.frames.ctypes.data, q.row_stride, q.frame_bytes, out.data_ptr(), capi.VS_MEM_HOST)
                assert k == q.ps.outputs, k
            return step_e2e

        e2e_ms, _, _ = timed_inflight(R, [q.stream for q in pes], [make_e2e(q, o) for q, o in zip(pes, outs_host)], e2e_steps, 1)
        h2d = R.reduce(pe.h2d_bytes, "sum")
        d2h = R.reduce(pe.d2h_bytes + (len(pe.ps.local_frames)) * 36, "sum")
        res["e2e"] = {"value": R.world * F / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": e2e_ms, "frames_per_step": R.world * F, "sub_chunk_frames": e2e_sub,
                      "h2d_gbs": h2d / (e2e_ms / 1e3) / 1e9, "d2h_gbs": d2h / (e2e_ms / 1e3) / 1e9}
        res["e2e"]["videos_in_flight"] = inflight
        del outs_host
        for q in pes:
            q.ps.close()
        del pes, pe
        torch.cuda.empty_cache()
    return res


def run_gpu_arm(args):
    R = Ranks()
    torch = R.torch
    W, H, F, crop = args.width, args.height, args.frames, args.crop
    world, rank = R.world, R.rank

    main = measure_partitioned(R, args, W, H, F, args.steps, args.warmup, args.passes, max(1, args.steps // 2), args.e2e_sub, True, "m")
    ceiling = host_copy_ceiling(R)

    extra = {}
    if not args.no_extra:
        # 4K (BASELINE.json metric names 1080p AND 4K; configs[2]): 120 frames per GPU through the same partitioned pipeline
        a4 = types.SimpleNamespace(**vars(args))
        a4.width, a4.height, a4.frames = 3840, 2160, args.frames_4k
        r4 = measure_partitioned(R, a4, 3840, 2160, args.frames_4k, max(2, args.steps // 4), 2, max(1, args.passes // 4),
                                 2, args.e2e_sub, False, "k", seed=CLIP_SEED + 7)
        extra["4k"] = {"metric": "stabilized_frames_per_sec_4k", "value": r4["value"], "unit": UNIT, "n_gpus": world,
                       "ms_per_step": r4["ms_per_step"], "frames_per_step": r4["frames_per_step"],
                       "frames_per_video_per_gpu": args.frames_4k, "pairs_converged": r4["pairs_converged"], "pairs_seen": r4["pairs_seen"],
                       "e2e": r4.get("e2e")}

        # the same 1080p video as NV12 frames in and out (decoder output fed straight in; VS_CLIP_NV12): half the
        # bytes over PCIe, no colour conversion; the alignment is that of the gray frames
        rn = None
        if W % 2 == 0 and H % 2 == 0 and crop % 2 == 0:
            rn = measure_partitioned(R, args, W, H, F, max(2, args.steps // 2), 2, args.passes, max(1, args.steps // 2), args.e2e_sub,
                                     False, "n", nv12=True)
        if rn:
            extra["nv12"] = {"metric": "stabilized_frames_per_sec_1080p_nv12", "value": rn["value"], "unit": UNIT, "n_gpus": world,
                             "ms_per_step": rn["ms_per_step"], "frames_per_step": rn["frames_per_step"], "size": "%dx%d" % (W, H),
                             "pairs_converged": rn["pairs_converged"], "pairs_seen": rn["pairs_seen"], "e2e": rn.get("e2e"),
                             "note": "frames are NV12 in and out (1.5 B per pixel each way); no counterpart upstream"}

    if rank != 0:
        R.close()
        return 0

    # ---- roofline of the dominant kernel (rank 0's share: frames incl. the halo, its pairs, its warped frames)
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    pr = main["per_rank"]
    alg = algorithmic_bytes(W, H, crop, pr["frames"], pr["keyframes"], pr["pairs"], pr["warped"])
    videos = main["prof_videos"]
    per_kernel = {}
    for name, (n, tot) in main["kernels_raw"].items():
        b = alg.get(name)
        ms_video = tot / videos
        gbs = (b / (ms_video / 1e3) / 1e9) if b else None
        per_kernel[name] = {"launches_per_video": n / videos, "ms_per_video": ms_video, "ms_per_launch": tot / n,
                            "algorithmic_bytes_per_video": b, "algorithmic_gbs": gbs,
                            "frac_of_hbm_peak": (gbs / peak) if gbs else None}
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_video"])
    d = per_kernel[dominant]
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")   # dram bytes per video per kernel from the committed ncu --set full capture
    if os.path.exists(tpath) and (W, H, F) == (1920, 1080, 300):      # the capture is of the default configuration only
        try:
            t = json.load(open(tpath)).get(dominant)
            traffic = t / d["launches_per_video"] if t else None
        except Exception:
            traffic = None
    second = sorted(per_kernel, key=lambda k: -per_kernel[k]["ms_per_video"])[1] if len(per_kernel) > 1 else None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": d["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                "frac": d["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_video"] / d["launches_per_video"],
                "launches_per_video": d["launches_per_video"], "ms_per_launch": d["ms_per_launch"],
                "kernel_share_of_step": d["ms_per_video"] / max(main["prof_ms_per_video"], 1e-9)}
    if second:
        d2 = per_kernel[second]
        roofline["runner_up"] = {"kernel": second, "achieved": d2["algorithmic_gbs"], "frac": d2["frac_of_hbm_peak"],
                                 "ms_per_video": d2["ms_per_video"],
                                 "kernel_share_of_step": d2["ms_per_video"] / max(main["prof_ms_per_video"], 1e-9)}
    step_bytes = sum(v["algorithmic_bytes_per_video"] or 0 for v in per_kernel.values())
    whole = {"algorithmic_bytes_per_video": step_bytes, "ms_per_video": main["ms_per_step"] / args.passes,
             "frac_of_hbm_peak": step_bytes / (main["ms_per_step"] / args.passes / 1e3) / 1e9 / peak}

    # ---- CPU baseline on this box's cores (N=1 only), the same clip's leading frames
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, args.cpu_threads or cores))
        n = cpu_sample_frames(W, H, threads, crop, 12.0, F)
        fps, kind, dt = cpu_reference(_cpu_clip(W, H, n, CLIP_SEED), threads, crop)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d threads, one VideoStabilizer each, every thread stabilizes the first %d frames of the %d-frame %dx%d video "
                         "(same bytes as the GPU arm), %.1f s" % (threads, n, F, W, H, dt)}

    # ---- N=1 only: the other BASELINE.json configs as driver-run records
    if world == 1 and not args.no_extra:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        try:
            import config_bench
            import kernel_bench
            q = types.SimpleNamespace(clips=64, clip_frames=32, iters=3, quiet=True)
            extra["clips720"] = config_bench.clips720(q)
            sweep = kernel_bench.bench_warp(types.SimpleNamespace(size=None, mode=None, transform=None, iters=5, quiet=True))
            extra["warp_sweep"] = [{k: r[k] for k in ("mode", "size", "bytes_per_launch", "ms_per_launch", "algorithmic_gbs", "frac_of_hbm_peak")}
                                   for r in sweep]
            # the drop-in per-frame API (VideoStabilizer::processFrame, pageable cv::Mat in and out) and the batched
            # VideoAlignerParams sweep, as their C++ tools measure them
            tool = os.path.join(REPO, "video_stabilizer_b200", "bin", "stream_bench")
            if os.path.exists(tool):
                out = subprocess.run([tool, "1920", "1080", "96", "1", "4", "8"], capture_output=True, text=True, timeout=120).stdout
                extra["process_frame"] = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
            tool = os.path.join(REPO, "video_stabilizer_b200", "bin", "grid_search_align")
            if os.path.exists(tool):
                out = subprocess.run([tool, "1280", "720", "48"], capture_output=True, text=True, timeout=120).stdout
                extra["grid_search_align"] = [l for l in out.splitlines() if l.startswith(("Input", "Best", "54 "))]
        except Exception as e:      # noqa: BLE001  (an extra record must not take the headline down)
            extra["error"] = repr(e)

    e2e = main["e2e"]
    achieved = e2e["h2d_gbs"] + e2e["d2h_gbs"]
    ceil_sum = ceiling["h2d_duplex_gbs"] + ceiling["d2h_duplex_gbs"]
    e2e["host_ceiling"] = dict(ceiling, note="pinned copies on every rank at once, both directions together; GB/s summed over ranks")
    e2e["frac_of_host_ceiling"] = achieved / ceil_sum if ceil_sum > 0 else None
    e2e["value_per_gpu"] = e2e["value"] / world
    for key in ("4k", "nv12"):
        if key in extra and extra[key].get("e2e"):
            x = extra[key]["e2e"]
            x["frac_of_host_ceiling"] = (x["h2d_gbs"] + x["d2h_gbs"]) / ceil_sum if ceil_sum > 0 else None

    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic",
        "config": dict(workload_config(args, world), videos_per_step=args.passes, frames_per_step=main["frames_per_step"],
                       sub_chunk_frames=main["sub"], sub_chunks_per_rank=main["block"], host_threads_per_rank=main["host_threads"],
                       videos_in_flight=main["videos_in_flight"],
                       kernel_times="`kernels` and `roofline` come from a second timed region with a rank's chunk as one sub-chunk, "
                                    "every stage one launch, back to back on one stream; `value` runs the chunk as %d sub-chunks "
                                    "whose solves overlap the other stages" % main["block"]),
        "e2e": e2e,
        "gpu_launches": int(main["launches"]),
        "clocks": main["clocks"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "kernels": per_kernel,
        "whole_step": whole,
        "value_per_gpu": main["value"] / world,
        "align_ms_per_pair": sum(per_kernel[k]["ms_per_video"] for k in per_kernel if k != "bgr_warp") / max(pr["pairs"], 1),
        "pairs_converged": main["pairs_converged"], "pairs": main["pairs_seen"],
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    R.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300, help="frames of the video per GPU")
    ap.add_argument("--frames-4k", type=int, default=120, help="frames per GPU of the 4K extra record")
    ap.add_argument("--crop", type=int, default=0)
    ap.add_argument("--passes", type=int, default=15, help="videos per step (device-resident number)")
    ap.add_argument("--inflight", type=int, default=2, help="videos in flight at a time (independent stabilizer instances)")
    ap.add_argument("--sub-chunks", type=int, default=3, help="sub-chunks a rank's resident chunk is cut into (one solver lane each)")
    ap.add_argument("--e2e-sub", type=int, default=16, help="sub-chunk (frames) of the host-streamed partition")
    ap.add_argument("--e2e-inflight", type=int, default=0, help="videos in flight in the end-to-end region (0 = 3 on up to two GPUs, else --inflight)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="CPU worker threads (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the 4K / configs[3] / warp-sweep records")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
