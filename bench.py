#!/usr/bin/env python
"""bench.py — stabilized frames/s of the alignment-and-warp hot path on B200.

Workload (BASELINE.json configs[1], "video_test"): a 1920x1080 synthetic, procedurally
jittered 300-frame BGR clip through the full pipeline — BGR->gray + pyramid, keyframe
features, batched sparse Lucas-Kanade solve, the sequential host L1 smoother, cv-exact BGR
warp — with VideoStabilizerParams as video_test.cpp:53-54 sets them (crop_pixels = 0).
One step = one pass of the whole clip.  With N GPUs every rank stabilizes its own clip
(independent clips, no collective on the data path): weak scaling, value = all frames / max
step time over ranks.

  value  frames/s with the clip already resident in HBM (CUDA events, max over ranks)
  e2e    frames/s through ClipStabilizer::feed() with HOST buffers: pinned H2D of every
         frame and D2H of every stabilized frame inside the timed region
  roofline      dominant kernel of the timed region: algorithmic bytes / its CUDA-event time
  cpu_baseline  the reference's own host sources (oracle/_ref) or the oracle port, timed on
                this box's cores on a bounded sample of the same workload (rank 0, N=1)

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

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "stabilized_frames_per_sec_1080p"
UNIT = "frames/s"


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
    """Compulsory HBM traffic per launch of each kernel (each input byte read once, each
    output byte written once), BASELINE.md section 4 / DESIGN.md."""
    lv = level_table(w, h)
    px = [L["w"] * L["h"] for L in lv]
    tiles = sum(L["tiles"] for L in lv)
    ow, oh = w - 2 * crop, h - 2 * crop
    return {
        "bgr2gray": 4 * w * h * n_frames,
        "ingest_bgr_gray_l1": (3 * w * h + px[0] + px[1]) * n_frames,
        # one launch per level; the per-launch figure reported is the L0->L1 launch (3/4 of the bytes)
        "pyr_down": (px[0] + px[1]) * n_frames,
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


# ----------------------------------------------------------------------------- CPU reference
def _cpu_frames(width, height, n, seed):
    """A short clip of the same synthetic workload for the CPU arm (numpy renderer: no GPU needed)."""
    from video_stabilizer_b200 import synth
    return synth.make_clip_numpy(width, height, n, seed)[0]


_CPU_CLIPS = {}


def cpu_reference(width, height, frames_per_thread, threads, crop, passes=1, clip=None):
    """Frames/s of the reference CPU implementation: `threads` workers, one stabilizer each (the
    reference's scale-out recipe, grid_search_align.cpp:105-118,159-210).  Every worker plays the
    clip forwards, then backwards, ... `passes` times (the turn-around keeps the motion continuous),
    so the sample is threads x frames_per_thread x passes frames.  Returns (frames_per_s, kind, seconds)."""
    from oracle import binding as ob
    use_ref = ob.ref_available()
    if use_ref:
        try:
            ob.load_ref(fast=True)
        except Exception:
            use_ref = False
    kind = "reference" if use_ref else "port"
    if clip is None:
        key = (width, height, frames_per_thread)
        if key not in _CPU_CLIPS:
            _CPU_CLIPS[key] = _cpu_frames(width, height, frames_per_thread, 4242)
        clip = _CPU_CLIPS[key]
    p = ob.stab_params_default()
    p.crop_pixels = crop

    def work(t):
        st = ob.RefStabilizer(p, fast=True) if use_ref else ob.Stabilizer(p, fast=True)
        for k in range(passes):
            for f in (clip if k % 2 == 0 else clip[::-1]):      # read-only input shared by the workers
                st.process(f)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return threads * len(clip) * passes / dt, kind, dt


def cpu_passes_for(width, height, frames_per_thread, threads, crop, target_seconds, clip=None):
    """One calibration pass, then how many passes make the sample last about target_seconds."""
    _, _, dt = cpu_reference(width, height, frames_per_thread, threads, crop, 1, clip)
    return int(max(1, min(200, round(target_seconds / max(dt, 1e-3)))))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.cpu_threads or cores))
    per_thread = args.cpu_frames
    # each step is a bounded sample sized so that the whole --steps/--warmup run stays around 2-3 minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    passes = cpu_passes_for(args.width, args.height, per_thread, threads, args.crop, min(budget, 20.0))
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        fps, kind, dt = cpu_reference(args.width, args.height, per_thread, threads, args.crop, passes)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    frames_per_step = threads * per_thread * passes
    value = frames_per_step / (ms / 1000.0)
    sample = "%d threads x a %d-frame %dx%d clip played %d times (forwards/backwards) per step, one VideoStabilizer per thread" % (
        threads, per_thread, args.width, args.height, passes)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic",
        "config": workload_config(args, frames=frames_per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, frames):
    return {
        "workload": "video_test: %dx%d synthetic procedurally-jittered BGR clip, full pipeline (BGR->gray pyramid, "
                    "keyframe features, sparse LK solve, host L1 smoother, cv-exact BGR warp), "
                    "VideoStabilizerParams of video_test.cpp:53-54 (crop_pixels=%d, lag=10)" % (args.width, args.height, args.crop),
        "frames_per_step_per_gpu": frames, "width": args.width, "height": args.height,
        "partition": "one independent clip per GPU, no collective on the data path",
        "l2": "inputs larger than L2: %.2f GB of frames read per step vs 126 MB L2" % (frames * args.width * args.height * 3 / 1e9),
    }


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200 import host, synth
    from video_stabilizer_b200.imgproc import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    W, H, F, crop = args.width, args.height, args.frames, args.crop
    frame_bytes = W * H * 3

    # ---- synthetic clip in pinned host memory, rendered on the GPU through vs_bgr_warp_u8
    pinned = torch.empty((F, H, W, 3), dtype=torch.uint8, pin_memory=True)
    frames = pinned.numpy()
    gen_ctx = Context(local)
    synth.make_clip_gpu(gen_ctx, W, H, F, seed=100 + rank, out=frames, chunk=50)
    gen_ctx.close()

    p = host.stab_params_default()
    p.crop_pixels = crop
    cs = host.ClipStabilizer(W, H, F, p, device=local)
    if args.pipeline_frames:
        cs.set_pipeline_frames(args.pipeline_frames)
    # a real (non-default) torch stream, borrowed by the library: the CUDA events of the timed
    # region are recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    cs.set_stream(stream.cuda_stream)
    SOLVER_LANES = int(os.environ.get("VSTAB_LANES", "3"))
    cs.set_solver_lanes(SOLVER_LANES)
    lib = capi.load()
    n_out = F - p.lag
    out_dev = torch.empty((n_out, cs.out_h, cs.out_w, 3), dtype=torch.uint8, device="cuda")
    out_host = torch.empty((n_out, cs.out_h, cs.out_w, 3), dtype=torch.uint8, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        n0 = cs.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1) / steps
        timed.launches = cs.launches - n0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    # ---- device-resident: frames already in the ring; a step = pyramids + features + solve +
    #      D2H of the transforms + host trajectory + warp, outputs stay in HBM
    cs.upload_only(0, frames.ctypes.data, F, W * 3, frame_bytes, capi.VS_MEM_HOST)
    cs.synchronize()

    def step_resident():
        cs.reset()
        k = cs.feed_resident(F, out_dev.data_ptr(), capi.VS_MEM_DEVICE)
        assert k == n_out, k

    # ---- end to end: host frames in, host frames out, copies inside the timed region
    def step_e2e():
        cs.reset()
        k = cs.feed_ptr(frames.ctypes.data, F, W * 3, frame_bytes, capi.VS_MEM_HOST, out_host.data_ptr(), capi.VS_MEM_HOST)
        assert k == n_out, k

    sampler = ClockSampler(local) if rank == 0 else None
    ms, t0, t1 = timed(step_resident, args.steps, args.warmup)
    launches = timed.launches
    clocks = sampler.stop(t0, t1) if sampler else None
    meas, ok, corr = cs.last_records(F)

    # ---- per-kernel CUDA-event times over a second timed region of the same steps, with the stages back to back on one
    #      stream (one solver lane): with the lanes of the timed run a kernel's events also span the kernels beside it
    cs.set_solver_lanes(1)
    lib.vs_ctx_profile_enable(cs.ctx_handle, 1)
    lib.vs_ctx_profile_reset(cs.ctx_handle)
    prof_ms, _, _ = timed(step_resident, args.steps, 1)
    cs.set_solver_lanes(SOLVER_LANES)
    kernels = {}
    for k in range(capi.VS_KERNEL_COUNT):
        n, tot = C.c_int64(), C.c_double()
        lib.vs_ctx_profile_read(cs.ctx_handle, k, C.byref(n), C.byref(tot))
        if n.value:
            kernels[lib.vs_kernel_name(k).decode()] = (n.value, tot.value)
    lib.vs_ctx_profile_enable(cs.ctx_handle, 0)

    e2e_ms, _, _ = timed(step_e2e, max(1, args.steps // 2), 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_key = F // 2
    alg = algorithmic_bytes(W, H, crop, F, n_key, F - 1, n_out)
    steps_profiled = args.steps + 1
    per_kernel = {}
    lv = level_table(W, H)
    px = [L["w"] * L["h"] for L in lv]
    # algorithmic bytes of one STEP per kernel (a kernel may be launched several times per step:
    # pyr_down once per level, bgr_warp in batches that overlap the host trajectory)
    alg_step = dict(alg)
    alg_step["pyr_down"] = sum(px[i] + px[i + 1] for i in range(len(px) - 1)) * F
    for name, (n, tot) in kernels.items():
        launches_per_step = n / steps_profiled
        ms_step = tot / steps_profiled
        b = alg_step.get(name)
        gbs = (b / (ms_step / 1e3) / 1e9) if b else None
        per_kernel[name] = {"launches_per_step": launches_per_step, "ms_per_step": ms_step,
                            "ms_per_launch": tot / n, "algorithmic_bytes_per_step": b, "algorithmic_gbs": gbs,
                            "frac_of_hbm_peak": (gbs / peak) if gbs else None}
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    d = per_kernel[dominant]
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")   # dram bytes per step from the committed ncu --set full capture
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath)).get(dominant)
            traffic = t / d["launches_per_step"] if t else None
        except Exception:
            traffic = None
    second = sorted(per_kernel, key=lambda k: -per_kernel[k]["ms_per_step"])[1] if len(per_kernel) > 1 else None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": d["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                "frac": d["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_step"] / d["launches_per_step"],
                "launches_per_step": d["launches_per_step"], "ms_per_launch": d["ms_per_launch"],
                "kernel_share_of_step": d["ms_per_step"] / max(prof_ms, 1e-9)}
    if second:
        d2 = per_kernel[second]
        roofline["runner_up"] = {"kernel": second, "achieved": d2["algorithmic_gbs"], "frac": d2["frac_of_hbm_peak"],
                                 "ms_per_step": d2["ms_per_step"], "kernel_share_of_step": d2["ms_per_step"] / max(prof_ms, 1e-9)}

    # ---- CPU baseline on this box's cores (N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, args.cpu_threads or cores))
        cpu_clip = frames[: args.cpu_frames]          # the first frames of the very clip the GPU arm stabilizes
        passes = cpu_passes_for(W, H, args.cpu_frames, threads, crop, 12.0, cpu_clip)
        fps, kind, dt = cpu_reference(W, H, args.cpu_frames, threads, crop, passes, cpu_clip)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d threads x the first %d frames of the %dx%d clip played %d times (forwards/backwards), one VideoStabilizer per thread, %.1f s" % (
                   threads, args.cpu_frames, W, H, passes, dt)}

    value = world * F / (ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic",
        "config": dict(workload_config(args, F), solver_lanes=SOLVER_LANES,
                       kernel_times="`kernels` and `roofline` come from a second timed region with the stages back to back on one "
                                    "stream (one solver lane); `value` runs the chunk as %d pieces whose solves overlap the other stages" % SOLVER_LANES),
        "e2e": {"value": world * F / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": F * frame_bytes,
                "d2h_bytes_per_step": n_out * cs.out_frame_bytes + (F - 1) * 36, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "kernels": per_kernel,
        "align_ms_per_pair": sum(per_kernel[k]["ms_per_step"] for k in per_kernel if k != "bgr_warp") / (F - 1),
        "pairs_converged": int(ok.sum()), "pairs": F - 1,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--crop", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=16, help="frames per CPU worker thread in the CPU arm / baseline")
    ap.add_argument("--cpu-threads", type=int, default=0, help="CPU worker threads (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-frames", type=int, default=0, help="sub-chunk of the host-to-host pipeline (0 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
