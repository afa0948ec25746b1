"""Parity of the batched device pipeline (vs_clip_*) against the restated VideoAligner.

BASELINE.json bars: keypoint index sets bit-exact, recovered transforms within 0.01 px
corner displacement, warped frames within 1 LSB (here: bit-exact for the cv-exact mode).
"""
import numpy as np
import pytest

from util import corner_displacement

pytestmark = pytest.mark.gpu

TOL_PX = 0.01


def run_oracle(ob, frames, params=None):
    """Feed frames one by one through the restated VideoAligner, collecting everything."""
    al = ob.Aligner(params)
    res = []
    for i, f in enumerate(frames):
        ok, T = al.align(f)
        rec = dict(ok=ok, T=T.copy(), curr=al.lib.vo_aligner_curr_index(al.h), phase=al.phase())
        if i > 0:
            L = al.levels
            rec["iters"] = [al.iterations(l) for l in range(L)]
            rec["wd"] = [[al.warpdiff(l, a) if rec["iters"][l] else None for a in range(2)] for l in range(L)]
            rec["sel"] = [[al.selected(l, a) if rec["iters"][l] else None for a in range(2)] for l in range(L)]
            rec["kp"] = [[al.keypoints(l, a) for a in range(2)] for l in range(L)]
            rec["jac"] = [[al.jacobians(l, a) for a in range(2)] for l in range(L)]
            rec["pyr"] = [[al.pyramid(s, l) for l in range(L)] for s in range(2)]
        res.append(rec)
    return al, res


def make_params(ob, **kw):
    """The same VideoAlignerParams for the oracle and for the C ABI."""
    from video_stabilizer_b200 import _capi as capi
    po = ob.align_params_default()
    pg = capi.VsAlignParams()
    capi.load().vs_align_params_default(pg)
    for k, v in kw.items():
        setattr(po, k, v)
        setattr(pg, k, v)
    return po, pg


def check_clip_against_oracle(gpu, ob, frames, deep=True, **param_overrides):
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    n, h, w, _ = frames.shape
    po, pg = make_params(ob, **param_overrides)
    clip = Clip(w, h, n, params=pg, debug=True, ctx=gpu)
    clip.upload(0, frames)
    clip.build_pyramids(0, n)
    pairs, keyframes = pairs_for_frames(0, n)
    clip.build_keyframes(keyframes)
    T, status, iters = clip.align(pairs)
    al, ref = run_oracle(ob, frames, po)
    assert clip.levels == al.levels
    worst = 0.0
    for i in range(1, n):
        r = ref[i]
        p = i - 1
        # frame i sits in oracle slot r["curr"]; the other slot holds frame i-1
        if deep:
            for l in range(clip.levels):
                assert np.array_equal(clip.get_gray(i, l), r["pyr"][r["curr"]][l]), ("pyramid", i, l)
            kf = i if i % 2 == 1 else i - 1
            for l in range(clip.levels):
                for a in range(2):
                    assert np.array_equal(clip.get_keypoints(kf, l, a), r["kp"][l][a]), ("keypoints", i, l, a)
                    assert np.array_equal(clip.get_jacobians(kf, l, a), r["jac"][l][a]), ("jacobians", i, l, a)
        if param_overrides.get("phase_correlate"):
            # shift and response of cv::phaseCorrelate (alignment.cpp:374): same serial f64 sums, bit for bit
            assert np.array_equal(clip.get_phase(p), r["phase"]), ("phase", i, clip.get_phase(p), r["phase"])
        assert list(iters[p]) == r["iters"], ("iterations", i, list(iters[p]), r["iters"])
        assert bool(status[p]) == r["ok"], ("status", i)
        for l in range(clip.levels):
            if r["iters"][l] == 0:
                continue
            for a in range(2):
                assert np.array_equal(clip.get_warpdiff(p, l, a), r["wd"][l][a]), ("warpdiff", i, l, a)
                assert np.array_equal(clip.get_selected(p, l, a), r["sel"][l][a]), ("selection", i, l, a)
        d = corner_displacement(T[p], r["T"], w, h)
        worst = max(worst, d)
        assert d <= TOL_PX, ("transform", i, T[p], r["T"], d)
    clip.close()
    return worst, T, status, ref


@pytest.mark.parametrize("w,h,n,seed", [(320, 180, 8, 0), (640, 360, 6, 1), (250, 141, 5, 2), (1280, 720, 4, 3)])
def test_clip_alignment_matches_oracle(gpu, ob, w, h, n, seed):
    from video_stabilizer_b200 import synth
    frames, _ = synth.make_clip_numpy(w, h, n, seed)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames)
    assert status.sum() >= n - 2          # the synthetic jitter is inside the convergence basin
    assert worst < 1e-6                   # in practice the only difference is f64 summation order


def test_clip_alignment_1080p(gpu, ob):
    """configs[0]: one 1920x1080 pair with known homography (plus its mirror-parity pair)."""
    from video_stabilizer_b200 import synth
    frames, poses = synth.make_clip_gpu(gpu, 1920, 1080, 3, 7)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames)
    assert status.all()
    # sanity against ground truth: the reference's quarter-step GN stops ~0.1 px short
    for p in range(2):
        truth = ob.tf_compose(ob.tf_inverse(poses[p + 1]), poses[p])
        assert corner_displacement(T[p], truth, 1920, 1080) < 0.5


def test_failure_paths_match_oracle(gpu, ob):
    """Non-convergence / over-displacement must return false exactly when the reference does."""
    from video_stabilizer_b200 import synth
    w, h = 320, 180
    frames, _ = synth.make_clip_numpy(w, h, 6, 5, step=14.0, limit=60.0, ab=0.02)
    # hitting max_iters without converging returns false (alignment.cpp:661-667)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames, deep=False, max_iters=3)
    assert not status.all(), "max_iters=3 should make the reference give up on some pairs"
    # converged displacement above max_displacement returns false (alignment.cpp:670-677)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames, deep=False, max_displacement=0.75)
    assert not status.all(), "max_displacement=0.75 should reject some pairs"
    # other parameter settings keep parity too
    check_clip_against_oracle(gpu, ob, frames, deep=False, threshold=0.1, smallest_fraction=0.5)
    check_clip_against_oracle(gpu, ob, frames, deep=False, pyramid_min_width=60, pyramid_min_height=30)
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (4, h, w, 3), dtype=np.uint8)     # unrelated frames
    check_clip_against_oracle(gpu, ob, noise, deep=False)
    flat = np.full((3, h, w, 3), 77, np.uint8)                      # zero gradients, singular Hessian
    check_clip_against_oracle(gpu, ob, flat, deep=False)


def test_phase_correlate_operator(gpu, ob):
    """cv::phaseCorrelate on the device: bit-identical to the restatement (same serial f64 sums) on even, odd and
    padded sizes, and within 1e-4 of the cv2 4.13 fixtures."""
    import os
    from video_stabilizer_b200 import imgproc as ip
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "phase_correlate.npz"))
    ref = g["result"]
    for k in range(len(ref)):
        a, b = g["a%d" % k], g["b%d" % k]
        (sx, sy), resp = ip.PhaseCorrelate(a, b, gpu)
        got = np.array([sx, sy, resp])
        assert np.array_equal(got, ob.phase_correlate_u8(a, b)), (k, got, ob.phase_correlate_u8(a, b))
        assert np.abs(got - ref[k]).max() < 1e-4, (k, got, ref[k])
    # the reference's own shift pattern (align_test.cpp:358-400): (5, 7) within half a pixel
    a = np.zeros((64, 64), np.uint8); a[20:30, 20:30] = 255
    b = np.zeros((64, 64), np.uint8); b[27:37, 25:35] = 255
    (sx, sy), resp = ip.PhaseCorrelate(a, b, gpu)
    assert abs(sx - 5) <= 0.5 and abs(sy - 7) <= 0.5
    # strided views and a level-2 sized image (480 x 270: the size the 1080p path runs)
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (300, 520), dtype=np.uint8)
    a, b = big[10:280, 20:500], big[13:283, 15:495]
    (sx, sy), resp = ip.PhaseCorrelate(a, b, gpu)
    assert np.array_equal(np.array([sx, sy, resp]), ob.phase_correlate_u8(a, b))
    assert abs(sx - 5) < 0.01 and abs(sy + 3) < 0.01 and resp > 0.9   # content moved by (+5, -3)
    # very wide / very tall images: twiddle and row tables beyond the default 48 KB of dynamic shared memory
    for (hh, ww) in ((48, 1536), (3072, 32)):
        big = rng.integers(0, 256, (hh + 8, ww + 8), dtype=np.uint8)
        a, b = np.ascontiguousarray(big[2:2 + hh, 3:3 + ww]), np.ascontiguousarray(big[4:4 + hh, 1:1 + ww])
        (sx, sy), resp = ip.PhaseCorrelate(a, b, gpu)
        assert np.array_equal(np.array([sx, sy, resp]), ob.phase_correlate_u8(a, b)), (hh, ww)
        assert abs(sx - 2) < 0.1 and abs(sy + 2) < 0.1


def test_phase_correlate_errors_are_reported(gpu):
    import ctypes as C
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    lib = gpu.lib
    a, b = np.zeros((16, 24), np.uint8), np.zeros((16, 20), np.uint8)
    out = np.zeros(3)
    r = lib.vs_phase_correlate_u8(gpu.handle, C.byref(capi.img_of(a)), C.byref(capi.img_of(b)), capi.ptr(out), capi.VS_MEM_HOST)
    assert r == -1 and b"one non-empty size" in lib.vs_last_error(gpu.handle)
    r = lib.vs_phase_correlate_u8(gpu.handle, C.byref(capi.img_of(a)), C.byref(capi.img_of(a)), None, capi.VS_MEM_HOST)
    assert r == -1
    # the tap needs an align call made with phase_correlate on
    clip = Clip(64, 48, 2, debug=True, ctx=gpu)
    assert lib.vs_clip_get_phase(clip.handle, 0, capi.ptr(out)) == -1
    clip.close()
    # level 2 must exist (alignment.hpp:69 indexes it unconditionally): a 2-level pyramid is refused, not read out of bounds
    pg = capi.VsAlignParams()
    capi.load().vs_align_params_default(pg)
    pg.phase_correlate = 1
    pg.pyramid_min_width, pg.pyramid_min_height = 40, 30
    clip = Clip(96, 64, 2, params=pg, ctx=gpu)
    assert clip.levels == 2
    clip.upload(0, np.zeros((2, 64, 96, 3), np.uint8))
    clip.build_pyramids(0, 2)
    pairs, keys = pairs_for_frames(0, 2)
    clip.build_keyframes(keys)
    with pytest.raises(Exception, match="level 2"):
        clip.align(pairs)
    clip.close()


@pytest.mark.parametrize("w,h,n,seed,step", [(320, 180, 8, 0, 2.0), (640, 360, 6, 1, 9.0), (250, 141, 5, 2, 5.0)])
def test_clip_alignment_with_phase_correlate_matches_oracle(gpu, ob, w, h, n, seed, step):
    """VideoAlignerParams::phase_correlate (alignment.cpp:369-388): the seed, and everything the solver derives from
    it (warp-diffs, selections, iteration counts, status), identical to the restated aligner."""
    from video_stabilizer_b200 import synth
    frames, _ = synth.make_clip_numpy(w, h, n, seed, step=step, limit=40.0)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames, deep=False, phase_correlate=1)
    assert worst < 1e-6
    assert any(r["phase"][2] > 0.5 for r in ref[1:]), "the seed should be accepted on some pairs"
    # a threshold nobody passes: the seed is computed and ignored, results equal the plain run
    worst, T2, status2, _ = check_clip_against_oracle(gpu, ob, frames, deep=False, phase_correlate=1, phase_correlate_threshold=1e9)
    _, T0, status0, _ = check_clip_against_oracle(gpu, ob, frames, deep=False)
    assert np.array_equal(T2, T0) and np.array_equal(status2, status0)


def test_clip_warp_matches_oracle(gpu, ob):
    from video_stabilizer_b200 import synth
    from video_stabilizer_b200.clip import Clip
    w, h, n = 640, 360, 4
    frames, _ = synth.make_clip_numpy(w, h, n, 9)
    clip = Clip(w, h, n, ctx=gpu)
    clip.upload(0, frames)
    T = np.array([[0, 0, 0, 0], [0.002, -0.001, 3.7, -1.2], [-0.004, 0.003, -12.5, 8.25], [0.01, 0.01, 0.5, 0.5]], np.float64)
    for crop in (0, 32):
        for mode in (0, 1, 2):
            out = clip.warp([0, 1, 2, 3], T, mode=mode, crop=crop)
            for i in range(n):
                assert np.array_equal(out[i], ob.warp_bgr(frames[i], T[i], mode, 0, crop)), (mode, crop, i)
    out = clip.warp([3, 0], T[:2])
    assert np.array_equal(out[0], ob.warp_bgr(frames[3], T[0])) and np.array_equal(out[1], ob.warp_bgr(frames[0], T[1]))
    assert np.array_equal(clip.get_bgr(2), frames[2])
    clip.close()


@pytest.mark.parametrize("w,h", [(640, 360), (650, 187), (161, 40)])
def test_clip_warp_row_group_kernel_every_path(gpu, ob, w, h):
    """The production warp (k_bgr_warp_cv_rows) against the oracle on transforms that exercise each of its paths:
    regular groups (near identity), fx = fy = 0 everywhere (identity, integer shifts), irregular groups (moderate
    rotation / scale: the per-pixel list), tiles whose source box does not fit (large rotation / scale: global
    path), content leaving the frame on every side, widths and crops that are not multiples of 4 or 128."""
    from video_stabilizer_b200.clip import Clip
    rng = np.random.default_rng(w * 7 + h)
    frames = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    clip = Clip(w, h, 3, ctx=gpu)
    clip.upload(0, frames)
    Ts = [[0, 0, 0, 0], [0, 0, 7, -3], [0, 0, -33, 18], [0.0007, -0.0004, 1.37, -2.61], [-0.002, 0.0015, -9.8, 4.4],
          [0.001, 0.02, 3.3, 2.2], [0.03, -0.015, -4.1, 0.7], [-0.04, 0.0, 0.25, 0.25], [0.0, 0.2, 5.5, -7.5],
          [-0.3, 0.05, 2.0, 1.0], [0.5, -0.4, 0.0, 0.0], [0.001, 0.001, w * 0.9, 0.0], [0.001, -0.001, 0.0, -h * 0.95],
          [0.0, 0.0, 3.0 * w, 0.0], [0, 0, 0.5, 0.5], [0, 0, 1.0 / 32, 0], [1e-13, 0, 0.49999, 0.0]]
    for crop in (0, 5, 32):
        if w - 2 * crop < 8 or h - 2 * crop < 8:
            continue
        for i in range(0, len(Ts), 3):
            T = np.array(Ts[i:i + 3], np.float64)
            slots = [(i + k) % 3 for k in range(len(T))]
            out = clip.warp(slots, T, crop=crop)
            for k in range(len(T)):
                want = ob.warp_bgr(frames[slots[k]], T[k], 0, 0, crop)
                assert np.array_equal(out[k], want), (crop, Ts[i + k], int(np.abs(out[k].astype(int) - want).max()))
    clip.close()


def test_synth_gpu_renderer_is_bit_identical_to_numpy(gpu):
    from video_stabilizer_b200 import synth
    a, pa = synth.make_clip_numpy(320, 180, 3, 4)
    b, pb = synth.make_clip_gpu(gpu, 320, 180, 3, 4)
    assert np.array_equal(pa, pb) and np.array_equal(a, b)


def test_frame_chunk_partition_equals_single_run(gpu):
    """A video split into even-aligned chunks with a one-frame halo (one chunk per GPU in
    production; here the chunks run one after the other on one device) gives bit-identical
    transforms to the single run: pairs are independent and the keyframe parity is global."""
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200 import partition, synth
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    w, h, n = 320, 180, 21
    frames, _ = synth.make_clip_numpy(w, h, n, 13)
    clip = Clip(w, h, n, ctx=gpu)
    clip.upload(0, frames)
    clip.build_pyramids(0, n)
    pairs, keys = pairs_for_frames(0, n)
    clip.build_keyframes(keys)
    T_all, st_all, it_all = clip.align(pairs)
    clip.close()
    for world in (2, 3):
        T_parts, st_parts = [], []
        for rank, chunk in enumerate(partition.frame_chunks(n, world)):
            up0, up1 = partition.chunk_upload_range(chunk, rank)
            plist, kslots = partition.chunk_pairs(chunk, rank)
            if not plist:
                continue
            c = Clip(w, h, up1 - up0, ctx=gpu)
            c.upload(0, frames[up0:up1])
            c.build_pyramids(0, up1 - up0)
            c.build_keyframes(kslots)
            arr = (capi.VsPair * len(plist))(*[capi.VsPair(*p) for p in plist])
            T, st, _ = c.align(arr)
            T_parts.append(T)
            st_parts.append(st)
            c.close()
        assert np.array_equal(np.concatenate(T_parts), T_all)
        assert np.array_equal(np.concatenate(st_parts), st_all)


def test_clip_alignment_4k(gpu, ob):
    """BASELINE.json configs[2] shape: 3840x2160, 7 pyramid levels, 20736 tiles at L0 — the
    selection's candidate lists no longer fit in shared memory and come from global scratch."""
    from video_stabilizer_b200 import synth
    frames, poses = synth.make_clip_gpu(gpu, 3840, 2160, 9, 17)
    # 8 pairs, every stage compared: pyramids (the fused ingest at 4K), keypoints, Jacobians, warp-diffs, selections,
    # iteration counts, status, transforms
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames, deep=True)
    assert status.all() and worst < 1e-6


def test_clip_alignment_8k(gpu, ob):
    """7680x4320: 82 944 tiles at L0, more than a 16-bit tile index and more keys than shared memory holds: the selection
    keys of a pair live in global memory (17-bit tile index) and only its chunk masks on chip.  Pyramids, warp-diffs,
    selections, iteration counts, status and transforms of two pairs against the oracle."""
    from video_stabilizer_b200 import synth
    frames, poses = synth.make_clip_gpu(gpu, 7680, 4320, 3, 23, chunk=4)
    worst, T, status, ref = check_clip_against_oracle(gpu, ob, frames, deep=False)
    assert status.all() and worst < 1e-6


def test_clips_in_shared_launches_equal_each_clip_alone(gpu):
    """BASELINE.json configs[3] (many independent 720p clips aligned and warped concurrently): the pyramids, keyframe
    features, ONE solver launch and ONE warp launch over all clips give each clip exactly what it gets alone."""
    from video_stabilizer_b200 import _capi as capi, synth
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    w, h, nclips, nf = 1280, 720, 6, 6
    clips = [synth.make_clip_gpu(gpu, w, h, nf, 200 + c)[0] for c in range(nclips)]
    T_corr = np.array([0.001, -0.002, 3.25, -1.5])
    alone = []
    for fr in clips:
        c = Clip(w, h, nf, ctx=gpu)
        c.upload(0, fr)
        c.build_pyramids(0, nf)
        pairs, keys = pairs_for_frames(0, nf)
        c.build_keyframes(keys)
        T, st, it = c.align(pairs)
        warped = c.warp(list(range(nf)), np.tile(T_corr, (nf, 1)))
        alone.append((T, st, it, warped))
        c.close()
    big = Clip(w, h, nclips * nf, max_pairs=nclips * (nf - 1), ctx=gpu)
    allp, keys = [], []
    for ci, fr in enumerate(clips):
        big.upload(ci * nf, fr)
        p, k = pairs_for_frames(0, nf, slot_of=lambda f, ci=ci: ci * nf + f)
        allp.extend((q.template_slot, q.keyframe_slot, q.invert) for q in p)
        keys.extend(ci * nf + f for f in k)
    big.build_pyramids(0, nclips * nf)
    big.build_keyframes(keys)
    arr = (capi.VsPair * len(allp))(*[capi.VsPair(*q) for q in allp])
    T, st, it = big.align(arr)
    warped = big.warp(list(range(nclips * nf)), np.tile(T_corr, (nclips * nf, 1)))
    for ci in range(nclips):
        a = alone[ci]
        sl = slice(ci * (nf - 1), (ci + 1) * (nf - 1))
        assert np.array_equal(T[sl], a[0]) and np.array_equal(st[sl], a[1]) and np.array_equal(it[sl], a[2]), ci
        assert np.array_equal(warped[ci * nf:(ci + 1) * nf], a[3]), ci
    big.close()


def test_lane_parallel_svd_inverse_equals_the_serial_one_bit_for_bit(gpu, ob):
    """The solver inverts its 4x4 Hessian with a Jacobi SVD spread over four lanes; it must reproduce the serial
    restatement of cv::SVD + Mat::inv(DECOMP_SVD) bit for bit (same operations, same order), on well-conditioned
    Hessians, on the regularised branch (cond > 1e6, alignment.cpp:566-571) and on singular input."""
    from video_stabilizer_b200 import _capi as capi
    rng = np.random.default_rng(5)
    mats = []
    for i in range(200):
        n = int(rng.integers(8, 4000))
        J = rng.normal(size=(n, 4)) * rng.uniform(0.1, 300, size=4)
        if i % 3 == 0:
            J[:, 3] = 0 if i % 2 else J[:, 2] * (1 + 1e-9 * rng.normal(size=n))     # singular / nearly dependent columns
        if i % 7 == 0:
            J[:, 0] *= 1e-5
        mats.append(J.T @ J)
    mats.append(np.zeros((4, 4)))
    mats.append(np.eye(4))
    H = np.ascontiguousarray(np.stack(mats), np.float64)
    n = H.shape[0]
    quad, serial, cond = np.zeros_like(H), np.zeros_like(H), np.zeros(n)
    capi.check(gpu.handle, gpu.lib.vs_debug_invert4(gpu.handle, capi.ptr(H), n, capi.ptr(quad), capi.ptr(serial), capi.ptr(cond)),
               "vs_debug_invert4")
    assert np.array_equal(quad.view(np.uint64), serial.view(np.uint64))
    assert (cond > 1e6).sum() > 10 and (cond < 1e6).sum() > 10
    for i in range(n):
        w, _, _ = ob.svd4(H[i])
        Hi = H[i].copy()
        if w[0] / (w[3] + 1e-10) > 1e6:
            Hi[np.arange(4), np.arange(4)] += 1e-6 * w[0]
        want = ob.inv4_svd(Hi)
        scale = np.abs(want).max() + 1e-300
        assert np.abs(quad[i] - want).max() <= 1e-9 * scale, i
