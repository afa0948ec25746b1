"""The C++ host layer (VideoAligner, VideoStabilizer, ClipStabilizer, imgproc.hpp operators)
on the GPU against the oracle: the reference-facing classes are the thing under test, called
through libvstab_host.so exactly as a C++ caller would call them."""
import os

import numpy as np
import pytest

from util import corner_displacement, noise_image

pytestmark = pytest.mark.gpu

TOL_PX = 0.01     # BASELINE.json: recovered transforms within 0.01 px corner displacement
TOL_LSB = 1       # BASELINE.json: warped frames within 1 LSB


@pytest.fixture(scope="module")
def host(gpu):
    from video_stabilizer_b200 import host
    host.load()
    return host


def _clip(w, h, n, seed, **kw):
    from video_stabilizer_b200 import synth
    return synth.make_clip_numpy(w, h, n, seed, **kw)[0]


def test_operators_through_cpp_wrappers(host, ob):
    rng = np.random.default_rng(0)
    img = noise_image(rng, 135, 240, smooth=2)
    ok, half = host.PyrDown(img)
    assert ok and np.array_equal(half, ob.pyr_down(img))
    ok, gx, gy = host.GradXY(img)
    ogx, ogy = ob.grad_xy(img)
    assert ok and np.array_equal(gx, ogx) and np.array_equal(gy, ogy)
    ok, tile, lmx, lmy = host.GradArgMax(gx, gy)
    assert ok and tile == ob.tile_size(240, 135)
    olx, oly = ob.grad_argmax(ogx, ogy, tile)
    assert np.array_equal(lmx, olx) and np.array_equal(lmy, oly)
    ok, jx, jy = host.SparseJacobian(gx, gy, lmx, lmy)
    ojx, ojy = ob.sparse_jac(ogx, ogy, olx, oly)
    assert ok and np.array_equal(jx, ojx) and np.array_equal(jy, ojy)
    key = noise_image(rng, 135, 240, smooth=2)
    T = np.array([0.002, -0.001, 1.3, -0.7])
    ok, wd = host.SparseWarpDiff(img, key, lmx, T)
    assert ok and np.array_equal(wd, ob.sparse_warpdiff(img, key, olx, T))
    k = 300
    selx = lmx.reshape(2, -1)[:, :k]
    sely = lmy.reshape(2, -1)[:, 5:5 + k]
    sjx, sjy = jx.reshape(4, -1)[:, :k], jy.reshape(4, -1)[:, 5:5 + k]
    ok, b = host.SparseICA(img, key, selx, sely, sjx, sjy, T)
    ob_b = ob.sparse_ica(img, key, selx, sely, sjx, sjy, T)
    assert ok and np.allclose(b, ob_b, rtol=1e-12, atol=1e-9)     # f64 summation order only
    ok, warped = host.ImageWarp(img, T)
    assert ok and np.array_equal(warped, ob.image_warp(img, T))
    bgr = noise_image(rng, 90, 160, channels=3)
    assert np.array_equal(host.warpBySimilarityTransform(bgr, T), ob.warp_bgr(bgr, T))


def test_reference_image_warp_shift_kat(host):
    """align_test.cpp:358-400: a 10x10 white square at (20,20) of a 64x64 image warped by
    T{0,0,5,7}.inverse() moves by (+5,+7)."""
    img = np.zeros((64, 64), np.uint8)
    img[20:30, 20:30] = 255
    ok, out = host.ImageWarp(img, host.tf_inverse(np.array([0, 0, 5.0, 7.0])))
    assert ok
    ys, xs = np.nonzero(out > 127)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (25, 34, 27, 36)


@pytest.mark.parametrize("w,h,n,seed", [(320, 180, 9, 0), (640, 360, 5, 1)])
def test_video_aligner_class_matches_oracle(host, ob, w, h, n, seed):
    frames = _clip(w, h, n, seed)
    a, o = host.VideoAligner(0), ob.Aligner()
    for i, f in enumerate(frames):
        ok, T = a.AlignNextFrame(f)
        ok_o, T_o = o.align(f)
        assert ok == ok_o, i
        assert corner_displacement(T, T_o, w, h) <= TOL_PX, (i, T, T_o)
    # per-call params are honoured (alignment.hpp:55-58): max_iters=1 makes pairs fail like upstream
    from video_stabilizer_b200 import _capi as capi
    p = capi.VsAlignParams()
    capi.load().vs_align_params_default(p)
    p.max_iters = 1
    po = ob.align_params_default()
    po.max_iters = 1
    o.params = po
    ok, T = a.AlignNextFrame(frames[0], p)
    ok_o, T_o = o.align(frames[0])
    assert ok == ok_o and corner_displacement(T, T_o, w, h) <= TOL_PX


def test_video_aligner_class_with_phase_correlate(host, ob):
    """AlignNextFrame with VideoAlignerParams::phase_correlate through the drop-in class, against the restated
    aligner; the flag can be switched per call (alignment.hpp:55-58)."""
    from video_stabilizer_b200 import _capi as capi
    w, h = 320, 180
    frames = _clip(w, h, 9, 4, step=8.0, limit=40.0)
    p = capi.VsAlignParams()
    capi.load().vs_align_params_default(p)
    po = ob.align_params_default()
    a, o = host.VideoAligner(0), ob.Aligner(po)
    for i, f in enumerate(frames):
        p.phase_correlate = po.phase_correlate = 0 if i in (4, 5) else 1
        ok, T = a.AlignNextFrame(f, p)
        ok_o, T_o = o.align(f)
        assert ok == ok_o, i
        assert corner_displacement(T, T_o, w, h) <= 1e-6, (i, T, T_o)


def test_video_aligner_size_change_resets(host, ob):
    a, o = host.VideoAligner(0), ob.Aligner()
    seq = list(_clip(320, 180, 3, 1)) + list(_clip(256, 144, 3, 2)) + list(_clip(320, 180, 2, 3))
    for i, f in enumerate(seq):
        ok, T = a.AlignNextFrame(f)
        ok_o, T_o = o.align(f)
        assert ok == ok_o, i
        assert corner_displacement(T, T_o, f.shape[1], f.shape[0]) <= TOL_PX


def _run_oracle_stabilizer(ob, frames, p):
    po = ob.stab_params_default()
    po.lag, po.smoother_memory, po.lambda_, po.enable_smoother, po.crop_pixels = p.lag, p.smoother_memory, p.lambda_, p.enable_smoother, p.crop_pixels
    st = ob.Stabilizer(po)
    return [st.process(f)[0] for f in frames]


@pytest.mark.parametrize("crop,enable,lag", [(32, 1, 10), (0, 1, 10), (8, 0, 4)])
def test_video_stabilizer_class_matches_oracle(host, ob, crop, enable, lag):
    w, h, n = 320, 180, 30
    frames = _clip(w, h, n, 11, step=3.0)
    p = host.stab_params_default()
    p.crop_pixels, p.enable_smoother, p.lag = crop, enable, lag
    st = host.VideoStabilizer(p, 0)
    want = _run_oracle_stabilizer(ob, frames, p)
    produced = 0
    for i, f in enumerate(frames):
        got = st.processFrame(f)
        assert (got is None) == (want[i] is None), i
        if got is not None:
            produced += 1
            assert got.shape == want[i].shape == (h - 2 * crop, w - 2 * crop, 3)
            assert np.abs(got.astype(int) - want[i].astype(int)).max() <= TOL_LSB, i
    assert produced == n - lag


def test_video_stabilizer_survives_a_size_change(host, ob):
    p = host.stab_params_default()
    p.lag, p.smoother_memory, p.crop_pixels = 4, 2, 8
    seq = list(_clip(320, 180, 8, 1)) + list(_clip(256, 144, 8, 2))
    st = host.VideoStabilizer(p, 0)
    want = _run_oracle_stabilizer(ob, seq, p)
    for i, f in enumerate(seq):
        got = st.processFrame(f)
        assert (got is None) == (want[i] is None), i
        if got is not None:
            assert got.shape == want[i].shape, i
            assert np.abs(got.astype(int) - want[i].astype(int)).max() <= TOL_LSB, i


@pytest.mark.parametrize("chunks", [(40,), (7, 13, 1, 19), (16, 16, 8)])
def test_clip_stabilizer_equals_frame_by_frame(host, ob, chunks):
    """Batched feed() == n processFrame() calls (bit-exact, both on the GPU) == oracle (<= 1 LSB)."""
    w, h, n = 320, 180, sum(chunks)
    frames = _clip(w, h, n, 21, step=3.0)
    p = host.stab_params_default()
    seq = host.VideoStabilizer(p, 0)
    want_seq = [seq.processFrame(f) for f in frames]
    want_seq = [f for f in want_seq if f is not None]
    want_oracle = [f for f in _run_oracle_stabilizer(ob, frames, p) if f is not None]
    cs = host.ClipStabilizer(w, h, max(chunks), p, 0)
    got = []
    pos = 0
    for c in chunks:
        out = cs.feed(frames[pos:pos + c])
        got.extend(list(out))
        pos += c
    assert len(got) == len(want_seq) == n - p.lag
    for i, (g, s, o) in enumerate(zip(got, want_seq, want_oracle)):
        assert np.array_equal(g, s), ("batched vs sequential", i)
        assert np.abs(g.astype(int) - o.astype(int)).max() <= TOL_LSB, ("batched vs oracle", i)
    # a second video through the same object starts from scratch
    cs.reset()
    again = cs.feed(frames[:max(chunks)])
    assert len(again) == max(0, max(chunks) - p.lag)
    if len(again):
        assert np.array_equal(again[0], got[0])


@pytest.mark.parametrize("chunks,phase", [((160,), 0), ((139, 150), 0), ((130, 129, 131), 0), ((300,), 0), ((139, 150), 1), ((40, 7), 1)])
def test_clip_stabilizer_solver_lanes_equal_frame_by_frame(host, chunks, phase):
    """Device-resident chunks of >= 128 pairs run as 2-4 pieces on the clip's solver lanes (the solve of one piece
    beside the pyramids / warps of the others): the frames must still be those of n processFrame() calls, bit for bit,
    whatever the parity of the chunk's first frame."""
    import torch
    from video_stabilizer_b200 import _capi as capi
    w, h, n = 320, 180, sum(chunks)
    frames = _clip(w, h, n, 5, step=2.0)
    p = host.stab_params_default()
    p.aligner.phase_correlate = phase      # the seed kernels run on the context stream ahead of each lane's solve
    seq = host.VideoStabilizer(p, 0)
    want = [f for f in (seq.processFrame(f) for f in frames) if f is not None]
    cs = host.ClipStabilizer(w, h, max(chunks), p, 0)
    got = []
    pos = 0
    for c in chunks:
        part = np.ascontiguousarray(frames[pos:pos + c])
        out = torch.empty((c, cs.out_h, cs.out_w, 3), dtype=torch.uint8, device="cuda")
        k = cs.feed_ptr(part.ctypes.data, c, part.strides[1], part.strides[0], capi.VS_MEM_HOST, out.data_ptr(), capi.VS_MEM_DEVICE)
        cs.synchronize()
        got.extend(list(out[:k].cpu().numpy()))
        pos += c
    assert len(got) == len(want) == n - p.lag
    for i, (g, s) in enumerate(zip(got, want)):
        assert np.array_equal(g, s), i


def test_clip_stabilizer_records(host, ob):
    w, h, n = 320, 180, 16
    frames = _clip(w, h, n, 3)
    cs = host.ClipStabilizer(w, h, n, None, 0)
    out = cs.feed(frames)
    meas, ok, corr = cs.last_records(n)
    o = ob.Aligner()
    for i, f in enumerate(frames):
        ok_o, T_o = o.align(f)
        assert bool(ok[i]) == ok_o
        assert corner_displacement(meas[i], T_o, w, h) <= TOL_PX
    assert len(corr) == len(out) == n - 10


@pytest.mark.parametrize("workers,phase", [(1, 0), (2, 0), (3, 0), (2, 1)])
def test_multi_gpu_stabilizer_equals_single_stream(host, workers, phase):
    """Frame-chunk partition of one video over several workers (here: contexts on one device; on a
    multi-GPU box, one per device) with the transforms gathered to the host and the sequential
    smoother run there: the output equals the single-stream ClipStabilizer bit for bit."""
    from video_stabilizer_b200 import _capi as capi
    w, h, n = 320, 180, 45
    frames = _clip(w, h, n, 31, step=3.0)
    p = host.stab_params_default()
    p.crop_pixels = 8
    p.aligner.phase_correlate = phase
    cs = host.ClipStabilizer(w, h, n, p, 0)
    want = cs.feed(frames)
    meas_want, ok_want, _ = cs.last_records(n)
    ndev = max(1, capi.load().vs_device_count())
    devices = [i % ndev for i in range(workers)]
    mg = host.MultiGpuStabilizer(devices, w, h, n, p)
    got, meas, ok = mg.stabilize(frames)
    assert np.array_equal(meas, meas_want) and np.array_equal(ok, ok_want)
    assert got.shape == want.shape and np.array_equal(got, want)
    # a shorter second video through the same object
    got2, _, _ = mg.stabilize(frames[:20])
    assert np.array_equal(got2, cs_feed_fresh(host, frames[:20], p))


def cs_feed_fresh(host, frames, p):
    n, h, w, _ = frames.shape
    return host.ClipStabilizer(w, h, n, p, 0).feed(frames)


# ---------------------------------------------------------------- one video partitioned over several workers (partitioned.hpp)
def _partitioned_run(host, frames, p, world, sub, block, resident, name):
    """`world` workers as threads of this process (one context each; on a multi-GPU box one per device), the table in POSIX
    shared memory exactly as between processes.  Returns {frame: stabilized frame}, and the last worker's records."""
    import threading
    from video_stabilizer_b200 import _capi as capi
    n, h, w, _ = frames.shape
    ndev = max(1, capi.load().vs_device_count())
    workers = [host.PartitionedStabilizer(r, world, w, h, n, sub, block, p, name, resident, device=r % ndev, host_threads=2)
               for r in range(world)]          # rank 0 first: it creates the segment
    out, errors = {}, []

    def run(r):
        try:
            ps = workers[r]
            local = np.ascontiguousarray(frames[ps.local_frames])
            for video in range(2):          # twice: the second video uses the other copy of the table
                if resident:
                    if video == 0:
                        ps.upload_resident(local.ctypes.data, local.strides[1], local.strides[0])
                    got = ps.stabilize(None)
                else:
                    got = ps.stabilize(local)
                assert len(got) == ps.outputs
                if video == 1:
                    for f, g in zip(ps.output_frames, got):
                        out[int(f)] = g
        except Exception as e:       # noqa: BLE001
            errors.append((r, repr(e)))

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
    return out, workers[-1].records(n)


@pytest.mark.parametrize("world,sub,block,resident", [(1, 16, 3, True), (1, 12, 1, False), (2, 12, 2, True), (2, 12, 1, False),
                                                      (3, 10, 1, True), (3, 16, 1, False)])
def test_partitioned_stabilizer_equals_single_stream(host, world, sub, block, resident):
    """ONE video, frame-chunk partitioned over `world` workers (contiguous chunks when block * sub covers a worker's share,
    interleaved sub-chunks otherwise), device-resident and host-streamed: measurements, status and every stabilized frame
    equal the single-stream pipeline bit for bit; no collective, 65 bytes per frame through the shared host table."""
    w, h, n = 320, 180, 50
    frames = _clip(w, h, n, 41, step=3.0)
    p = host.stab_params_default()
    p.crop_pixels = 8
    cs = host.ClipStabilizer(w, h, n, p, 0)
    want = cs.feed(frames)
    meas_want, ok_want, corr_want = cs.last_records(n)
    name = "/vstab_gputest_%d_%d%d%d%d" % (os.getpid(), world, sub, block, int(resident))
    got, (corr, meas, ok) = _partitioned_run(host, frames, p, world, sub, block, resident, name if world > 1 else "")
    assert sorted(got) == list(range(n - p.lag))
    for f in range(n - p.lag):
        assert np.array_equal(got[f], want[f]), f
    seen = len(meas)
    assert seen >= n - 2 * sub
    assert np.array_equal(ok, ok_want[:seen])
    # the solver's f64 sums follow its CTA size (pairs in flight): ~1e-9 px, selections and iteration counts are identical
    for f in range(seen):
        assert corner_displacement(meas[f], meas_want[f], w, h) <= 1e-6, f
