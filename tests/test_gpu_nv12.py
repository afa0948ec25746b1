"""NV12 frames (decoder output fed straight in; VS_CLIP_NV12, vs_plane_warp_u8).

Upstream has no NV12 path (its frames are BGR cv::Mat), so the definition is pinned where it can be: the plane warps against
cv2.warpAffine on 1- and 2-channel images (tests/test_oracle_golden.py::test_nv12_plane_warps_match_cv2, CPU), the alignment
against the reference's alignment of the BGR frame whose three channels equal Y (cv::cvtColor(BGR2GRAY) of it is Y).
"""
import numpy as np
import pytest

from util import corner_displacement

pytestmark = pytest.mark.gpu


def forward_matrix(T, w, h):
    A, B, TX, TY = T
    cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
    return np.array([1.0 + A, -B, TX - A * cx + B * cy, B, 1.0 + A, TY - B * cx - A * cy], np.float64)


def nv12_clip(ob, w, h, n, seed):
    """NV12 frames of a synthetic jittered clip: Y = BGR2GRAY of the rendered frame, U / V from its blue / red channels."""
    from video_stabilizer_b200 import synth
    frames, poses = synth.make_clip_numpy(w, h, n, seed)
    out = np.empty((n, h * 3 // 2, w), np.uint8)
    for i, f in enumerate(frames):
        out[i, :h] = ob.bgr2gray(f)
        uv = out[i, h:].reshape(h // 2, w // 2, 2)
        uv[:, :, 0] = f[::2, ::2, 0] // 2 + 64
        uv[:, :, 1] = 255 - f[1::2, 1::2, 2] // 2
    return out, poses


@pytest.mark.parametrize("h,w", [(97, 131), (62, 86), (360, 640), (1080, 1920)])
@pytest.mark.parametrize("ch", [1, 2])
def test_plane_warp_operator(gpu, ob, h, w, ch):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h * 7 + ch)
    src = rng.integers(0, 256, (h, w) if ch == 1 else (h, w, ch), dtype=np.uint8)
    cases = [(0, 0, 0, 0), (0, 0, 5, -3), (0.013, -0.021, 3.37, -2.81), (-0.2, 0.1, -30.5, 12.25)]
    if h > 400:
        cases = cases[2:3]
    for T in cases:
        M = forward_matrix(T, w, h)
        assert np.array_equal(ip.PlaneWarp(src, M, ctx=gpu), ob.warp_plane_matrix(src, M, w, h)), (T, "full")
        # a window of the output (crop fused), odd extents: the tail columns and unaligned rows
        ow, oh = w - 13, h - 9
        assert np.array_equal(ip.PlaneWarp(src, M, ow, oh, 6, 4, ctx=gpu), ob.warp_plane_matrix(src, M, ow, oh, 6, 4)), (T, "window")


def test_plane_warp_errors_are_reported(gpu):
    import ctypes as C
    from video_stabilizer_b200 import _capi as capi
    lib = capi.load()
    a = np.zeros((8, 8, 3), np.uint8)
    M = np.array([1, 0, 0, 0, 1, 0], np.float64)
    simg = capi.VsImg(a.ctypes.data, 8, 8, a.strides[0], 1, 0)
    assert lib.vs_plane_warp_u8(gpu.handle, C.byref(simg), 3, capi.ptr(M), C.byref(simg), 0, 0, capi.VS_MEM_HOST) == -1
    assert b"channels" in lib.vs_last_error(gpu.handle)
    assert lib.vs_plane_warp_u8(gpu.handle, C.byref(simg), 1, None, C.byref(simg), 0, 0, capi.VS_MEM_HOST) == -1


@pytest.mark.parametrize("w,h,n,seed", [(320, 180, 8, 0), (640, 360, 5, 1), (1280, 720, 3, 3)])
def test_nv12_clip_alignment_is_that_of_the_gray_frames(gpu, ob, w, h, n, seed):
    """Pyramids, transforms, status and iteration counts of an NV12 clip equal those of the restated VideoAligner fed the BGR
    frames whose channels all equal Y, and bit for bit those of a BGR clip of the same frames."""
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    nv12, _ = nv12_clip(ob, w, h, n, seed)
    gray_bgr = np.repeat(nv12[:, :h, :, None], 3, 3)
    clip = Clip(w, h, n, ctx=gpu, nv12=True)
    clip.upload(0, nv12)
    clip.build_pyramids(0, n)
    pairs, keyframes = pairs_for_frames(0, n)
    clip.build_keyframes(keyframes)
    T, status, iters = clip.align(pairs)
    ref = Clip(w, h, n, ctx=gpu)
    ref.upload(0, gray_bgr)
    ref.build_pyramids(0, n)
    ref.build_keyframes(keyframes)
    T2, status2, iters2 = ref.align(pairs)
    assert np.array_equal(T, T2) and np.array_equal(status, status2) and np.array_equal(iters, iters2)
    for l in range(clip.levels):
        assert np.array_equal(clip.get_gray(n - 1, l), ref.get_gray(n - 1, l)), l
    assert np.array_equal(clip.get_bgr(2), nv12[2])
    al = ob.Aligner(None)
    for i, f in enumerate(gray_bgr):
        ok, To = al.align(f)
        if i == 0:
            continue
        assert bool(status[i - 1]) == ok
        assert [al.iterations(l) for l in range(al.levels)] == list(iters[i - 1])
        assert corner_displacement(T[i - 1], To, w, h) < 1e-6
    assert status.sum() >= n - 2
    clip.close(); ref.close()


@pytest.mark.parametrize("w,h,crop", [(320, 180, 0), (322, 182, 4), (1920, 1080, 0), (1920, 1080, 32)])
def test_nv12_clip_warp_matches_oracle(gpu, ob, w, h, crop):
    from video_stabilizer_b200.clip import Clip
    rng = np.random.default_rng(w + crop)
    n = 3
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    clip = Clip(w, h, n, ctx=gpu, nv12=True)
    clip.upload(0, nv12)
    T = np.array([[0, 0, 0, 0], [0.004, -0.003, 5.3, -3.7], [-0.02, 0.015, -21.25, 14.5]], np.float64)
    out = clip.warp([2, 0, 1], T, crop=crop)
    for i, s in enumerate([2, 0, 1]):
        assert np.array_equal(out[i], ob.warp_nv12(nv12[s], w, h, T[i], crop)), (i, s)
    if crop == 0:
        assert np.array_equal(out[0], nv12[2])          # the identity returns the frame
    clip.close()


def test_nv12_clip_rejects_what_it_cannot_do(gpu):
    import ctypes as C
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.clip import Clip
    lib = capi.load()
    h = C.c_void_p()
    assert lib.vs_clip_create(gpu.handle, 321, 180, 2, 2, None, capi.VS_CLIP_NV12, C.byref(h)) == -1      # odd width
    clip = Clip(320, 180, 2, ctx=gpu, nv12=True)
    T = np.zeros((1, 4))
    with pytest.raises(capi.VsError):
        clip.warp([0], T, crop=3)                       # odd crop
    with pytest.raises(capi.VsError):
        clip.warp([0], T, mode=capi.VS_WARP_LANCZOS2)   # BGR-only modes
    clip.close()


# ---------------------------------------------------------------- the batched host classes on NV12 frames
@pytest.fixture(scope="module")
def host(gpu):
    from video_stabilizer_b200 import host
    host.load()
    return host


@pytest.mark.parametrize("chunks,crop", [((40,), 8), ((7, 13, 1, 19), 0), ((48,), 32)])
def test_clip_stabilizer_nv12(host, ob, chunks, crop):
    """ClipStabilizer on NV12 frames: the measurements are those of the BGR pipeline fed the gray frames (Y in every channel),
    the corrections follow, and every produced frame is the oracle's plane-by-plane warp of its NV12 frame."""
    w, h, n = 320, 180, sum(chunks)
    nv12, _ = nv12_clip(ob, w, h, n, 21)
    gray_bgr = np.repeat(nv12[:, :h, :, None], 3, 3)
    p = host.stab_params_default()
    p.crop_pixels = crop
    ref = host.ClipStabilizer(w, h, n, p, 0)
    ref.feed(gray_bgr)
    meas_want, ok_want, corr_want = ref.last_records(n)
    cs = host.ClipStabilizer(w, h, max(chunks), p, 0, nv12=True)
    cs.set_pipeline_frames(8)            # the host-to-host pipeline: asynchronous uploads into pyramid level 0
    got, corr = [], []
    pos = 0
    for c in chunks:
        out = cs.feed(nv12[pos:pos + c])
        got.extend(list(out))
        corr.extend(list(cs.last_records(c)[2]))
        pos += c
    assert len(got) == len(corr_want) == n - p.lag
    assert got[0].shape == ((h - 2 * crop) * 3 // 2, w - 2 * crop)
    for k in range(len(got)):
        assert corner_displacement(corr[k], corr_want[k], w, h) <= 1e-6, k
        assert np.array_equal(got[k], ob.warp_nv12(nv12[k], w, h, corr[k], crop)), k


@pytest.mark.parametrize("world,sub,block,resident", [(1, 16, 3, True), (2, 12, 1, False), (3, 10, 1, True)])
def test_partitioned_stabilizer_nv12_equals_single_stream(host, world, sub, block, resident):
    """ONE NV12 video frame-chunk partitioned over `world` workers equals the single-stream NV12 pipeline bit for bit."""
    import os
    import threading
    from video_stabilizer_b200 import _capi as capi
    from oracle import binding as ob
    w, h, n = 320, 180, 50
    nv12, _ = nv12_clip(ob, w, h, n, 41)
    p = host.stab_params_default()
    p.crop_pixels = 8
    want = host.ClipStabilizer(w, h, n, p, 0, nv12=True).feed(nv12)
    name = "/vstab_gputest_nv12_%d_%d%d" % (os.getpid(), world, sub) if world > 1 else ""
    ndev = max(1, capi.load().vs_device_count())
    workers = [host.PartitionedStabilizer(r, world, w, h, n, sub, block, p, name, resident, device=r % ndev, host_threads=2, nv12=True)
               for r in range(world)]
    out, errors = {}, []

    def run(r):
        try:
            ps = workers[r]
            local = np.ascontiguousarray(nv12[ps.local_frames])
            if resident:
                ps.upload_resident(local.ctypes.data, local.strides[1], local.strides[0])
                got = ps.stabilize(None)
            else:
                got = ps.stabilize(local)
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
    assert sorted(out) == list(range(n - p.lag))
    for f in range(n - p.lag):
        assert np.array_equal(out[f], want[f]), f
