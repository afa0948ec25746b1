"""Full-size (BASELINE.json shapes) checks through size-independent properties, where running the
CPU oracle on every frame would take minutes: batched == frame-by-frame (bit-exact), identity and
integer-shift warps, keypoints inside their tiles, known synthetic motion recovered, and oracle
spot checks on a few frames of the same clip."""
import numpy as np
import pytest

from util import corner_displacement

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(gpu):
    from video_stabilizer_b200 import host
    host.load()
    return host


@pytest.fixture(scope="module")
def clip1080(gpu):
    from video_stabilizer_b200 import synth
    frames, poses = synth.make_clip_gpu(gpu, 1920, 1080, 48, 77, chunk=24)
    return frames, poses


def test_1080p_batched_pipelined_equals_frame_by_frame(host, clip1080):
    frames, _ = clip1080
    p = host.stab_params_default()
    p.crop_pixels = 0                                    # video_test.cpp:54
    seq = host.VideoStabilizer(p, 0)
    want = [f for f in (seq.processFrame(f) for f in frames) if f is not None]
    cs = host.ClipStabilizer(1920, 1080, len(frames), p, 0)
    got = cs.feed(frames)                                 # host -> host: the 3-stream pipelined path (48 > 32 frames)
    assert len(got) == len(want) == len(frames) - p.lag
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), i
    meas, ok, corr = cs.last_records(len(frames))
    assert ok[1:].all() and not ok[0]


def test_1080p_recovers_the_synthetic_motion_and_matches_oracle_spot_checks(gpu, ob, clip1080):
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    frames, poses = clip1080
    n = 12
    clip = Clip(1920, 1080, n, ctx=gpu)
    clip.upload(0, frames[:n])
    clip.build_pyramids(0, n)
    pairs, keys = pairs_for_frames(0, n)
    clip.build_keyframes(keys)
    T, st, it = clip.align(pairs)
    assert st.all()
    for i in range(1, n):
        truth = ob.tf_compose(ob.tf_inverse(poses[i]), poses[i - 1])
        assert corner_displacement(T[i - 1], truth, 1920, 1080) < 0.5, i   # the reference's quarter-step GN stops ~0.1 px short
    # oracle on the first three frames only (seconds, not minutes)
    al = ob.Aligner()
    for i in range(3):
        ok, To = al.align(frames[i])
        if i:
            assert ok and corner_displacement(T[i - 1], To, 1920, 1080) <= 0.01
    # structural property at every level: the keypoint of tile (tx,ty) lies inside that tile
    for l in range(clip.levels):
        li = clip.level_info(l)
        for axis in range(2):
            kp = clip.get_keypoints(1, l, axis).astype(int)
            tx = np.arange(li["tw"])[None, :] * li["tile"]
            ty = np.arange(li["th"])[:, None] * li["tile"]
            assert ((kp[0] >= tx) & (kp[0] < tx + li["tile"]) & (kp[1] >= ty) & (kp[1] < ty + li["tile"])).all(), (l, axis)
    clip.close()


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160)])
def test_warp_identity_and_integer_shift_are_exact(gpu, w, h):
    from video_stabilizer_b200.clip import Clip
    rng = np.random.default_rng(w)
    frame = rng.integers(0, 256, (1, h, w, 3), dtype=np.uint8)
    clip = Clip(w, h, 1, ctx=gpu)
    clip.upload(0, frame)
    out = clip.warp([0], np.zeros((1, 4)))
    assert np.array_equal(out[0], frame[0])                                   # identity is a copy
    out = clip.warp([0], np.array([[0.0, 0.0, 7.0, -3.0]]))                    # content moves by (+7, -3), zeros enter
    assert np.array_equal(out[0, : h - 3, 7:], frame[0, 3:, : w - 7])
    assert not out[0, h - 3:, :].any() and not out[0, :, :7].any()
    out = clip.warp([0], np.array([[0.0, 0.0, 7.0, -3.0]]), crop=32)           # the crop is a window of the same warp
    full = clip.warp([0], np.array([[0.0, 0.0, 7.0, -3.0]]))
    assert np.array_equal(out[0], full[0, 32: h - 32, 32: w - 32])
    clip.close()


def test_pyramid_of_a_constant_frame_is_constant_at_4k(gpu):
    from video_stabilizer_b200.clip import Clip
    clip = Clip(3840, 2160, 1, ctx=gpu)
    clip.upload(0, np.full((1, 2160, 3840, 3), 201, np.uint8))
    clip.build_pyramids(0, 1)
    assert clip.levels == 7
    for l in range(clip.levels):
        g = clip.get_gray(0, l)
        assert g.min() == g.max() == 201, l
    clip.close()
