"""Quality tooling on the batched API: the parameter sweep of grid_search_align.cpp:134-210 as one
solver launch, and the jitter score in the spirit of eval_jitter.cpp:21-75 / measure_jitter (grid_search_align.cpp:27-60)."""
import ctypes as C

import numpy as np
import pytest

from util import corner_displacement

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(gpu):
    from video_stabilizer_b200 import host
    host.load()
    return host


def _clip(w, h, n, seed, **kw):
    from video_stabilizer_b200 import synth
    return synth.make_clip_numpy(w, h, n, seed, **kw)[0]


def test_align_sweep_equals_one_launch_per_parameter_set(gpu, ob):
    """vs_clip_align_sweep(pairs x sets) == vs_clip_set_params + vs_clip_align per set, bit for bit, and == the oracle run with
    those VideoAlignerParams (status, <= 0.01 px); includes sets that make pairs fail (max_displacement, max_iters)."""
    from video_stabilizer_b200 import _capi as capi
    from video_stabilizer_b200.clip import Clip, pairs_for_frames
    w, h, n = 320, 180, 9
    frames = _clip(w, h, n, 61, step=4.0)
    sets = [(0.02, 10.0, 0.8, 64, 0), (0.05, 6.0, 0.3, 64, 0), (0.03, 8.0, 0.5, 64, 1), (0.02, 1.5, 0.8, 64, 0),
            (0.0005, 10.0, 0.8, 3, 0), (0.02, 10.0, 1.0, 64, 1), (0.02, 10.0, 0.0, 64, 0)]
    clip = Clip(w, h, n, max_pairs=(n - 1) * len(sets), ctx=gpu)
    clip.upload(0, frames)
    clip.build_pyramids(0, n)
    pairs, keys = pairs_for_frames(0, n)
    clip.build_keyframes(keys)
    arr = (capi.VsSweepParams * len(sets))(*[capi.VsSweepParams(t, md, fr, it, pc) for (t, md, fr, it, pc) in sets])
    T = np.zeros((len(sets), n - 1, 4))
    st = np.zeros((len(sets), n - 1), np.int32)
    capi.check(gpu.handle, gpu.lib.vs_clip_align_sweep(clip.handle, C.cast(pairs, C.c_void_p), n - 1, C.cast(arr, C.c_void_p), len(sets),
                                                       capi.ptr(T), capi.ptr(st), capi.VS_MEM_HOST), "vs_clip_align_sweep")
    for s, (t, md, fr, it, pc) in enumerate(sets):
        p = capi.VsAlignParams()
        gpu.lib.vs_align_params_default(C.byref(p))
        p.threshold, p.max_displacement, p.smallest_fraction, p.max_iters, p.phase_correlate = t, md, fr, it, pc
        capi.check(gpu.handle, gpu.lib.vs_clip_set_params(clip.handle, C.byref(p)), "set_params")
        T1, st1, _ = clip.align(pairs)
        assert np.array_equal(st[s], st1), s
        assert np.array_equal(T[s], T1), s
        po = ob.align_params_default()
        po.threshold, po.max_displacement, po.smallest_fraction, po.max_iters, po.phase_correlate = t, md, fr, it, pc
        o = ob.Aligner(po)
        for f in range(n):
            ok, To = o.align(frames[f])
            if f == 0:
                continue
            assert bool(st[s, f - 1]) == ok, (s, f)
            if ok:
                assert corner_displacement(T[s, f - 1], To, w, h) <= 0.01, (s, f)
    assert (st == 0).any() and (st == 1).any()


def test_jitter_score_and_grid_search(host, ob):
    """The batched sweep's per-combination measurements are the ones a VideoStabilizer with those parameters measures; the
    jitter of a clip is the median flow magnitude of its frame-to-frame motion; stabilizing lowers it."""
    w, h, n = 320, 180, 24
    frames = _clip(w, h, n, 71, step=3.0)
    gs = host.AlignerGridSearch(w, h, n, 54, crop_pixels=16)
    jit = gs.measure_jitter(frames)
    # the same statistic from the oracle's measurements
    o = ob.Aligner()
    meds = []
    for f in range(n):
        ok, T = o.align(frames[f])
        if f and ok:
            meds.append(host.flow_median_px(T, w, h))
    want = float(np.median(meds))
    assert jit["pairs"] == n - 1 and jit["failed"] == n - 1 - len(meds)
    assert abs(jit["median_px"] - want) <= 0.01
    assert 0.5 < jit["median_px"] < 6.0                      # the synthetic walk moves up to 3 px per axis per frame

    grid = host.reference_grid()
    assert grid.shape == (54, 4)
    inp, res, T, st, launches = gs.run(frames, grid)
    assert abs(inp["median_px"] - jit["median_px"]) < 1e-12
    assert res.shape == (54, 5) and launches > 0
    # combination 8 = {pc 0, thr 0.02, frac 0.8, maxDisp 10} is the default VideoAlignerParams: its measurements are the oracle's
    d = np.where((grid[:, 0] == 0) & (grid[:, 1] == 0.02) & (np.abs(grid[:, 2] - 0.8) < 1e-6) & (grid[:, 3] == 10.0))[0]
    assert len(d) == 1
    o = ob.Aligner()
    for f in range(n):
        ok, To = o.align(frames[f])
        if f:
            assert bool(st[d[0], f - 1]) == ok
            if ok:
                assert corner_displacement(T[d[0], f - 1], To, w, h) <= 0.01
    # removing the measured motion frame by frame (smoother off, lag 1) leaves far less jitter than the input has
    good = res[:, 3] == 0
    assert good.any()
    assert (res[good, 1] < 0.9).all() and res[good, 1].min() < 0.3, res[good][:, :2]
    assert res[d[0], 4] == n - 2
