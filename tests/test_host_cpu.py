"""Host C++ layer (libvstab_host.so) without a GPU: it loads, exports every symbol
include/vstab_host.h declares, reproduces the reference's transform-algebra known-answer
tests, and its smoother / trajectory match the oracle (and, through it, the reference's own
smoother.cpp / stabilizer.cpp — tests/test_oracle_vs_ref.py)."""
import os
import re

import numpy as np
import pytest

from mt19937 import MT19937

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    from video_stabilizer_b200 import host
    host.load()
    return host


def test_host_header_symbols_are_exported_and_bound(host):
    text = open(os.path.join(REPO, "include", "vstab_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(vsh_[A-Za-z0-9_]+)\s*\(", text)))
    assert len(names) >= 35
    lib = host.load()
    for n in names:
        assert hasattr(lib, n), "libvstab_host.so does not export %s" % n
        assert n in host.SYMBOLS, "%s is declared in vstab_host.h but not bound in host.py" % n


def test_params_defaults_match_reference_headers(host):
    p = host.stab_params_default()   # stabilizer.hpp:13-30, alignment.hpp:5-41
    assert (p.lag, p.smoother_memory, p.lambda_, p.enable_smoother, p.crop_pixels) == (10, 5, 4.0, 1, 32)
    assert (p.min_disp, p.max_disp, p.min_decay, p.max_decay) == (48.0, 64.0, 0.9, 0.7)
    assert (p.aligner.threshold, p.aligner.max_iters, p.aligner.max_displacement) == (0.02, 64, 10.0)


def test_transform_algebra_matches_oracle_bitwise(host, ob):
    rng = np.random.default_rng(1)
    for _ in range(300):
        T1 = rng.uniform(-1, 1, 4) * np.array([0.3, 0.3, 80, 80])
        T2 = rng.uniform(-1, 1, 4) * np.array([0.3, 0.3, 80, 80])
        assert np.array_equal(host.tf_inverse(T1), ob.tf_inverse(T1))
        assert np.array_equal(host.tf_compose(T1, T2), ob.tf_compose(T1, T2))
        assert np.array_equal(host.tf_warp(T1, 12.5, -7.0), ob.tf_warp(T1, 12.5, -7.0))
        assert np.array_equal(host.tf_warp(T1, 12.5, -7.0, (640, 360)), ob.tf_warp(T1, 12.5, -7.0, (640, 360)))
        assert host.tf_max_corner_displacement(T1, 1280, 720) == ob.tf_max_corner_displacement(T1, 1280, 720)


def test_reference_randomized_inverse_kat(host):
    """align_test.cpp:444-480 (TestRandomizedInverse): seed 12345, T(Tinv(p)) == p within 1e-5."""
    g = MT19937(12345)
    for _ in range(50):
        T = np.array([g.uniform(-0.5, 0.5), g.uniform(-0.5, 0.5), g.uniform(-100, 100), g.uniform(-100, 100)])
        Ti = host.tf_inverse(T)
        for _ in range(10):
            x, y = g.uniform(-1000, 1000), g.uniform(-1000, 1000)
            q = host.tf_warp(T, x, y)
            back = host.tf_warp(Ti, q[0], q[1])
            assert abs(np.float32(back[0]) - np.float32(x)) <= 1e-3 and abs(np.float32(back[1]) - np.float32(y)) <= 1e-3
            assert abs(back[0] - x) < 1e-9 * max(1, abs(x)) + 1e-9 and abs(back[1] - y) < 1e-9 * max(1, abs(y)) + 1e-9


def test_compose_is_this_then_argument_and_associative(host):
    """align_test.cpp:487-551 (TestRandomizedCompose): compose order and associativity."""
    g = MT19937(6789)
    for _ in range(50):
        Ts = [np.array([g.uniform(-0.2, 0.2), g.uniform(-0.2, 0.2), g.uniform(-50, 50), g.uniform(-50, 50)]) for _ in range(3)]
        T12 = host.tf_compose(Ts[0], Ts[1])
        x, y = g.uniform(-500, 500), g.uniform(-500, 500)
        p1 = host.tf_warp(Ts[0], x, y)
        p2 = host.tf_warp(Ts[1], p1[0], p1[1])
        q = host.tf_warp(T12, x, y)
        assert np.allclose(q, p2, atol=1e-9)
        left = host.tf_compose(host.tf_compose(Ts[0], Ts[1]), Ts[2])
        right = host.tf_compose(Ts[0], host.tf_compose(Ts[1], Ts[2]))
        assert np.allclose(left, right, atol=1e-9)
    ident = host.tf_compose(Ts[0], host.tf_inverse(Ts[0]))   # align_test.cpp:557-601
    assert np.allclose(ident, 0, atol=1e-12)


def test_smoother_matches_oracle_bitwise(host, ob):
    rng = np.random.default_rng(2)
    d = rng.normal(0, 5, 16)
    assert np.array_equal(host.tvl1_relax(d, 4.0), ob.tvl1_smooth(d, 4.0))
    assert np.array_equal(host.tvl1_relax(d[:1], 4.0), ob.tvl1_smooth(d[:1], 4.0))
    a, b = host.L1SmootherCenter(10, 5, 4.0), ob.Smoother(10, 5, 4.0)
    for i in range(80):
        m = rng.normal(0, 1, 4) * np.array([0.002, 0.002, 6.0, 6.0])
        ra, rb = a.update(m), b.update(m)
        assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1]), i


@pytest.mark.parametrize("enable,lag,mem", [(1, 10, 5), (0, 10, 5), (1, 4, 2), (1, 3, 6)])
def test_trajectory_matches_reference_stabilizer_glue(host, ob, enable, lag, mem):
    """StabilizerTrajectory against the same recurrence written out with the oracle's
    smoother and algebra (stabilizer.cpp:19-88), including resets on failed alignments and
    the decay branches."""
    rng = np.random.default_rng(5)
    p = host.stab_params_default()
    p.enable_smoother, p.lag, p.smoother_memory = enable, lag, mem
    traj = host.StabilizerTrajectory(p)
    sm = ob.Smoother(lag, mem, p.lambda_)
    accum = np.zeros(4)
    queue = []
    w, h = 1280, 720
    dues = 0
    for i in range(120):
        scale = 25.0 if 40 <= i < 60 else 4.0            # drive displacement through both decay branches
        meas = rng.normal(0, 1, 4) * np.array([0.003, 0.003, scale, scale])
        ok = not (i == 0 or i % 37 == 36)
        due, corr = traj.push(meas, ok, w, h)
        smoothed = np.zeros(4)
        if enable:
            f, s = sm.update(meas)
            if f:
                smoothed = s
        if not ok:
            accum = np.zeros(4)
        queue.append(meas)
        if len(queue) > lag:
            oldest = queue.pop(0)
            jitter = ob.tf_compose(oldest, ob.tf_inverse(smoothed)) if enable else oldest
            na = ob.tf_compose(accum, jitter)
            disp = ob.tf_max_corner_displacement(na, w, h)
            if disp > p.max_disp:
                decay = p.max_decay
            elif disp > p.min_disp:
                f = min(1.0, max(0.0, (disp - p.min_disp) / (p.max_disp - p.min_disp)))
                decay = p.min_decay * (1.0 - f) + p.max_decay * f
            else:
                decay = p.min_decay
            na = np.array([na[0] * decay, na[1] * decay, na[2] * decay, na[3] * decay])
            accum = na
            assert due and np.array_equal(corr, ob.tf_inverse(na)), i
            dues += 1
        else:
            assert not due
    assert dues == 120 - lag


def test_flow_median_of_a_translation_and_a_rotation():
    """jitter statistic (grid_search.hpp): median |W(p) - p| over a grid; a pure translation moves every point alike"""
    from video_stabilizer_b200 import host
    host.load()
    assert abs(host.flow_median_px([0, 0, 3, 4], 1920, 1080) - 5.0) < 1e-12
    assert host.flow_median_px([0, 0, 0, 0], 1920, 1080) == 0.0
    r = host.flow_median_px([0, 0.001, 0, 0], 1920, 1080)       # small rotation about the centre: grows with the radius
    assert 0.2 < r < 1.2
    g = host.reference_grid()
    assert g.shape == (54, 4) and g[0].tolist() == [0.0, 0.02, np.float32(0.3), 6.0] and g[-1, 0] == 1.0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own host sources on the host cores) prints ONE JSON line with the keys the
    driver reads, on the GPU arm's metric / unit / config; and the GPU arm refuses to run without a device."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--width", "320", "--height", "180", "--frames", "20", "--cpu-threads", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "stabilized_frames_per_sec_1080p" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["width"] == 320 and "workload" in d["config"]
    import torch
    if not torch.cuda.is_available():
        gpu = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--steps", "1", "--warmup", "0"],
                             capture_output=True, text=True, timeout=300)
        assert gpu.returncode != 0 and "no CUDA device" in (gpu.stderr + gpu.stdout)
