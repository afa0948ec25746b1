"""The reference's driver programs (align_test, video_test) rebuilt from the public C++ headers
against the drop-in host library, run as processes: every check prints [PASS] and the exit code
is 0.  This is the C++ caller's view of the boundary (no Python in the loop)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(REPO, "video_stabilizer_b200", "bin")


def _run(args):
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    return r


def test_align_test_executable():
    r = _run([os.path.join(BIN, "align_test")])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "[FAIL]" not in r.stdout and "ALL CHECKS PASSED" in r.stdout
    assert r.stdout.count("[PASS]") >= 20


def test_video_test_executable():
    r = _run([os.path.join(BIN, "video_test"), "640", "360", "40"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "[PASS] batched output equals frame-by-frame output" in r.stdout


# ---------------------------------------------------------------- the reference's OWN drivers, compiled unmodified
# oracle/_ref/align_test_ref and video_test_ref are /root/reference/align_test.cpp and video_test.cpp compiled as they
# are against the product's drop-in headers and linked with libvstab_host.so / libvstab.so (oracle/Makefile: ref_drivers).
# Their fixtures (input.png, template.png, recordings/*.mp4) are not in the reference's repository: the tests write them
# as raw dumps (oracle/ref_shim/cv_io.cpp) and read the drivers' outputs back.
REF = os.path.join(REPO, "oracle", "_ref")


def _write_raw(path, frames):
    import numpy as np
    a = np.ascontiguousarray(frames, np.uint8)
    if a.ndim == 3:
        a = a[None]
    n, h, w, c = a.shape
    with open(path, "wb") as f:
        f.write(b"VSRAW1\n%d %d %d %d\n" % (w, h, c, n))
        f.write(a.tobytes())


def _read_raw(path):
    import numpy as np
    with open(path, "rb") as f:
        assert f.readline() == b"VSRAW1\n"
        w, h, c, n = [int(v) for v in f.readline().split()]
        data = np.frombuffer(f.read(), np.uint8)
    return data[: n * h * w * c].reshape(n, h, w, c)


def _need(exe):
    p = os.path.join(REF, exe)
    if not os.path.exists(p):
        pytest.skip("%s is built from /root/reference in the build container (oracle/Makefile ref_drivers)" % exe)
    return p


def test_reference_align_test_unmodified(tmp_path, ob):
    """align_test.cpp:43-247 (pyramid / gradients / ImageWarp shift checks / GradArgMax), :606-623 (transform algebra and
    the ImageWarp shift KAT) and :625-691 (AlignImagePair) run against the drop-in library; the transform it prints must
    be the oracle's for the same pair and aligned.png the oracle's ImageWarp of the input."""
    import re

    import numpy as np
    from video_stabilizer_b200 import synth
    exe = _need("align_test_ref")
    w, h = 1920, 1080                      # BASELINE.json configs[0]: a 1920x1080 pair with a known transform
    from oracle import binding
    canvas = synth.make_canvas(w, h, 1077)
    poses = synth.jitter_path(2, 1078, step=3.0)
    pair = [binding.warp_bgr_matrix(canvas, synth.forward_matrix_for_pose(p, w, h), w, h) for p in poses]
    cwd = tmp_path / "a" / "b"
    cwd.mkdir(parents=True)
    _write_raw(tmp_path / "a" / "input.png", pair[0])          # TestPyrDown reads ../input.png
    _write_raw(tmp_path / "template.png", pair[0])             # AlignImagePair reads ../../template.png, ../../input.png
    _write_raw(tmp_path / "input.png", pair[1])
    r = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=600)
    out = r.stdout
    print(out[-6000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "[FAIL]" not in out and "[FAIL]" not in r.stderr and "Error" not in r.stderr
    assert out.count("Shift verification passed.") == 6            # every pyramid level: ImageWarp by (4,4), found by phase correlation
    assert "[PASS] Warp shift matched expected transform" in out    # the reference's ImageWarp KAT
    assert out.count("[PASS]") >= 5
    m = re.search(r"Alignment successful\. Transform = (.*)", out)
    assert m, out[-2000:]
    got = [float(v) for v in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", m.group(1))]
    assert len(got) == 4
    o = ob.Aligner()
    assert o.align(pair[0])[0] is False
    ok, T = o.align(pair[1])
    assert ok
    from util import corner_displacement
    assert corner_displacement(got, T, w, h) <= 0.01 + 1e-4        # printed with 6 significant digits
    aligned = _read_raw(tmp_path / "aligned.png")[0, :, :, 0]
    want = ob.image_warp(ob.bgr2gray(pair[1]), T)
    # the driver rounds the float image to 8 bits (convertTo); the printed-vs-exact transform moves a value by << 1 LSB
    assert np.abs(aligned.astype(int) - np.clip(np.rint(want), 0, 255).astype(int)).max() <= 1


def test_reference_video_test_unmodified(tmp_path, ob):
    """video_test.cpp:10-128 run as it is over recordings/*.mp4 (raw dumps): VideoCapture -> VideoStabilizer(crop 0)
    ::processFrame per frame -> VideoWriter.  Every frame it writes must be the oracle stabilizer's (<= 1 LSB)."""
    import numpy as np
    from video_stabilizer_b200 import synth
    exe = _need("video_test_ref")
    w, h, n = 640, 360, 48
    frames, _ = synth.make_clip_numpy(w, h, n, 55, step=3.0)
    (tmp_path / "recordings").mkdir()
    (tmp_path / "build").mkdir()
    _write_raw(tmp_path / "recordings" / "clip.mp4", frames)
    r = subprocess.run([exe], cwd=tmp_path / "build", capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "All videos have been processed successfully." in r.stdout
    got = _read_raw(tmp_path / "build" / "output" / "processed_clip.mp4")
    po = ob.stab_params_default()
    po.crop_pixels = 0                                    # video_test.cpp:54
    st = ob.Stabilizer(po)
    want = [f for f in (st.process(f)[0] for f in frames) if f is not None]
    assert len(got) == len(want) == n - po.lag
    for i, (g, o) in enumerate(zip(got, want)):
        assert np.abs(g.astype(int) - o.astype(int)).max() <= 1, i
