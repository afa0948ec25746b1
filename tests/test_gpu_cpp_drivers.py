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
