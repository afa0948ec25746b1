import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle (test infrastructure)."""
    from oracle import binding
    binding.load()
    return binding


@pytest.fixture(scope="session")
def hooks():
    """CPU build of the product's host+device headers (introselect replay, 4x4 SVD)."""
    import ctypes as C
    d = os.path.join(REPO, "tests", "native")
    so = os.path.join(d, "libhooks.so")
    src = os.path.join(d, "hooks.cpp")
    csrc = os.path.join(REPO, "video_stabilizer_b200", "csrc")
    deps = [src, os.path.join(csrc, "vs_introselect.cuh"), os.path.join(csrc, "vs_linalg4.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", csrc, "-o", so, src],
                       check=True)
    lib = C.CDLL(so)
    lib.th_condition_and_invert.restype = C.c_double
    return lib


@pytest.fixture(scope="session")
def gpu():
    """A vs_ctx on cuda:0.  Fails (does not skip) when the library or device is missing."""
    from video_stabilizer_b200.imgproc import Context
    return Context(0)
