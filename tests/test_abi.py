"""The C-ABI library loads without a GPU and exports every symbol include/vstab.h declares."""
import ctypes as C
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "vstab.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from video_stabilizer_b200 import _capi as capi
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libvstab.so does not export %s" % n
        assert n in capi.SYMBOLS, "%s is declared in vstab.h but not bound in _capi.py" % n
    assert lib.vs_abi_version() == 1


def test_structs_match_header_layout():
    from video_stabilizer_b200 import _capi as capi
    assert C.sizeof(capi.VsImg) == 40
    assert C.sizeof(capi.VsPair) == 12
    assert C.sizeof(capi.VsAlignParams) == 48
    p = capi.VsAlignParams()
    capi.load().vs_align_params_default(C.byref(p))
    assert (p.phase_correlate, p.threshold, p.max_iters, p.pyramid_min_width, p.max_displacement) == (0, 0.02, 64, 20, 10.0)
    assert abs(p.smallest_fraction - 0.8) < 1e-7


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a device vs_ctx_create fails with a message; nothing computes on the CPU."""
    from video_stabilizer_b200 import _capi as capi
    lib = capi.load()
    if lib.vs_device_count() > 0:
        return
    h = C.c_void_p()
    assert lib.vs_ctx_create(0, C.byref(h)) == -2
    assert b"no CPU fallback" in lib.vs_last_error(None)


def test_tile_size_rule():
    from video_stabilizer_b200 import _capi as capi
    lib = capi.load()
    assert [lib.vs_grad_argmax_tile_size(w, h) for (w, h) in ((1920, 1080), (960, 540), (480, 270), (240, 135), (120, 67), (60, 33))] == [20, 20, 10, 4, 2, 2]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "video_stabilizer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "vs_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
