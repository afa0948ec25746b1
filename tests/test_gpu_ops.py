"""Parity of every single operator of the C ABI against the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from util import noise_image

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (2, 3), (5, 4), (33, 17), (64, 64), (135, 240), (131, 97), (360, 640), (67, 120), (1080, 1920)]


@pytest.mark.parametrize("h,w", SIZES)
def test_bgr2gray(gpu, ob, h, w):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h * 10007 + w)
    bgr = noise_image(rng, h, w, 3)
    assert np.array_equal(ip.BGR2Gray(bgr, gpu), ob.bgr2gray(bgr))


# the fused kernel runs for widths that are multiples of 16 (strips of 72 rows: sizes around the strip boundaries, odd
# heights, one and several column warps, partial warps); the other sizes take the two-kernel path behind the same entry
INGEST_SIZES = [(2, 16), (3, 16), (72, 32), (73, 48), (74, 512), (75, 528), (143, 1040), (144, 16), (145, 64), (146, 80),
                (180, 320), (360, 640), (720, 1280), (1080, 1920), (1081, 1936), (2160, 3840), (135, 240), (131, 97), (67, 120)]


@pytest.mark.parametrize("h,w", INGEST_SIZES)
def test_ingest_bgr_gray_l1(gpu, ob, h, w):
    """BGR -> gray + first pyramid level in one kernel == cvtColor then pyr_down of the oracle, bit for bit."""
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h * 4099 + w)
    bgr = noise_image(rng, h, w, 3)
    g0, g1 = ip.IngestBGR(bgr, gpu)
    want0 = ob.bgr2gray(bgr)
    assert np.array_equal(g0, want0)
    assert np.array_equal(g1, ob.pyr_down(want0, w // 2, h // 2))


def test_ingest_extreme_values(gpu, ob):
    """saturated channels exercise the top of the 24-bit gray accumulator and of the 16-bit pyramid sums"""
    from video_stabilizer_b200 import imgproc as ip
    h, w = 150, 96
    for fill in ((255, 255, 255), (0, 0, 0), (255, 0, 0), (0, 255, 0), (0, 0, 255)):
        bgr = np.empty((h, w, 3), np.uint8)
        bgr[:] = fill
        bgr[::7, ::5] = 255 - np.array(fill, np.uint8)
        g0, g1 = ip.IngestBGR(bgr, gpu)
        want0 = ob.bgr2gray(bgr)
        assert np.array_equal(g0, want0) and np.array_equal(g1, ob.pyr_down(want0, w // 2, h // 2))


@pytest.mark.parametrize("h,w", SIZES)
def test_pyr_down(gpu, ob, h, w):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h * 7 + w)
    img = noise_image(rng, h, w)
    oh, ow = max(h // 2, 1), max(w // 2, 1)
    out = np.zeros((oh, ow), np.uint8)
    assert ip.PyrDown(img, out, gpu)
    assert np.array_equal(out, ob.pyr_down(img, ow, oh))


def test_pyr_down_output_extent_defines_work(gpu, ob):
    """The reference lets the caller pick any output extent (repeat-edge input)."""
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(3)
    img = noise_image(rng, 50, 70)
    for (oh, ow) in ((25, 35), (30, 40), (7, 90), (1, 1)):
        out = np.zeros((oh, ow), np.uint8)
        assert ip.PyrDown(img, out, gpu)
        assert np.array_equal(out, ob.pyr_down(img, ow, oh))


def test_pyr_down_strided_views(gpu, ob):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(4)
    big = noise_image(rng, 100, 203)
    img = big[3:93, 5:186]          # non-contiguous rows, unaligned start
    outbig = np.zeros((60, 120), np.uint8)
    out = outbig[2:47, 7:97]
    assert ip.PyrDown(img, out, gpu)
    assert np.array_equal(out, ob.pyr_down(np.ascontiguousarray(img), 90, 45))
    assert outbig[:2].sum() == 0 and outbig[:, :7].sum() == 0


@pytest.mark.parametrize("h,w", SIZES)
def test_grad_xy(gpu, ob, h, w):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h * 31 + w)
    img = noise_image(rng, h, w)
    gx = np.zeros((h, w), np.float32)
    gy = np.zeros((h, w), np.float32)
    assert ip.GradXY(img, gx, gy, gpu)
    ox, oy = ob.grad_xy(img)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)


@pytest.mark.parametrize("h,w", [(22, 40), (45, 80), (90, 160), (180, 320), (360, 640), (135, 240), (1080, 1920)])
def test_grad_argmax_and_jacobian(gpu, ob, h, w):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(h + w)
    img = noise_image(rng, h, w, smooth=1)
    gx, gy = ob.grad_xy(img)
    ok, tile, lmx, lmy = ip.GradArgMax(gx, gy, gpu)
    assert ok and tile == ob.tile_size(w, h)
    ox, oy = ob.grad_argmax(gx, gy, tile)
    assert np.array_equal(lmx, ox) and np.array_equal(lmy, oy)
    ok, jx, jy = ip.SparseJacobian(gx, gy, lmx, lmy, gpu)
    assert ok
    rx, ry = ob.sparse_jac(gx, gy, ox, oy)
    assert np.array_equal(jx, rx) and np.array_equal(jy, ry)


def test_grad_argmax_ties_and_flat_tiles(gpu, ob):
    """All-zero tiles return the tile origin; ties keep the first maximum in scan order."""
    from video_stabilizer_b200 import imgproc as ip
    h, w = 64, 96
    gx = np.zeros((h, w), np.float32)
    gy = np.zeros((h, w), np.float32)
    gx[10:20, 30:50] = 3.5       # a plateau spanning tiles
    gy[::7, ::5] = -2.0
    ok, tile, lmx, lmy = ip.GradArgMax(gx, gy, gpu)
    ox, oy = ob.grad_argmax(gx, gy, tile)
    assert ok and np.array_equal(lmx, ox) and np.array_equal(lmy, oy)


@pytest.mark.parametrize("T", [(0, 0, 0, 0), (0.001, -0.0007, 1.3, -2.2), (-0.01, 0.02, 5.5, 3.25), (0.0, 0.0, -40.0, 25.0)])
def test_sparse_warpdiff_and_ica(gpu, ob, T):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(11)
    h, w = 180, 320
    key = noise_image(rng, h, w, smooth=2)
    tmpl = np.roll(key, (1, -2), (0, 1))
    gx, gy = ob.grad_xy(key)
    tile = ob.tile_size(w, h)
    lmx, lmy = ob.grad_argmax(gx, gy, tile)
    st = ip.SimilarityTransform(*T)
    ok, wd = ip.SparseWarpDiff(tmpl, key, lmx, st, gpu)
    assert ok and np.array_equal(wd, ob.sparse_warpdiff(tmpl, key, lmx, T))
    jx, jy = ob.sparse_jac(gx, gy, lmx, lmy)
    k = lmx.shape[1] * lmx.shape[2]
    selx = lmx.reshape(2, k)[:, : k * 4 // 5]
    sely = lmy.reshape(2, k)[:, 5: k * 4 // 5]
    sjx = jx.reshape(4, k)[:, : k * 4 // 5]
    sjy = jy.reshape(4, k)[:, 5: k * 4 // 5]
    ok, b = ip.SparseICA(tmpl, key, selx, sely, sjx, sjy, st, gpu)
    ref = ob.sparse_ica(tmpl, key, selx, sely, sjx, sjy, T)
    assert ok
    # same f32 products; only the f64 summation order differs
    assert np.allclose(b, ref, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("T", [(0, 0, 5, 7), (0.01, -0.02, 3.3, -2.1), (-0.2, 0.1, -30.5, 12.25)])
def test_image_warp(gpu, ob, T):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(5)
    img = noise_image(rng, 97, 131)
    for (oh, ow) in ((97, 131), (40, 200)):
        out = np.zeros((oh, ow), np.float32)
        assert ip.ImageWarp(img, ip.SimilarityTransform(*T), out, gpu)
        assert np.array_equal(out, ob.image_warp(img, T, ow, oh))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("border", [0, 1])
def test_bgr_warp(gpu, ob, mode, border):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(17)
    for (h, w) in ((97, 131), (360, 640)):
        img = noise_image(rng, h, w, 3)
        for T in ((0, 0, 0, 0), (0.01, -0.02, 3.3, -2.1), (-0.03, 0.05, -7.7, 4.2), (0.2, 0.1, 30, -20)):
            for crop in (0, 8):
                got = ip.warpBySimilarityTransform(img, ip.SimilarityTransform(*T), gpu, mode, border, crop)
                ref = ob.warp_bgr(img, T, mode, border, crop)
                assert got.shape == ref.shape
                # integer mode must be bit-exact; float modes are the same f32 operations
                assert np.array_equal(got, ref), (mode, border, T, crop, int(np.abs(got.astype(int) - ref).max()))


def test_bgr_warp_1080p_exact(gpu, ob):
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(23)
    img = noise_image(rng, 1080, 1920, 3)
    T = (0.0013, -0.0021, 6.37, -3.81)
    got = ip.warpBySimilarityTransform(img, ip.SimilarityTransform(*T), gpu)
    assert np.array_equal(got, ob.warp_bgr(img, T))


def test_errors_are_reported_not_swallowed(gpu):
    import ctypes as C
    from video_stabilizer_b200 import _capi as capi
    lib = gpu.lib
    bad = capi.VsImg(None, 8, 8, 8, 1, 0)
    out = np.zeros((4, 4), np.uint8)
    r = lib.vs_pyr_down_u8(gpu.handle, C.byref(bad), C.byref(capi.img_of(out)), capi.VS_MEM_HOST)
    assert r == -1 and b"NULL" in lib.vs_last_error(gpu.handle)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (7680, 4320)])
@pytest.mark.parametrize("mode", [1, 2])
def test_bgr_warp_float_modes_full_size(gpu, ob, w, h, mode):
    """BASELINE.json configs[4] sizes: the float-bilinear and Lanczos-2 BGR warps on whole 1080p / 4K / 8K frames, compared
    with the oracle on bands of rows (top, bottom, and spread over the frame) — bit-exact."""
    from video_stabilizer_b200 import imgproc as ip
    rng = np.random.default_rng(w + mode)
    src = noise_image(rng, h, w, 3)
    T = (0.0013, -0.0021, 6.37, -3.81)
    got = ip.warpBySimilarityTransform(src, ip.SimilarityTransform(*T), gpu, mode=mode, border=0)
    M = ip.forward_matrix(ip.SimilarityTransform(*T), w, h)
    bands = [0, 5, h // 3, h // 2 + 1, h - 24, h - 8]
    for y0 in bands:
        want = ob.warp_bgr_matrix(src, M, w, 8, dx0=0, dy0=y0, mode=mode, border=0, fast=False)
        assert np.array_equal(got[y0:y0 + 8], want), (y0, int(np.abs(got[y0:y0 + 8].astype(int) - want.astype(int)).max()))
