"""Generates tests/golden/cv2_golden.npz: outputs of OpenCV (cv2, this container: 4.13) for the
third-party calls on the reference's hot path, on small seeded inputs.  OpenCV's C++ library
is what the reference links (alignment.cpp:212,558,582; imgproc.cpp:473); the oracle restates
those calls and tests/test_oracle_golden.py pins it against this file.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261018)
out = {"cv2_version": np.array(cv2.__version__)}

# --- cvtColor(BGR2GRAY): all 256^3 combinations are too many; random + corner cases
bgr = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
bgr[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 254], [128, 127, 129]]
out["gray_in"] = bgr
out["gray_out"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)

# --- warpAffine(INTER_LINEAR, BORDER_CONSTANT 0), forward matrix (no WARP_INVERSE_MAP),
#     built exactly like warpBySimilarityTransform (imgproc.cpp:458-481)
img = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
transforms = np.array([
    [0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 5.0, 7.0],
    [0.0, 0.0, 0.5, -0.5],
    [0.013, -0.021, 3.37, -2.81],
    [-0.05, 0.08, -9.6, 4.4],
    [0.25, 0.1, 20.0, -15.0],
    [-0.3, -0.2, 1.015625, 2.984375],
], np.float64)
warps = []
h, w, _ = img.shape
for A, B, TX, TY in transforms:
    cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
    M = np.array([[1.0 + A, -B, TX - A * cx + B * cy], [B, 1.0 + A, TY - B * cx - A * cy]], np.float64)
    warps.append(cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
out["warp_in"] = img
out["warp_T"] = transforms
out["warp_out"] = np.stack(warps)

# --- cv::SVD / Mat::inv(DECOMP_SVD) on 4x4 SPD matrices shaped like the aligner's Hessian
mats, ws, invs = [], [], []
for i in range(12):
    k = 200 + 50 * i
    J = np.zeros((2 * k, 4))
    g = rng.normal(0, 20, 2 * k)
    u = rng.uniform(-0.5, 0.5, 2 * k)
    v = rng.uniform(-0.3, 0.3, 2 * k)
    J[:k] = np.stack([g[:k] * u[:k], -g[:k] * v[:k], g[:k], np.zeros(k)], 1)
    J[k:] = np.stack([g[k:] * v[k:], g[k:] * u[k:], np.zeros(k), g[k:]], 1)
    if i >= 9:   # badly conditioned cases
        J[:, 0] *= 10.0 ** -(i - 6)
    H = J.T @ J
    w_, u_, vt_ = cv2.SVDecomp(H)
    ok, Hi = cv2.invert(H, flags=cv2.DECOMP_SVD)
    mats.append(H); ws.append(w_.ravel()); invs.append(Hi)
out["svd_H"] = np.stack(mats)
out["svd_w"] = np.stack(ws)
out["svd_inv"] = np.stack(invs)

# --- phaseCorrelate on the reference's ImageWarp test pattern (align_test.cpp:358-400):
#     64x64 black image, 10x10 white square at (20,20), shifted by (5,7)
a = np.zeros((64, 64), np.float32); a[20:30, 20:30] = 255
b = np.zeros((64, 64), np.float32); b[27:37, 25:35] = 255
(sx, sy), resp = cv2.phaseCorrelate(a, b)
out["phase_shift"] = np.array([sx, sy, resp])

np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)

# --- phaseCorrelate on u8-valued textures the way the aligner calls it (alignment.cpp:225-229, 374: level-2 image as
#     CV_32F, no window): even, odd and padded (non-5-smooth) sizes, translations, a small rotation + translation,
#     and two unrelated images (low response).  Written to a second file so cv2_golden.npz stays byte-stable.
prng = np.random.default_rng(20261019)
pc = {"cv2_version": np.array(cv2.__version__)}
cases = []
for i, (h, w) in enumerate([(64, 96), (45, 75), (67, 120), (90, 160), (61, 83), (135, 240)]):
    big = cv2.GaussianBlur(prng.random((h + 48, w + 48)).astype(np.float32), (0, 0), 1.5 + 0.5 * (i % 3))
    big = ((big - big.min()) / (big.max() - big.min()) * 255).astype(np.uint8)
    shifts = [(0, 0), (3, -5), (-7, 2), (11, 9)]
    for (dy, dx) in shifts:
        cases.append((big[24:24 + h, 24:24 + w], big[24 + dy:24 + dy + h, 24 + dx:24 + dx + w]))
    R = cv2.getRotationMatrix2D((big.shape[1] / 2, big.shape[0] / 2), 0.4, 1.002); R[:, 2] += (2.3, -1.6)
    rot = cv2.warpAffine(big, R, (big.shape[1], big.shape[0]), flags=cv2.INTER_LINEAR)
    cases.append((big[24:24 + h, 24:24 + w], rot[24:24 + h, 24:24 + w]))
    other = prng.integers(0, 256, (h, w), dtype=np.uint8)
    cases.append((big[24:24 + h, 24:24 + w], other))
res = []
for k, (a, b) in enumerate(cases):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    (sx, sy), resp = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    pc["a%d" % k] = a; pc["b%d" % k] = b
    res.append([sx, sy, resp])
pc["result"] = np.array(res)
np.savez_compressed(os.path.join(HERE, "phase_correlate.npz"), **pc)
print("wrote", os.path.join(HERE, "cv2_golden.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})
