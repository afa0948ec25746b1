"""Generates tests/golden/plane_warp.npz: cv2.warpAffine (cv2 4.13 in this container) on 1-channel and 2-channel u8
images, the planes of an NV12 frame, for the product's VS_CLIP_NV12 clips (no counterpart upstream).  The forward matrix
is built like warpBySimilarityTransform (imgproc.cpp:458-481); the UV plane of a w x h frame is a (w/2) x (h/2)
two-channel image warped by the same similarity with half the translation.  tests/test_oracle_golden.py pins the oracle's
vo_warp_plane_matrix / vo_warp_nv12 against this file.

Run from the repo root:  python tests/golden/make_plane_warp_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261019)
w, h = 86, 62
nv12 = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
transforms = np.array([
    [0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 6.0, -4.0],
    [0.0, 0.0, 0.5, -0.25],
    [0.013, -0.021, 3.37, -2.81],
    [-0.05, 0.08, -9.6, 4.4],
    [0.25, 0.1, 20.0, -15.0],
], np.float64)


def matrix(T, pw, ph):
    A, B, TX, TY = T
    cx, cy = (pw - 1) * 0.5, (ph - 1) * 0.5
    return np.array([[1.0 + A, -B, TX - A * cx + B * cy], [B, 1.0 + A, TY - B * cx - A * cy]], np.float64)


def warp(img, M):
    hh, ww = img.shape[:2]
    return cv2.warpAffine(img, M, (ww, hh), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)


Y = nv12[:h]
UV = nv12[h:].reshape(h // 2, w // 2, 2)
outs = []
for T in transforms:
    y = warp(Y, matrix(T, w, h))
    uv = warp(UV, matrix([T[0], T[1], T[2] * 0.5, T[3] * 0.5], w // 2, h // 2))
    outs.append(np.concatenate([y, uv.reshape(h // 2, w)], 0))
np.savez_compressed(os.path.join(HERE, "plane_warp.npz"), cv2_version=np.array(cv2.__version__), nv12=nv12,
                    T=transforms, out=np.stack(outs))
print("wrote plane_warp.npz", np.stack(outs).shape)
