import numpy as np


def corner_displacement(Ta, Tb, w, h):
    """Max distance between the images of the 4 frame corners under two centre-based
    similarity transforms (the 0.01 px parity metric of BASELINE.json)."""
    cx, cy = w * 0.5, h * 0.5
    d = 0.0
    for (x, y) in ((0.0, 0.0), (w, 0.0), (0.0, h), (w, h)):
        pts = []
        for T in (Ta, Tb):
            A, B, TX, TY = [float(v) for v in T]
            px, py = x - cx, y - cy
            pts.append(((1 + A) * px - B * py + cx + TX, B * px + (1 + A) * py + cy + TY))
        d = max(d, float(np.hypot(pts[0][0] - pts[1][0], pts[0][1] - pts[1][1])))
    return d


def noise_image(rng, h, w, channels=None, smooth=0):
    shape = (h, w) if channels is None else (h, w, channels)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    if smooth:
        from scipy.ndimage import gaussian_filter
        f = gaussian_filter(img.astype(np.float32), [smooth, smooth] + ([0] if channels else []), mode="nearest")
        f = (f - f.min()) / max(f.max() - f.min(), 1e-6) * 255
        img = f.astype(np.uint8)
    return img
