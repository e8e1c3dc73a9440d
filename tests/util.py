"""Shared helpers for the parity tests (seeded inputs, random cameras, scenes)."""
import math

import numpy as np


def rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (Rz @ Ry @ Rx).astype(np.float32)


def random_camera(rng, W, H, yaw=None):
    f = float(rng.uniform(0.8, 1.6) * W / 2)
    K = np.array([[f, rng.uniform(-2, 2), W / 2 + rng.uniform(-5, 5)],
                  [0, f * rng.uniform(0.9, 1.1), H / 2 + rng.uniform(-5, 5)], [0, 0, 1]], np.float32)
    R = rot(rng.uniform(-0.2, 0.2), rng.uniform(-3.1, 3.1) if yaw is None else yaw, rng.uniform(-0.2, 0.2))
    return K, R


def smooth_image(rng, h, w, cn=3, dtype=np.uint8):
    a = rng.integers(0, 256, (h, w, cn)).astype(np.float64)
    for _ in range(2):
        a = (np.roll(a, 1, 0) + 2 * a + np.roll(a, -1, 0)) / 4
        a = (np.roll(a, 1, 1) + 2 * a + np.roll(a, -1, 1)) / 4
    a = np.clip(np.round(a), 0, 255).astype(dtype)
    return a if cn > 1 else a[:, :, 0]


def special_maps(rng, h, w, W, H):
    """Random float maps with out-of-range, integer, tie, huge, NaN and Inf coordinates injected."""
    xm = rng.uniform(-40, W + 40, (h, w)).astype(np.float32)
    ym = rng.uniform(-40, H + 40, (h, w)).astype(np.float32)
    xm[0, :20] = np.arange(20) - 3
    ym[0, :20] = 5
    xm[1, :64] = (np.arange(64) - 10) / 64.0 + 7
    ym[1, :64] = np.arange(64) / 64.0 + 3 + 1 / 64
    bad = np.array([1e9, -1e9, 3e9, -3e9, np.nan, np.inf, -np.inf, 1e20], np.float32)
    xm[2, :8] = bad
    ym[2, :8] = 5
    ym[3, :8] = bad
    xm[3, :8] = 5
    xm[4, :6] = [-1, -1, W - 1, W - 0.5, W, -0.5]
    ym[4, :6] = [-1, H - 1, H - 0.5, 3, 3, -0.5]
    return xm, ym


def blend_scene(rng, n, H, W, spread, nonbinary=True, dtype=np.int16):
    """n overlapping images with masks and corners (negative corners included)."""
    imgs, masks, tls = [], [], []
    for _ in range(n):
        h, w = int(rng.integers(H // 2, H)), int(rng.integers(W // 2, W))
        img = smooth_image(rng, h, w).astype(dtype)
        m = np.full((h, w), 255, np.uint8)
        m[:int(rng.integers(0, max(h // 4, 1))), :] = 0
        m[:, :int(rng.integers(0, max(w // 4, 1)))] = 0
        if nonbinary and w > 16:
            m[:, w // 2:w // 2 + 6] = np.array([40, 80, 120, 160, 200, 240], np.uint8)
        imgs.append(img)
        masks.append(m)
        tls.append((int(rng.integers(-spread, spread)), int(rng.integers(-20, 20))))
    return imgs, masks, tls


def exposure_scene(n, w, h, seed, overlap=0.3):
    """n overlapping 8UC3 views of one smooth scene with different exposure, masks with holes, corners (some negative):
    the input shape of ExposureCompensator::feed (exposure_compensate.cpp:64-71)."""
    r = np.random.default_rng(seed)
    step = int(w * (1 - overlap))
    base = r.integers(0, 256, (h + 40, step * n + w, 3), dtype=np.uint8).astype(np.float32)
    for _ in range(3):                                       # cheap smoothing without cv2 (the GPU box may lack it)
        base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, -1, 0) + np.roll(base, -1, 1)) / 5
    imgs, masks, corners = [], [], []
    for i in range(n):
        x0, y0 = i * step, (7 * i) % 30
        g = 0.75 + 0.12 * i
        im = np.clip(base[y0:y0 + h, x0:x0 + w] * g, 0, 255).astype(np.uint8)
        m = np.full((h, w), 255, np.uint8)
        m[:4] = 0
        m[:, :3] = 0
        m[r.random((h, w)) > 0.97] = 0
        m[r.random((h, w)) > 0.995] = 128                    # not the level value: excluded like 0
        imgs.append(np.ascontiguousarray(im)); masks.append(m); corners.append((x0 - 60, y0 - 11))
    return corners, imgs, masks
