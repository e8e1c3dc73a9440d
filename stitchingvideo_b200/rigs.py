"""Deterministic synthetic camera rigs and frames (SURVEY.md §8d "Synthetic rigs").

Camera i of n: K = [[f,0,W/2],[0,f,H/2],[0,0,1]] (float32), R_i = R_y(180deg + i*360deg/n) (float32):
one camera is centred on the +-180deg seam, so exactly one warped image is panorama-wide.
Pure numpy host code (calibration-side inputs); no image processing happens here.
"""
import math

import numpy as np

RIGS = {
    # name: (n_cameras, W, H, focal, warper, scale, blender, gains?)
    "c1": dict(n=5, pick=(3, 4), W=1920, H=1080, f=1050.0, warper="spherical", scale=1050.0, blender="multiband", gains=False),
    "c2": dict(n=5, W=1920, H=1080, f=1050.0, warper="cylindrical", scale=1050.0, blender="feather", gains=False),
    "c3": dict(n=5, W=1920, H=1080, f=1050.0, warper="spherical", scale=1050.0, blender="multiband", gains=True),
    "c4": dict(n=8, W=3840, H=2160, f=2900.0, warper="spherical", scale=2900.0, blender="multiband", gains=True),
    "c5": dict(n=8, W=3840, H=2160, f=2900.0, warper="spherical", scale=16384.0 / (2.0 * math.pi), blender="multiband", gains=True),
    # the live app's published per-frame case: 6 x 1920x1088, cached-map
    # cylindrical remap + composite without blending or gain, panorama ~8040 x 1105 (2 pi f = 8040 -> f = 1280)
    # Round 2: as the app runs it (APP64:748-759) - BlockApply with per-camera block gain maps, the composite cropped by the
    # app's margins (upblack = downblack = 0.1, leftblack = rightblack = 10, APP64:47) with feedSizeRemap's unconditional gather
    "app6": dict(n=6, W=1920, H=1088, f=1280.0, warper="cylindrical", scale=1280.0, blender="no", gains=False,
                 block_gains=True, crop=(0.1, 0.1, 10, 10), crop_app_fill=True),
    # small rigs for fast parity tests (same construction, scaled down)
    "mini": dict(n=5, W=240, H=136, f=131.0, warper="spherical", scale=131.0, blender="multiband", gains=True),
    "mini_cyl": dict(n=5, W=240, H=136, f=131.0, warper="cylindrical", scale=131.0, blender="feather", gains=False),
    # shape coverage of the streaming frame kernel: warper scale far below / above the focal length (source boxes too
    # large for shared memory -> direct gathers; large boxes; tiny boxes) and heavy overlap (3+ cameras per tile)
    "mini_cyl_s3": dict(n=5, W=480, H=272, f=262.0, warper="cylindrical", scale=262.0 / 3.0, blender="feather", gains=False),
    "mini_cyl_s17": dict(n=5, W=480, H=272, f=262.0, warper="cylindrical", scale=262.0 / 1.7, blender="feather", gains=True),
    "mini_cyl_up": dict(n=5, W=240, H=136, f=131.0, warper="cylindrical", scale=131.0 * 1.6, blender="feather", gains=False),
    "mini_cyl_n9": dict(n=9, W=240, H=136, f=100.0, warper="cylindrical", scale=100.0, blender="feather", gains=True),
    "mini_sph_n7": dict(n=7, W=240, H=136, f=110.0, warper="spherical", scale=140.0, blender="feather", gains=False),
}
GAINS = [0.95, 1.02, 1.00, 0.98, 1.05, 0.97, 1.03, 0.99]


def rot_y(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64).astype(np.float32)


def cameras(name):
    """-> (Ks, Rs, spec) for a named rig."""
    spec = dict(RIGS[name])
    n, W, H, f = spec["n"], spec["W"], spec["H"], spec["f"]
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)
    idx = spec.get("pick") or tuple(range(n))
    Ks = [K.copy() for _ in idx]
    Rs = [rot_y(180.0 + i * 360.0 / n) for i in idx]
    spec["n_used"] = len(idx)
    spec["gain_values"] = [GAINS[i % len(GAINS)] for i in range(len(idx))] if spec["gains"] else None
    return Ks, Rs, spec


def block_gain_maps(name, warped_sizes):
    """Deterministic BlocksGainCompensator::gain_maps_ for a rig (32x32 blocks of each warped image, gains in 0.85..1.15)."""
    rng = np.random.default_rng(7)
    return [rng.uniform(0.85, 1.15, ((h + 31) // 32, (w + 31) // 32)).astype(np.float32) for (w, h) in warped_sizes]


def _blur121(a, passes):
    """cheap separable [1 2 1]/4 smoothing in uint16 (natural-ish spectrum; the path is data independent)."""
    a = a.astype(np.uint16)
    for _ in range(passes):
        p = np.pad(a, ((1, 1), (0, 0), (0, 0)), mode="edge")
        a = (p[:-2] + 2 * p[1:-1] + p[2:] + 2) >> 2
        p = np.pad(a, ((0, 0), (1, 1), (0, 0)), mode="edge")
        a = (p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:] + 2) >> 2
    return a.astype(np.uint8)


def frame(name, frame_idx, cam_idx, smooth=3):
    """Seeded synthetic frame: default_rng(frame_idx * n + cam).integers(0, 256) then smoothing."""
    spec = RIGS[name]
    rng = np.random.default_rng(frame_idx * spec["n"] + cam_idx)
    img = rng.integers(0, 256, (spec["H"], spec["W"], 3), dtype=np.uint8)
    return _blur121(img, smooth) if smooth else img


def seam_mask(warped_size, x_lo_frac, x_hi_frac, ramp=6):
    """Vertical-seam mask in warped coordinates: 255 inside [lo, hi) with a `ramp`-px ramp of
    non-binary values at each edge (exercises the float weights)."""
    w, h = warped_size
    lo, hi = int(w * x_lo_frac), int(w * x_hi_frac)
    row = np.zeros(w, np.float64)
    row[lo:hi] = 255
    for k in range(ramp):
        v = 255.0 * (k + 1) / (ramp + 1)
        if 0 <= lo + k < w:
            row[lo + k] = min(row[lo + k], v)
        if 0 <= hi - 1 - k < w:
            row[hi - 1 - k] = min(row[hi - 1 - k], v)
    return np.repeat(np.round(row).astype(np.uint8)[None, :], h, axis=0)
