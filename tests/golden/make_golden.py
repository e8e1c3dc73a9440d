"""Generate the golden input/output vectors under tests/golden/ from OpenCV (cv2 4.13, IPP off).

The reference ships no golden vectors for this path and its OpenCV 2.4.11 binaries
cannot run here, so the vectors come from the cv2 build in this container, which reproduces the
2.4.11 arithmetic of the path bit for bit.  Run from the repo root:
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box (no /root/reference, possibly no cv2) only reads them.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import util  # noqa: E402

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})


def gen_maps():
    rng = np.random.default_rng(100)
    out = {}
    for kind in ("spherical", "cylindrical", "plane"):
        W, H = 160, 100
        K, R = util.random_camera(rng, W, H, yaw=0.4 if kind == "plane" else None)   # plane: keep the ROI finite
        scale = float(rng.uniform(100, 200))
        w = cv2.PyRotationWarper(kind, scale)
        roi, xm, ym = w.buildMaps((W, H), K, R)
        out.update({kind + "_K": K, kind + "_R": R, kind + "_scale": np.float32(scale), kind + "_size": np.array([W, H]),
                    kind + "_roi": np.array(roi), kind + "_xmap": xm, kind + "_ymap": ym,
                    kind + "_warproi": np.array(w.warpRoi((W, H), K, R)),
                    kind + "_pt": np.array(w.warpPoint((W / 3.0, H / 5.0), K, R), np.float32)})
    save("maps", **out)


def gen_remap():
    rng = np.random.default_rng(101)
    H, W = 45, 61
    out = {}
    for cn in (1, 3):
        src = rng.integers(0, 256, (H, W, cn), dtype=np.uint8) if cn == 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
        xm, ym = util.special_maps(rng, 40, 70, W, H)
        out["src%d" % cn], out["xmap%d" % cn], out["ymap%d" % cn] = src, xm, ym
        for border in (0, 1, 2, 3, 4):
            for interp in (0, 1):
                out["dst%d_b%d_i%d" % (cn, border, interp)] = cv2.remap(src, xm, ym, interp, borderMode=border, borderValue=(7, 9, 11))
    save("remap", **out)


def gen_remap_fixed():
    """Fixed-point map pair (cv::convertMaps) and cv::remap through it: the app's video front-end format."""
    rng = np.random.default_rng(105)
    H, W = 45, 61
    out = {}
    for cn in (1, 3):
        src = rng.integers(0, 256, (H, W, cn), dtype=np.uint8) if cn == 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
        xm, ym = util.special_maps(rng, 40, 70, W, H)
        m1, m2 = cv2.convertMaps(xm, ym, cv2.CV_16SC2)
        n1, _ = cv2.convertMaps(xm, ym, cv2.CV_16SC2, nninterpolation=True)
        out.update({"src%d" % cn: src, "xmap%d" % cn: xm, "ymap%d" % cn: ym, "map1_%d" % cn: m1, "map2_%d" % cn: m2, "nnmap1_%d" % cn: n1})
        for border in (0, 1, 2, 3, 4):
            for interp in (0, 1):
                out["dst%d_b%d_i%d" % (cn, border, interp)] = cv2.remap(src, m1, m2, interp, borderMode=border, borderValue=(7, 9, 11))
            out["nndst%d_b%d" % (cn, border)] = cv2.remap(src, n1, None, cv2.INTER_NEAREST, borderMode=border, borderValue=(7, 9, 11))
    save("remap_fixed", **out)


def gen_pyr():
    rng = np.random.default_rng(102)
    out = {}
    for name, a in (("s16", rng.integers(-32768, 32768, (38, 50, 3)).astype(np.int16)),
                    ("u8", rng.integers(0, 256, (37, 51, 3), dtype=np.uint8)),
                    ("s16c1", rng.integers(0, 257, (40, 48)).astype(np.int16))):
        out[name] = a
        out[name + "_down"] = cv2.pyrDown(a)
        out[name + "_up"] = cv2.pyrUp(a)
    save("pyr", **out)


def gen_misc():
    rng = np.random.default_rng(103)
    img = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    out = {"img": img}
    for i, g in enumerate((0.95, 1.02, 1.5, 2.5)):
        out["gain%d" % i] = np.float64(g)
        out["gain%d_out" % i] = cv2.convertScaleAbs(img, alpha=g)        # saturate_cast<uchar>(p * (float)g)
    gm = rng.uniform(0.8, 1.2, (5, 7)).astype(np.float32)
    out["gmap"], out["gmap_resized"] = gm, cv2.resize(gm, (40, 30), interpolation=cv2.INTER_LINEAR)
    c = cv2.detail_BlocksGainCompensator()
    m = np.zeros((40, 60), np.uint8)
    m[5:35, 8:50] = 255
    m[15:18, 20:30] = 0
    out["mask"], out["dist"] = m, cv2.distanceTransform(m, cv2.DIST_L1, 3)
    save("misc", **out)


def gen_blend():
    rng = np.random.default_rng(104)
    imgs, masks, tls = util.blend_scene(rng, 3, 70, 110, 80)
    sizes = [(im.shape[1], im.shape[0]) for im in imgs]
    tl = np.array(tls)
    br = np.array([(t[0] + s[0], t[1] + s[1]) for t, s in zip(tls, sizes)])
    roi = (int(tl[:, 0].min()), int(tl[:, 1].min()), int(br[:, 0].max() - tl[:, 0].min()), int(br[:, 1].max() - tl[:, 1].min()))
    out = {"tls": np.array(tls), "n": np.int32(len(imgs))}
    for i, (im, m) in enumerate(zip(imgs, masks)):
        out["img%d" % i], out["mask%d" % i] = im, m
    cases = {"no": cv2.detail.Blender_createDefault(0), "feather": cv2.detail_FeatherBlender(0.02),
             "mb16_5": cv2.detail_MultiBandBlender(0, 5, cv2.CV_16S), "mb16_2": cv2.detail_MultiBandBlender(0, 2, cv2.CV_16S),
             "mb32_5": cv2.detail_MultiBandBlender(0, 5, cv2.CV_32F)}
    for name, b in cases.items():
        b.prepare(roi)
        for im, m, t in zip(imgs, masks, tls):
            b.feed(im, m, t)
        d, dm = b.blend(None, None)
        out[name + "_dst"], out[name + "_mask"] = d, dm
    save("blend", **out)


def gen_calib():
    """ExposureCompensator::feed and the seam-mask refinement from cv2."""
    rng = np.random.default_rng(106)
    out = {}
    for k, (shp, ds) in enumerate((((37, 53), (106, 74)), ((40, 64), (32, 20)), ((64, 2), (5, 100)), ((50, 70), (70, 50)),
                                   ((31, 45), (100, 77)), ((1, 9), (20, 3)))):
        a = (rng.random(shp) > 0.7).astype(np.uint8) * 255
        a[rng.random(shp) > 0.9] = rng.integers(1, 255)
        out["m%d" % k] = a
        out["m%d_dilate" % k] = cv2.dilate(a, None)
        out["m%d_resize" % k] = cv2.resize(a, ds, interpolation=cv2.INTER_LINEAR)
        mw = (rng.random((ds[1], ds[0])) > 0.2).astype(np.uint8) * 255
        out["m%d_warped" % k] = mw
        out["m%d_refined" % k] = cv2.resize(cv2.dilate(a, None), ds) & mw          # stitcher.cpp:291-294
    for n, w, h in ((2, 120, 90), (3, 160, 100), (5, 200, 120)):
        corners, imgs, masks = util.exposure_scene(n, w, h, seed=n)
        masks = [np.where(m == 255, 255, 0).astype(np.uint8) for m in masks]        # cv2 4.x tests mask != 0, 2.4.11 mask == 255
        c = cv2.detail_GainCompensator(1)
        c.feed(corners, [cv2.UMat(i) for i in imgs], [cv2.UMat(m) for m in masks])
        out["gains%d" % n] = np.array(c.getMatGains(), np.float64).reshape(-1)
        bc = cv2.detail_BlocksGainCompensator(32, 32, 1)
        bc.setNrGainsFilteringIterations(2)
        bc.feed(corners, [cv2.UMat(i) for i in imgs], [cv2.UMat(m) for m in masks])
        for i, m in enumerate(bc.getMatGains()):
            out["blocks%d_%d" % (n, i)] = np.array(m, np.float32)
    g = rng.uniform(0.8, 1.2, (9, 13)).astype(np.float32)
    ker = np.array([[0.25, 0.5, 0.25]], np.float32)
    out["gmap"], out["gmap_smooth"] = g, cv2.sepFilter2D(g, cv2.CV_32F, ker, ker)
    save("calib", **out)


if __name__ == "__main__":
    gen_maps()
    gen_remap()
    gen_remap_fixed()
    gen_pyr()
    gen_misc()
    gen_blend()
    gen_calib()
