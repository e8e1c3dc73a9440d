"""CPU: live cross-check of the oracle against OpenCV (cv2 4.13, which reproduces the OpenCV 2.4.11
arithmetic of this path bit for bit) on fresh seeded inputs — larger and more varied
than the committed fixtures.  Skipped where cv2 is not importable (the fixtures still pin the oracle)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P
from tests import util

cv2 = pytest.importorskip("cv2")
cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)


def same(got, ref, what):
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    d = (got.view(np.uint32) != ref.view(np.uint32)) if got.dtype == np.float32 else (got != ref)
    assert not d.any(), "%s: %d of %d values differ" % (what, int(d.sum()), d.size)


@pytest.mark.parametrize("kind", ["spherical", "cylindrical"])
def test_build_maps_full_size(kind):
    """A 1080p camera, the seam-straddling one included (panorama-wide ROI)."""
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras("c3")
    for i in (0, 2):
        w, cw = O.Warper(kind, spec["scale"]), cv2.PyRotationWarper(kind, spec["scale"])
        roi, xm, ym = w.build_maps((spec["W"], spec["H"]), Ks[i], Rs[i])
        croi, cxm, cym = cw.buildMaps((spec["W"], spec["H"]), Ks[i], Rs[i])
        assert tuple(roi) == tuple(croi)
        same(xm, cxm, "xmap cam %d" % i)
        same(ym, cym, "ymap cam %d" % i)


def test_remap_random():
    rng = np.random.default_rng(200)
    for cn in (1, 3):
        H, W = 173, 241
        src = rng.integers(0, 256, (H, W, cn), dtype=np.uint8) if cn == 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
        xm, ym = util.special_maps(rng, 150, 210, W, H)
        for border in (cv2.BORDER_REFLECT, cv2.BORDER_CONSTANT, cv2.BORDER_REPLICATE, cv2.BORDER_REFLECT_101, cv2.BORDER_WRAP):
            for interp in (cv2.INTER_LINEAR, cv2.INTER_NEAREST):
                ref = cv2.remap(src, xm, ym, interp, borderMode=border, borderValue=(7, 9, 11))
                same(O.remap(src, xm, ym, interp, border, (7, 9, 11, 0)), ref, "remap cn=%d b=%d i=%d" % (cn, border, interp))


def test_integer_pyramids_random():
    rng = np.random.default_rng(201)
    for shape in ((64, 96), (51, 77), (2, 2), (1, 9), (33, 1)):
        for a in (rng.integers(-32768, 32768, shape + (3,)).astype(np.int16), rng.integers(0, 256, shape + (3,), dtype=np.uint8)):
            same(O.pyr_down(a), cv2.pyrDown(a), "pyrDown %s %s" % (shape, a.dtype))
            if min(shape) > 1:
                same(O.pyr_up(a), cv2.pyrUp(a), "pyrUp %s %s" % (shape, a.dtype))


def _cv_blend(b, imgs, masks, tls):
    sizes = [(im.shape[1], im.shape[0]) for im in imgs]
    b.prepare(O.result_roi(tls, sizes))
    for im, m, t in zip(imgs, masks, tls):
        b.feed(im, m, t)
    return b.blend(None, None)


@pytest.mark.parametrize("scene", [(2, 100, 150, 80), (5, 200, 300, 500), (4, 373, 600, 900)])
def test_blenders_random(scene):
    rng = np.random.default_rng(210 + scene[0])
    imgs, masks, tls = util.blend_scene(rng, *scene)
    sizes = [(im.shape[1], im.shape[0]) for im in imgs]
    cases = [(O.BLEND_NO, {}, cv2.detail.Blender_createDefault(0), 0),
             (O.BLEND_FEATHER, {"sharpness": 0.02}, cv2.detail_FeatherBlender(0.02), 0),
             (O.BLEND_FEATHER, {"sharpness": 0.1}, cv2.detail_FeatherBlender(0.1), 0),
             (O.BLEND_MULTI_BAND, {"num_bands": 5, "weight_type": O.CV_16S}, cv2.detail_MultiBandBlender(0, 5, cv2.CV_16S), 0),
             (O.BLEND_MULTI_BAND, {"num_bands": 3, "weight_type": O.CV_16S}, cv2.detail_MultiBandBlender(0, 3, cv2.CV_16S), 0),
             (O.BLEND_MULTI_BAND, {"num_bands": 5, "weight_type": O.CV_32F}, cv2.detail_MultiBandBlender(0, 5, cv2.CV_32F), 1)]
    for kind, kw, cvb, tol in cases:
        b = O.Blender(kind, **kw)
        b.prepare(tls, sizes)
        for im, m, t in zip(imgs, masks, tls):
            b.feed(im, m, t)
        dst, dmask = b.blend()
        ref, rmask = _cv_blend(cvb, imgs, masks, tls)
        assert dst.shape == ref.shape
        d = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= tol, "blender %s %s: max |diff| %d" % (kind, kw, int(d.max()))
        same(dmask, rmask, "mask %s %s" % (kind, kw))


def test_frame_loop_against_cv2_pipeline():
    """The whole per-frame loop (stitcher.cpp:221-313) on the small spherical rig: cv2's warper,
    convertScaleAbs gain, MultiBandBlender(CV_16S weights: all-integer, exactly reproducible)."""
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras("mini")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, "spherical", spec["scale"])
    frames = [rigs.frame("mini", 0, i) for i in range(n)]
    got, gmask = P.compose(cal, frames, blender="multiband", num_bands=5, weight_type=O.CV_16S, gains=spec["gain_values"])
    cw = cv2.PyRotationWarper("spherical", spec["scale"])
    ones = np.full((size[1], size[0]), 255, np.uint8)
    corners, warped, masks = [], [], []
    for i in range(n):
        tl, img = cw.warp(frames[i], Ks[i], Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = cw.warp(ones, Ks[i], Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(tuple(tl))
        warped.append(cv2.convertScaleAbs(img, alpha=spec["gain_values"][i]).astype(np.int16))
        masks.append(m)
    assert corners == cal.corners
    ref, rmask = _cv_blend(cv2.detail_MultiBandBlender(0, 5, cv2.CV_16S), warped, masks, corners)
    same(got, np.clip(ref, 0, 255).astype(np.uint8), "panorama")
    same(gmask, rmask, "panorama mask")


@pytest.mark.parametrize("n,w,h", [(2, 120, 90), (3, 160, 100), (4, 90, 70)])
def test_feather_create_weight_maps_against_cv2(n, w, h):
    """FeatherBlender::createWeightMaps (blenders.cpp:158-186), incl. an all-zero mask"""
    corners, _, masks = util.exposure_scene(n, w, h, seed=n)
    masks = [np.where(m == 255, 255, 0).astype(np.uint8) for m in masks]
    if n == 4:
        masks[2][:] = 0
    roi, maps = cv2.detail_FeatherBlender(0.05).createWeightMaps([cv2.UMat(m) for m in masks], corners, None)
    oroi, omaps = O.feather_create_weight_maps(masks, corners, 0.05)
    assert tuple(roi) == tuple(oroi)
    for a, b in zip(maps, omaps):
        same(b, a.get(), "normalised weight map")


def test_mask_refinement_primitives_random_shapes():
    """cv::dilate (3x3 default) and cv::resize INTER_LINEAR on 8UC1 over random shapes and scale factors: up, down, exact
    2x decimation in both / one dimension (INTER_AREA rerouting only when both are 2), degenerate 1-pixel sizes."""
    rng = np.random.default_rng(2026)
    for k in range(60):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 120))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        a[rng.random((h, w)) < 0.5] = 0
        same(O.dilate3x3(a), cv2.dilate(a, None), "dilate %dx%d" % (w, h))
        mode = k % 4
        if mode == 0 and h % 2 == 0 and w % 2 == 0:
            ds = (w // 2, h // 2)
        elif mode == 1 and w % 2 == 0:
            ds = (w // 2, int(rng.integers(1, 200)))
        elif mode == 2:
            ds = (int(w * rng.uniform(1.0, 4.0)) + 1, int(h * rng.uniform(1.0, 4.0)) + 1)
        else:
            ds = (int(rng.integers(1, 200)), int(rng.integers(1, 150)))
        same(O.resize_linear_8u(a, ds), cv2.resize(a, ds, interpolation=cv2.INTER_LINEAR), "resize %dx%d -> %dx%d" % (w, h, ds[0], ds[1]))
        mw = (rng.random((ds[1], ds[0])) > 0.3).astype(np.uint8) * 255
        same(O.refine_seam_mask(a, mw), cv2.resize(cv2.dilate(a, None), ds) & mw, "refine")
