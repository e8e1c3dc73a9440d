"""CPU: the oracle's restatement of the reference's LOGIC (oracle/so_stitch.c) against the reference's own
sources — blenders.cpp, warpers.cpp (+ warpers_inl.hpp), util.cpp — compiled where they lie into
oracle/_ref/libstitch_ref.so on the OpenCV stand-in of oracle/ref_shim (primitives shared with the oracle,
pinned against cv2 by tests/golden).  Bit-exact everywhere, float weights included (same primitive)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as RF
from tests import util

pytestmark = pytest.mark.skipif(not RF.available(), reason="oracle/_ref not built (needs /root/reference; run __graft_entry__.build())")


def same(got, ref, what):
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    d = (got.view(np.uint32) != ref.view(np.uint32)) if got.dtype == np.float32 else (got != ref)
    assert not d.any(), "%s: %d of %d values differ" % (what, int(d.sum()), d.size)


@pytest.mark.parametrize("scene", [(2, 100, 150, 80), (5, 200, 300, 500), (3, 60, 70, 30), (4, 373, 600, 900)])
@pytest.mark.parametrize("cfg", [(O.BLEND_NO, {}), (O.BLEND_FEATHER, {"sharpness": 0.02}), (O.BLEND_FEATHER, {"sharpness": 0.1}),
                                 (O.BLEND_MULTI_BAND, {"num_bands": 5}), (O.BLEND_MULTI_BAND, {"num_bands": 5, "weight_type": O.CV_16S}),
                                 (O.BLEND_MULTI_BAND, {"num_bands": 1}), (O.BLEND_MULTI_BAND, {"num_bands": 3, "weight_type": O.CV_16S}),
                                 (O.BLEND_MULTI_BAND, {"num_bands": 7}), (O.BLEND_MULTI_BAND, {"num_bands": 0})])
def test_blenders_oracle_equals_reference_sources(scene, cfg):
    rng = np.random.default_rng(300 + scene[0])
    imgs, masks, tls = util.blend_scene(rng, *scene)
    sizes = [(im.shape[1], im.shape[0]) for im in imgs]
    ob, rb = O.Blender(cfg[0], **cfg[1]), RF.Blender(cfg[0], **cfg[1])
    ob.prepare(tls, sizes)
    rb.prepare(tls, sizes)
    assert rb.roi == O.result_roi(tls, sizes)                   # util.cpp:127-140
    for im, m, tl in zip(imgs, masks, tls):
        ob.feed(im, m, tl)
        rb.feed(im, m, tl)
    (od, om), (rd, rm) = ob.blend(), rb.blend()
    same(od, rd, "blend image %s" % (cfg,))
    same(om, rm, "blend mask %s" % (cfg,))


def test_multiband_8u_and_saturating_inputs():
    rng = np.random.default_rng(301)
    for dtype, full in ((np.uint8, False), (np.int16, True)):
        imgs, masks, tls = util.blend_scene(rng, 3, 120, 200, 150, dtype=dtype)
        if full:
            imgs = [rng.integers(-32768, 32768, im.shape).astype(np.int16) for im in imgs]
        sizes = [(im.shape[1], im.shape[0]) for im in imgs]
        for wt in (O.CV_32F, O.CV_16S):
            ob, rb = O.Blender(O.BLEND_MULTI_BAND, 4, wt), RF.Blender(O.BLEND_MULTI_BAND, 4, wt)
            ob.prepare(tls, sizes)
            rb.prepare(tls, sizes)
            for im, m, tl in zip(imgs, masks, tls):
                ob.feed(im, m, tl)
                rb.feed(im, m, tl)
            (od, om), (rd, rm) = ob.blend(), rb.blend()
            same(od, rd, "%s multiband image" % dtype.__name__)
            same(om, rm, "%s multiband mask" % dtype.__name__)


@pytest.mark.parametrize("dtype", [np.int16, np.uint8])
def test_laplace_pyramid_helpers(dtype):
    rng = np.random.default_rng(302)
    for shape, levels in (((64, 96), 5), ((160, 224), 3), ((32, 32), 5), ((8, 24), 2), ((40, 56), 0)):
        img = (rng.integers(-32768, 32768, shape + (3,)).astype(np.int16) if dtype == np.int16
               else rng.integers(0, 256, shape + (3,), dtype=np.uint8))
        op, rp = O.create_laplace_pyr(img, levels), RF.create_laplace_pyr(img, levels)
        for l, (a, b) in enumerate(zip(op, rp)):
            same(a, b, "laplace level %d" % l)
        same(O.restore_from_laplace_pyr(op), RF.restore_from_laplace_pyr(op), "restore")


def test_weight_map_and_normalize():
    import ctypes as C
    rng = np.random.default_rng(303)
    m = np.zeros((80, 120), np.uint8)
    m[10:70, 15:100] = 255
    m[30:35, 40:50] = 0
    for sharp in (0.02, 0.1, 1.0):
        same(O.create_weight_map(m, sharp), RF.create_weight_map(m, sharp), "createWeightMap")
    src = rng.integers(-3000, 3000, (33, 47, 3)).astype(np.int16)
    wf = rng.uniform(0, 3, (33, 47)).astype(np.float32)
    wf[0, :5] = [0, 1e-7, 1e-6, 1e-5, 1.0]
    ws = rng.integers(0, 700, (33, 47)).astype(np.int16)
    for w in (wf, ws):
        got = src.copy()
        mw, ms = O.mat(w), O.mat(got)
        O.lib().so_normalize_using_weight_map(C.byref(mw), C.byref(ms))
        same(got, RF.normalize_using_weight_map(w, src), "normalizeUsingWeightMap %s" % w.dtype)


@pytest.mark.parametrize("kind", ["spherical", "cylindrical", "plane"])
def test_warpers_oracle_equals_reference_sources(kind):
    rng = np.random.default_rng(304)
    for _ in range(4):
        W, H = int(rng.integers(120, 400)), int(rng.integers(90, 300))
        K, R = util.random_camera(rng, W, H, yaw=0.4 if kind == "plane" else None)
        scale = float(rng.uniform(200, 600))
        ow, rw = O.Warper(kind, scale), RF.Warper(kind, scale)
        assert ow.warp_roi((W, H), K, R) == rw.warp_roi((W, H), K, R)
        ou, rv = ow.warp_point((W / 3.0, H / 5.0), K, R), rw.warp_point((W / 3.0, H / 5.0), K, R)
        assert np.float32(ou[0]) == np.float32(rv[0]) and np.float32(ou[1]) == np.float32(rv[1])
        oroi, oxm, oym = ow.build_maps((W, H), K, R)
        rroi, rxm, rym = rw.build_maps((W, H), K, R)
        assert tuple(oroi) == tuple(rroi)
        same(oxm, rxm, kind + " xmap")
        same(oym, rym, kind + " ymap")
        img = util.smooth_image(rng, H, W)
        (otl, od), (rtl, rd) = ow.warp(img, K, R), rw.warp(img, K, R)
        assert tuple(otl) == tuple(rtl)
        same(od, rd, kind + " warp")


def test_seam_straddling_camera_through_reference_sources():
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras("mini")
    ow, rw = O.Warper("spherical", spec["scale"]), RF.Warper("spherical", spec["scale"])
    for i in range(spec["n_used"]):
        assert ow.warp_roi((spec["W"], spec["H"]), Ks[i], Rs[i]) == rw.warp_roi((spec["W"], spec["H"]), Ks[i], Rs[i])


@pytest.mark.parametrize("name,cv_name,ab", [
    ("fisheye", "fisheye", None), ("stereographic", "stereographic", None), ("compressedRectilinear", "compressedPlaneA2B1", (2.0, 1.0)),
    ("compressedRectilinearPortrait", "compressedPlanePortraitA2B1", (2.0, 1.0)), ("panini", "paniniA2B1", (2.0, 1.0)),
    ("paniniPortrait", "paniniPortraitA2B1", (2.0, 1.0)), ("mercator", "mercator", None)])
def test_remaining_projectors_reference_sources_equal_cv2(name, cv_name, ab):
    """The reference's own projector code (compiled into oracle/_ref) gives the maps OpenCV 4.13 gives, bit for bit, for
    every projector cv2 exposes except transverseMercator (where 4.13's formula differs from the 2.4.11 source in the
    last ulp for ~1 % of the pixels; the 2.4.11 source — what the GPU tests compare against — is authoritative)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(305)
    for _ in range(3):
        W, H = int(rng.integers(120, 260)), int(rng.integers(90, 200))
        K, R = util.random_camera(rng, W, H, yaw=float(rng.uniform(-0.5, 0.5)))
        scale = float(rng.uniform(150, 400))
        rw, cw = RF.Warper(name, scale, *(ab or (1.0, 1.0))), cv2.PyRotationWarper(cv_name, scale)
        roi, xm, ym = rw.build_maps((W, H), K, R)
        croi, cxm, cym = cw.buildMaps((W, H), K, R)
        assert tuple(roi) == tuple(croi)
        same(xm, cxm, name + " xmap")
        same(ym, cym, name + " ymap")


# ---- exposure_compensate.cpp, compiled where it lies -------------------------------------------
@pytest.mark.parametrize("n,w,h", [(2, 120, 90), (3, 160, 100), (5, 200, 120), (4, 97, 61)])
def test_gain_compensator_feed_is_the_reference_code(n, w, h):
    """The oracle's restatement of GainCompensator::feed (so_calib.c) against the reference's own exposure_compensate.cpp
    (:76-147): same overlap loops, same sequential double sums, same normal equations -> the gains are EQUAL doubles."""
    corners, imgs, masks = util.exposure_scene(n, w, h, seed=40 + n)
    assert np.array_equal(RF.gain_feed(corners, imgs, masks), O.gain_feed(corners, imgs, masks))
    img = imgs[0]
    for g in (0.95, 1.02, 1.5, 2.5):
        assert np.array_equal(RF.gain_apply(img, g), O.gain_apply(img, g))


@pytest.mark.parametrize("n,w,h,bl", [(2, 120, 90, 32), (3, 160, 100, 32), (3, 150, 97, 20), (2, 64, 40, 64)])
def test_blocks_gain_compensator_feed_is_the_reference_code(n, w, h, bl):
    """BlocksGainCompensator::feed (:165-222) and ::apply (:225-246): block lists, one gain solve over all blocks, two
    smoothing passes, resize + multiply — gain maps and the compensated image bit for bit."""
    corners, imgs, masks = util.exposure_scene(n, w, h, seed=50 + n)
    ref_maps, ref_img0 = RF.blocks_gain_feed(corners, imgs, masks, bl, bl, apply_first=True)
    maps = O.blocks_gain_feed(corners, imgs, masks, bl, bl)
    for a, b in zip(maps, ref_maps):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(O.blocks_gain_apply(imgs[0], maps[0]), ref_img0)


@pytest.mark.parametrize("n,w,h,sharp", [(2, 120, 90, 0.02), (3, 160, 100, 0.05), (4, 90, 70, 0.3)])
def test_feather_create_weight_maps_is_the_reference_code(n, w, h, sharp):
    """FeatherBlender::createWeightMaps (blenders.cpp:158-186), incl. an all-zero mask (the `tmp` VIEW write-back)"""
    corners, _, masks = util.exposure_scene(n, w, h, seed=60 + n)
    if n == 4:
        masks[2][:] = 0
    rroi, rmaps = RF.feather_create_weight_maps(masks, corners, sharp)
    oroi, omaps = O.feather_create_weight_maps(masks, corners, sharp)
    assert tuple(rroi) == tuple(oroi)
    for a, b in zip(omaps, rmaps):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
