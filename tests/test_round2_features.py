"""Round-2 surface: warpBackward and PlaneWarper with a translation against the reference's own
code (oracle/_ref); the remaining projectors, the fisheye-undistort stage and the live app's crop margins INSIDE the
compositor, bit-exact against the oracle chain; the warper's map cache."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P
from stitchingvideo_b200 import rigs

from tests import util


def assert_same(got, ref, what):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    d = got != ref
    assert not d.any(), "%s: %d of %d values differ, first at %s" % (what, int(d.sum()), d.size, np.argwhere(d)[:3].tolist())


def narrow_rig(n=3, W=320, H=240, f=300.0, step=0.45):
    """n cameras fanned around the optical axis: fits the projectors that blow up far from it (stereographic, Panini ...)."""
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)
    Rs = [util.rot(0.03 * (i - 1), step * (i - (n - 1) / 2.0), 0.02 * i) for i in range(n)]
    return [K.copy() for _ in range(n)], Rs, (W, H)


def frames_for(n, W, H, seed):
    rng = np.random.default_rng(seed)
    return [util.smooth_image(rng, H, W) for _ in range(n)]


# ------------------------------------------------------------------------------------------------- CPU
def test_crop_geometry_follows_the_apps_float_arithmetic():
    # APP64:47 defaults: upblack = downblack = 0.1f, leftblack = rightblack = 10
    w, h, xx, yy = P.crop_geometry(8047, 1106, 0.1, 0.1, 10, 10)
    keep = np.float32(np.float32(1) - np.float32(0.1)) - np.float32(0.1)
    assert (w, xx) == (8027, 10)
    assert h == int(np.float32(1106) * keep) and yy == int(np.float32(np.float32(h) / keep) * np.float32(0.1))
    assert yy + h <= 1106
    assert P.crop_geometry(100, 50, 0.0, 0.0, 0, 0) == (100, 50, 0, 0)


def test_app_composite_equals_noblend_where_covered():
    """feedSize + feedSizeRemap (APP64:115-177) give Blender::NO's panorama wherever a camera covers the pixel."""
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    frames = [rigs.frame("mini_cyl", 0, i) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender="no")
    app, amask = P.compose_app(cal, frames, fill=False)
    assert np.array_equal(app, ref) and np.array_equal(amask, rmask)
    filled, _ = P.compose_app(cal, frames, fill=True)
    assert np.array_equal(filled[rmask != 0], ref[rmask != 0])


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,ab", [("spherical", None), ("cylindrical", None), ("plane", None), ("fisheye", None), ("mercator", None),
                                      ("paniniPortrait", (1.5, 0.8))])
def test_warp_backward_against_reference_sources(gpu, name, ab):
    """RotationWarperBase<P>::warpBackward (warpers_inl.hpp:102-128): exported since round 1, tested now."""
    from oracle import ref as RF
    if not RF.available():
        pytest.skip("oracle/_ref not built")
    cls = getattr(gpu, name[0].upper() + name[1:] + "Warper")
    rng = np.random.default_rng(5)
    for trial in range(2):
        W, H = int(rng.integers(120, 220)), int(rng.integers(90, 160))
        K, R = util.random_camera(rng, W, H, yaw=float(rng.uniform(-0.3, 0.3)))
        scale = float(rng.uniform(150, 300))
        w = cls(scale, *ab) if ab else cls(scale)
        rw = RF.Warper(name, scale, *(ab or (1.0, 1.0)))
        img = util.smooth_image(rng, H, W)
        tl, warped = w.warp(img, K, R)
        for interp, border in ((O.INTER_LINEAR, O.BORDER_REFLECT), (O.INTER_NEAREST, O.BORDER_CONSTANT)):
            back = w.warpBackward(warped, K, R, interp, border, (W, H))
            assert_same(back, rw.warp_backward(warped, K, R, (W, H), interp, border), "%s warpBackward" % name)
            assert_same(w.warpBackward(warped, K, R, interp, border, (W, H)), back, "%s warpBackward (cached maps)" % name)


@pytest.mark.gpu
def test_plane_warper_with_translation(gpu):
    """PlaneWarper's T overloads (warpers.cpp:81-137) against the reference's own code."""
    from oracle import ref as RF
    if not RF.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(9)
    for trial in range(3):
        W, H = int(rng.integers(120, 240)), int(rng.integers(90, 180))
        K, R = util.random_camera(rng, W, H, yaw=float(rng.uniform(-0.3, 0.3)))
        scale = float(rng.uniform(150, 300))
        T = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
        w = gpu.PlaneWarper(scale)
        w.setTranslation(T)
        rw = RF.PlaneWarperT(scale, T)
        assert w.warpRoi((W, H), K, R) == rw.warp_roi((W, H), K, R)
        u, v = w.warpPoint((W / 3.0, H / 4.0), K, R)
        ru, rv = rw.warp_point((W / 3.0, H / 4.0), K, R)
        assert (np.float32(u), np.float32(v)) == (np.float32(ru), np.float32(rv))
        roi, xm, ym = w.buildMaps((W, H), K, R)
        rroi, rxm, rym = rw.build_maps((W, H), K, R)
        assert tuple(roi) == tuple(rroi)
        assert_same(xm, rxm, "plane+T xmap")
        assert_same(ym, rym, "plane+T ymap")
        img = util.smooth_image(rng, H, W)
        (tl, dst), (rtl, rdst) = w.warp(img, K, R), rw.warp(img, K, R)
        assert tuple(tl) == tuple(rtl)
        assert_same(dst, rdst, "plane+T warp")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cylindrical", "stereographic"])
def test_repeated_warp_reuses_the_maps(gpu, name):
    """warp() rebuilt the maps on every call - for the host-evaluated projectors a per-pixel libm loop plus an upload.
    Same (size, K, R, scale) -> the maps are reused: one kernel (the remap) instead of two, or no host loop at all."""
    cls = getattr(gpu, name[0].upper() + name[1:] + "Warper")
    rng = np.random.default_rng(3)
    W, H = 200, 150
    K, R = util.random_camera(rng, W, H, yaw=0.2)
    w = cls(220.0)
    img = util.smooth_image(rng, H, W)
    n0 = gpu.kernel_launch_count()
    tl1, d1 = w.warp(img, K, R)
    n1 = gpu.kernel_launch_count()
    tl2, d2 = w.warp(img, K, R)
    n2 = gpu.kernel_launch_count()
    assert tl1 == tl2 and np.array_equal(d1, d2)
    assert n2 - n1 == 1 and n1 - n0 >= n2 - n1
    K2 = K.copy(); K2[0, 2] += 1.0                                  # another calibration: rebuilt
    tl3, d3 = w.warp(img, K2, R)
    assert not (tl3 == tl1 and np.array_equal(d3, d1))


@pytest.mark.gpu
@pytest.mark.parametrize("name,ab", [("stereographic", None), ("mercator", None), ("fisheye", None), ("paniniPortrait", (1.5, 0.8)),
                                      ("compressedRectilinear", (1.2, 1.1)), ("transverseMercator", None), ("cylindricalPortrait", None)])
@pytest.mark.parametrize("blender", ["feather", "multiband", "no"])
def test_compositor_with_the_remaining_projectors(gpu, name, ab, blender):
    """every projector of detail/warpers.hpp inside the per-frame compositor (maps evaluated once per calibration
    as the reference does, tables on the device, frames never touch the host), against the oracle chain fed with the
    reference's own buildMaps (oracle/_ref)."""
    from oracle import ref as RF
    if not RF.available():
        pytest.skip("oracle/_ref not built")
    Ks, Rs, size = narrow_rig()
    scale = 280.0
    rw = RF.Warper(name, scale, *(ab or (1.0, 1.0)))
    cal = P.Calibration(size, Ks, Rs, None, scale, build_maps=rw.build_maps)
    gains = [0.95, 1.02, 1.05]
    comp = gpu.Compositor(size, Ks, Rs, warper=name, scale=scale, blender=blender, gains=gains, warper_ab=ab)
    for i in range(3):
        roi = comp.camera_roi(i)
        assert (roi[0], roi[1]) == cal.corners[i] and (roi[2], roi[3]) == cal.sizes[i]
    for fi in range(2):
        frames = frames_for(3, size[0], size[1], 40 + fi)
        ref, rmask = P.compose(cal, frames, blender=blender, gains=gains)
        pano, mask = comp.compose(frames)
        assert_same(pano, ref, "%s/%s pano" % (name, blender))
        assert_same(mask, rmask, "%s/%s mask" % (name, blender))
    comp.set_fused(0)                                               # the staged, reference-shaped path recomputes mapBackward on the device
    with pytest.raises(gpu.StitchError) as e:
        comp.compose(frames)
    assert e.value.code == -213


def fisheye_maps(W, H, k1=-0.30, k2=0.09):
    """A radial undistortion map pair in initUndistortRectifyMap's CV_16SC2 / CV_16UC1 format (APP64:201-238), without cv2."""
    f = 0.55 * W
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    x, y = (xs - W / 2) / f, (ys - H / 2) / f
    r2 = x * x + y * y
    d = 1 + k1 * r2 + k2 * r2 * r2
    return O.convert_maps((x * d * f + W / 2).astype(np.float32), (y * d * f + H / 2).astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("rig,blender,kw", [("mini_cyl", "no", {}), ("mini_cyl", "feather", {}), ("mini", "multiband", {}),
                                            ("mini_cyl", "no", {"blocks": True})])
def test_undistort_stage_inside_the_compositor(gpu, rig, blender, kw):
    """the live app's frame loop (APP64:736-756) - remap through the fisheye maps, THEN the cached-map warp of the
    8-bit result, gain, composite - as one device pipeline, bit-exact against the oracle's remap_fixed -> remap chain."""
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    umaps = [fisheye_maps(size[0], size[1], -0.30 + 0.02 * i, 0.09) for i in range(n)]
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    gains = ([0.95, 1.02, 1.0, 0.98, 1.05] * 2)[:n]
    gmaps = None
    if kw.get("blocks"):
        rng = np.random.default_rng(1)
        gmaps = [rng.uniform(0.8, 1.2, ((s[1] + 31) // 32, (s[0] + 31) // 32)).astype(np.float32) for s in cal.sizes]
        gains = None
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=gains, gain_maps=gmaps,
                          undistort_maps=umaps)
    comp.set_depth(2)
    for fi in range(2):
        frames = [rigs.frame(rig, fi, i) for i in range(n)]
        ref, rmask = P.compose(cal, frames, blender=blender, gains=gains, gain_maps=gmaps, undistort_maps=umaps)
        pano, mask = comp.compose(frames)
        assert_same(pano, ref, "undistort %s/%s frame %d" % (rig, blender, fi))
        assert_same(mask, rmask, "undistort mask")
    # without the stage the panorama differs (the maps are not the identity)
    plain = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=gains, gain_maps=gmaps)
    assert not np.array_equal(plain.compose(frames)[0], pano)


@pytest.mark.gpu
@pytest.mark.parametrize("crop", [(0.1, 0.1, 10, 10), (0.0, 0.25, 0, 37), (0.2, 0.0, 33, 1)])
@pytest.mark.parametrize("fill", [True, False])
def test_crop_margins_of_the_app_composite(gpu, crop, fill):
    """UpdateMat / feedSizeRemap's crop (APP64:47, 150-177, 702) - cropped tiles are never scheduled - incl. the
    unconditional gather: a pixel no camera covers takes camera 0's warped pixel (0, 0)."""
    rig = "mini_cyl"
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    # seam masks leave uncovered panorama pixels, so that the fill is exercised
    cal0 = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    seams = [rigs.seam_mask(cal0.sizes[i], 0.2, 0.8) for i in range(n)]
    for m in seams:
        m[:5] = 0
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"], seams)
    rng = np.random.default_rng(2)
    gmaps = [rng.uniform(0.8, 1.2, ((s[1] + 31) // 32, (s[0] + 31) // 32)).astype(np.float32) for s in cal.sizes]
    comp = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="no", gain_maps=gmaps, seam_masks=seams,
                          crop=crop, crop_app_fill=fill)
    roi = O.result_roi(cal.corners, cal.sizes)
    ow, oh, xx, yy = P.crop_geometry(roi[2], roi[3], *crop)
    assert comp.pano_size == (ow, oh)
    for fi in range(2):
        frames = [rigs.frame(rig, fi, i) for i in range(n)]
        ref, rmask = P.compose_app(cal, frames, gain_maps=gmaps, crop=crop, fill=fill)
        pano, mask = comp.compose(frames)
        assert_same(pano, ref, "cropped composite fill=%s" % fill)
        assert_same(mask, rmask[yy:yy + oh, xx:xx + ow], "cropped mask")
    if fill:
        assert (rmask[yy:yy + oh, xx:xx + ow] == 0).any(), "the scene must contain uncovered pixels for this test to mean anything"


@pytest.mark.gpu
def test_crop_margins_with_feather_and_as_a_persistent_lap(gpu):
    import torch
    from stitchingvideo_b200 import capi
    rig, crop = "mini_cyl", (0.1, 0.1, 10, 10)
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    comp = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather", crop=crop)
    roi = O.result_roi(cal.corners, cal.sizes)
    ow, oh, xx, yy = P.crop_geometry(roi[2], roi[3], *crop)
    sets = [[rigs.frame(rig, f, i) for i in range(n)] for f in range(2)]
    refs = [P.compose(cal, s, blender="feather") for s in sets]
    for s, (ref, rmask) in zip(sets, refs):
        pano, mask = comp.compose(s)
        assert_same(pano, ref[yy:yy + oh, xx:xx + ow], "cropped feather")
        assert_same(mask, rmask[yy:yy + oh, xx:xx + ow], "cropped feather mask")
    dev = [[capi.DeviceImage.from_torch(torch.from_numpy(a).cuda()) for a in s] for s in sets]
    wp = (ow + 7) & ~7
    pt = [torch.zeros((oh, wp, 3), dtype=torch.uint8, device="cuda") for _ in range(3)]
    panos = [capi.DeviceImage(t.data_ptr(), oh, ow, gpu.CV_8UC3, wp * 3, 0, owner=t) for t in pt]
    b = comp.batch([dev[f % 2] for f in range(3)], panos)
    assert b.mode == 1
    b.launch(); b.wait()
    for f in range(3):
        assert_same(pt[f][:, :ow].cpu().numpy(), refs[f % 2][0][yy:yy + oh, xx:xx + ow], "cropped persistent lap frame %d" % f)
    with pytest.raises(gpu.StitchError):
        gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="multiband", crop=crop)
