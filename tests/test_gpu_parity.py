"""-m gpu parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for every stage (the float-weight stages included, because the GPU
follows the oracle's summation order; north_star only asks for +-1 LSB there)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P
from tests import util

pytestmark = pytest.mark.gpu


def assert_same(got, ref, what):
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    if got.dtype == np.float32:
        d = got.view(np.uint32) != ref.view(np.uint32)
    else:
        d = got != ref
    assert not d.any(), "%s: %d of %d values differ, first at %s" % (what, int(d.sum()), d.size, np.argwhere(d)[:3].tolist())


# ------------------------------------------------------------------ a1-a3 warper geometry + maps
@pytest.mark.parametrize("kind", ["spherical", "cylindrical", "plane"])
def test_build_maps_bit_exact(gpu, kind):
    rng = np.random.default_rng(10)
    cls = {"spherical": gpu.SphericalWarper, "cylindrical": gpu.CylindricalWarper, "plane": gpu.PlaneWarper}[kind]
    for _ in range(4):
        W, H = int(rng.integers(200, 700)), int(rng.integers(150, 500))
        K, R = util.random_camera(rng, W, H)
        scale = float(rng.uniform(300, 900))
        w, ow = cls(scale), O.Warper(kind, scale)
        roi, xm, ym = w.buildMaps((W, H), K, R)
        oroi, oxm, oym = ow.build_maps((W, H), K, R)
        assert tuple(roi) == tuple(oroi)
        assert_same(xm, oxm, kind + " xmap")
        assert_same(ym, oym, kind + " ymap")
        assert w.warpRoi((W, H), K, R) == ow.warp_roi((W, H), K, R)
        u, v = w.warpPoint((W / 3.0, H / 5.0), K, R)
        ou, ov = ow.warp_point((W / 3.0, H / 5.0), K, R)
        assert (np.float32(u), np.float32(v)) == (np.float32(ou), np.float32(ov))
        assert w.getScale() == np.float32(scale)


def test_seam_straddling_camera_roi(gpu):
    """A camera centred on yaw 180deg gets a panorama-wide ROI."""
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras("mini")
    w, ow = gpu.SphericalWarper(spec["scale"]), O.Warper("spherical", spec["scale"])
    roi = w.warpRoi((spec["W"], spec["H"]), Ks[0], Rs[0])
    assert roi == ow.warp_roi((spec["W"], spec["H"]), Ks[0], Rs[0])
    assert roi[2] > 3 * w.warpRoi((spec["W"], spec["H"]), Ks[1], Rs[1])[2]


# ------------------------------------------------------------------ a4 remap
@pytest.mark.parametrize("cn", [1, 3])
@pytest.mark.parametrize("border", [O.BORDER_REFLECT, O.BORDER_CONSTANT, O.BORDER_REPLICATE, O.BORDER_REFLECT_101, O.BORDER_WRAP])
@pytest.mark.parametrize("interp", [O.INTER_LINEAR, O.INTER_NEAREST])
def test_remap_bit_exact(gpu, cn, border, interp):
    rng = np.random.default_rng(20 + cn)
    H, W = 97, 131
    src = rng.integers(0, 256, (H, W, cn), dtype=np.uint8) if cn == 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
    xm, ym = util.special_maps(rng, 120, 173, W, H)
    got = gpu.remap(src, xm, ym, interp, border, (7, 9, 11, 0))
    ref = O.remap(src, xm, ym, interp, border, (7, 9, 11, 0))
    assert_same(got, ref, "remap")


@pytest.mark.parametrize("cn", [1, 3])
def test_remap_fixed_point_maps(gpu, cn):
    """convertMaps + remap through the CV_16SC2 / CV_16UC1 pair: the app's fisheye front-end format (APP64:201-238, 741)."""
    rng = np.random.default_rng(22 + cn)
    H, W = 97, 131
    src = rng.integers(0, 256, (H, W, cn), dtype=np.uint8) if cn == 3 else rng.integers(0, 256, (H, W), dtype=np.uint8)
    xm, ym = util.special_maps(rng, 120, 173, W, H)
    m1, m2 = gpu.convertMaps(xm, ym)
    om1, om2 = O.convert_maps(xm, ym)
    assert_same(m1, om1, "convertMaps map1")
    assert_same(m2, om2, "convertMaps map2")
    n1, _ = gpu.convertMaps(xm, ym, nninterpolation=True)
    assert_same(n1, O.convert_maps(xm, ym, True)[0], "convertMaps nearest")
    for border in (O.BORDER_REFLECT, O.BORDER_CONSTANT, O.BORDER_REPLICATE, O.BORDER_REFLECT_101, O.BORDER_WRAP):
        for interp in (O.INTER_LINEAR, O.INTER_NEAREST):
            assert_same(gpu.remap(src, m1, m2, interp, border, (7, 9, 11, 0)), O.remap_fixed(src, m1, m2, interp, border, (7, 9, 11, 0)),
                        "fixed remap b=%d i=%d" % (border, interp))
        assert_same(gpu.remap(src, n1, None, O.INTER_NEAREST, border, (7, 9, 11, 0)), O.remap_fixed(src, n1, None, O.INTER_NEAREST, border, (7, 9, 11, 0)),
                    "fixed nearest b=%d" % border)
        # the fixed-point pair gives what the float maps give (cv::remap converts them the same way internally)
        assert_same(gpu.remap(src, m1, m2, O.INTER_LINEAR, border, (7, 9, 11, 0)), gpu.remap(src, xm, ym, O.INTER_LINEAR, border, (7, 9, 11, 0)),
                    "fixed vs float maps b=%d" % border)
    with pytest.raises(gpu.StitchError):
        gpu.remap(src, m1, None, O.INTER_LINEAR, O.BORDER_REFLECT)      # OpenCV needs the fractional map for INTER_LINEAR


def test_remap_degenerate_sources(gpu):
    rng = np.random.default_rng(21)
    for (H, W) in ((1, 1), (1, 9), (7, 1), (2, 2)):
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        xm = rng.uniform(-3, W + 3, (16, 24)).astype(np.float32)
        ym = rng.uniform(-3, H + 3, (16, 24)).astype(np.float32)
        for border in (O.BORDER_REFLECT, O.BORDER_CONSTANT):
            assert_same(gpu.remap(src, xm, ym, O.INTER_LINEAR, border), O.remap(src, xm, ym, O.INTER_LINEAR, border), "remap %dx%d" % (H, W))


@pytest.mark.parametrize("kind", ["spherical", "cylindrical", "plane"])
def test_warp_image_and_mask(gpu, kind):
    rng = np.random.default_rng(30)
    cls = {"spherical": gpu.SphericalWarper, "cylindrical": gpu.CylindricalWarper, "plane": gpu.PlaneWarper}[kind]
    W, H = 320, 200
    K, R = util.random_camera(rng, W, H, yaw=0.3)
    img = util.smooth_image(rng, H, W)
    w, ow = cls(260.0), O.Warper(kind, 260.0)
    tl, dst = w.warp(img, K, R, O.INTER_LINEAR, O.BORDER_REFLECT)
    otl, odst = ow.warp(img, K, R, O.INTER_LINEAR, O.BORDER_REFLECT)
    assert tuple(tl) == tuple(otl)
    assert_same(dst, odst, "warp image")
    mask = np.full((H, W), 255, np.uint8)
    tl, dm = w.warp(mask, K, R, O.INTER_NEAREST, O.BORDER_CONSTANT)
    otl, odm = ow.warp(mask, K, R, O.INTER_NEAREST, O.BORDER_CONSTANT)
    assert_same(dm, odm, "warp mask")
    # cached-map video path (APP64:752)
    w.buildMaps((W, H), K, R)
    assert_same(w.remap(img), odst, "cached remap")


# ------------------------------------------------------------------ a5-a7 exposure
def test_gain_apply(gpu):
    rng = np.random.default_rng(40)
    img = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    img[0, :86 // 3] = 255
    c = gpu.GainCompensator()
    gains = [0.95, 1.02, 1.0, 0.5, 1.5, 2.5, 0.98, 1.05]
    c.setGains(gains)
    assert c.gains() == gains
    for i, g in enumerate(gains):
        got = img.copy()
        c.apply(i, (0, 0), got, None)
        assert_same(got, O.gain_apply(img, g), "gain %g" % g)


def test_blocks_gain_apply(gpu):
    rng = np.random.default_rng(41)
    img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    gm = rng.uniform(0.7, 1.4, (9, 15)).astype(np.float32)
    full = rng.uniform(0.7, 1.4, (270, 480)).astype(np.float32)
    c = gpu.BlocksGainCompensator()
    c.setGainMaps([gm, full])
    for i, m in enumerate((gm, full)):
        got = img.copy()
        c.apply(i, (0, 0), got, None)
        assert_same(got, O.blocks_gain_apply(img, m), "blocks gain %d" % i)


# ------------------------------------------------------------------ a11/a11b/a15 pyramids
@pytest.mark.parametrize("dtype", [np.int16, np.uint8])
@pytest.mark.parametrize("shape,levels", [((64, 96), 5), ((160, 224), 3), ((32, 32), 5), ((8, 24), 2), ((40, 56), 0)])
def test_laplace_pyramid(gpu, dtype, shape, levels):
    rng = np.random.default_rng(50)
    if dtype == np.int16:
        img = rng.integers(-32768, 32768, shape + (3,)).astype(np.int16)     # full range: saturation paths
    else:
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    got = gpu.createLaplacePyr(img, levels)
    ref = O.create_laplace_pyr(img, levels)
    for l, (g, r) in enumerate(zip(got, ref)):
        assert_same(g, r, "laplace level %d" % l)
    assert_same(gpu.restoreImageFromLaplacePyr(ref), O.restore_from_laplace_pyr(ref), "restore")


# ------------------------------------------------------------------ a12-a18 blenders
def _run_pair(gpu, kind, imgs, masks, tls, **kw):
    sizes = [(im.shape[1], im.shape[0]) for im in imgs]
    if kind == "no":
        gb, ob = gpu.Blender(), O.Blender(O.BLEND_NO)
    elif kind == "feather":
        gb, ob = gpu.FeatherBlender(kw.get("sharpness", 0.02)), O.Blender(O.BLEND_FEATHER, sharpness=kw.get("sharpness", 0.02))
    else:
        gb = gpu.MultiBandBlender(False, kw.get("num_bands", 5), kw.get("weight_type", O.CV_32F))
        ob = O.Blender(O.BLEND_MULTI_BAND, kw.get("num_bands", 5), kw.get("weight_type", O.CV_32F))
    gb.prepare(tls, sizes)
    ob.prepare(tls, sizes)
    for im, m, tl in zip(imgs, masks, tls):
        gb.feed(im, m, tl)
        ob.feed(im, m, tl)
    return gb.blend(), ob.blend()


@pytest.mark.parametrize("scene", [(2, 100, 150, 80), (5, 200, 300, 500), (3, 60, 70, 30), (4, 373, 600, 900)])
@pytest.mark.parametrize("cfg", [("no", {}), ("feather", {}), ("feather", {"sharpness": 0.1}),
                                 ("mb", {"num_bands": 5}), ("mb", {"num_bands": 5, "weight_type": O.CV_16S}),
                                 ("mb", {"num_bands": 1}), ("mb", {"num_bands": 3, "weight_type": O.CV_16S}),
                                 ("mb", {"num_bands": 7}), ("mb", {"num_bands": 0})])
def test_blenders_bit_exact(gpu, scene, cfg):
    rng = np.random.default_rng(60 + scene[0])
    imgs, masks, tls = util.blend_scene(rng, *scene)
    (gd, gm), (od, om) = _run_pair(gpu, cfg[0], imgs, masks, tls, **cfg[1])
    assert_same(gd, od, "blend image %s" % (cfg,))
    assert_same(gm, om, "blend mask %s" % (cfg,))


def test_multiband_8u_input(gpu):
    rng = np.random.default_rng(61)
    imgs, masks, tls = util.blend_scene(rng, 3, 120, 200, 150, dtype=np.uint8)
    for wt in (O.CV_32F, O.CV_16S):
        (gd, gm), (od, om) = _run_pair(gpu, "mb", imgs, masks, tls, num_bands=4, weight_type=wt)
        assert_same(gd, od, "8U multiband image")
        assert_same(gm, om, "8U multiband mask")


def test_multiband_saturating_input(gpu):
    """Full-range int16 images exercise the saturating subtract/add and the wrapping accumulate."""
    rng = np.random.default_rng(62)
    imgs, masks, tls = util.blend_scene(rng, 3, 90, 130, 100)
    imgs = [rng.integers(-32768, 32768, im.shape).astype(np.int16) for im in imgs]
    (gd, gm), (od, om) = _run_pair(gpu, "mb", imgs, masks, tls, num_bands=4)
    assert_same(gd, od, "saturating multiband")


def test_blender_reuse_and_errors(gpu):
    rng = np.random.default_rng(63)
    imgs, masks, tls = util.blend_scene(rng, 2, 80, 100, 60)
    b = gpu.MultiBandBlender()
    with pytest.raises(gpu.StitchError) as e:
        b.feed(imgs[0], masks[0], tls[0])                       # feed before prepare
    assert e.value.code == -215
    b.prepare(tls, [(im.shape[1], im.shape[0]) for im in imgs])
    with pytest.raises(gpu.StitchError) as e:
        b.feed(imgs[0].astype(np.float32), masks[0], tls[0])    # CV_Assert(img.type() == CV_16SC3 || CV_8UC3)
    assert e.value.code == -215
    with pytest.raises(gpu.StitchError):
        b.feed(imgs[0], masks[0].astype(np.int16), tls[0])      # CV_Assert(mask.type() == CV_8U)
    for im, m, tl in zip(imgs, masks, tls):
        b.feed(im, m, tl)
    d1, m1 = b.blend()
    with pytest.raises(gpu.StitchError):
        b.blend()                                               # buffers were handed over: prepare again
    b.prepare(tls, [(im.shape[1], im.shape[0]) for im in imgs])
    for im, m, tl in zip(imgs, masks, tls):
        b.feed(im, m, tl)
    d2, m2 = b.blend()
    assert_same(d2, d1, "second panorama on the same handle")
    with pytest.raises(gpu.StitchError) as e:
        gpu.Blender.createDefault(7)
    assert e.value.code == -5                                   # CV_StsBadArg
    with pytest.raises(gpu.StitchError) as e:
        gpu.ExposureCompensator.createDefault(9)
    assert e.value.code == -5
    with pytest.raises(gpu.StitchError) as e:
        gpu.MultiBandBlender(False, 5, O.CV_8U)                 # weight_type assert (blenders.cpp:198)
    assert e.value.code == -215
    with pytest.raises(gpu.StitchError) as e:
        gpu.SphericalWarper(100.0).warp(np.zeros((4, 4, 3), np.uint8), np.eye(2, dtype=np.float32), np.eye(3, dtype=np.float32))
    assert e.value.code == -215                                 # K must be 3x3 CV_32F (warpers.cpp:52)


def test_weight_map_and_normalize(gpu):
    rng = np.random.default_rng(64)
    m = np.zeros((80, 120), np.uint8)
    m[10:70, 15:100] = 255
    m[30:35, 40:50] = 0
    m[0:3, :] = 255
    for sharp in (0.02, 0.1, 1.0):
        assert_same(gpu.createWeightMap(m, sharp), O.create_weight_map(m, sharp), "createWeightMap")
    assert_same(gpu.createWeightMap(np.full((40, 50), 255, np.uint8), 0.02), O.create_weight_map(np.full((40, 50), 255, np.uint8), 0.02), "all-ones mask")
    assert_same(gpu.createWeightMap(np.zeros((5, 7), np.uint8), 0.02), O.create_weight_map(np.zeros((5, 7), np.uint8), 0.02), "all-zero mask")
    src = rng.integers(-3000, 3000, (33, 47, 3)).astype(np.int16)
    wf = rng.uniform(0, 3, (33, 47)).astype(np.float32)
    wf[0, :5] = [0, 1e-7, 1e-6, 1e-5, 1.0]                       # tiny weights: out-of-range float->short casts
    ws = rng.integers(0, 700, (33, 47)).astype(np.int16)
    for w in (wf, ws):
        got = gpu.normalizeUsingWeightMap(w, src.copy())
        ref = src.copy()
        mw, ms = O.mat(w), O.mat(ref)
        import ctypes as C
        O.lib().so_normalize_using_weight_map(C.byref(mw), C.byref(ms))
        assert_same(got, ref, "normalizeUsingWeightMap %s" % w.dtype)


# ------------------------------------------------------------------ the per-frame loop
def _compositor_case(gpu, rig, blender, weight_type=O.CV_32F, seams=False, gains=True, num_bands=5, out16=False):
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras(rig)
    size = (spec["W"], spec["H"])
    n = spec["n_used"]
    g = ([0.95, 1.02, 1.0, 0.98, 1.05] * 2)[:n] if gains else None
    cal0 = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    seam = None
    if seams:
        seam = [rigs.seam_mask(cal0.sizes[i], 0.15, 0.85) for i in range(n)]
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"], seam)
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, num_bands=num_bands,
                          weight_type=weight_type, gains=g, seam_masks=seam,
                          output_type=gpu.CV_16SC3 if out16 else gpu.CV_8UC3)
    for i in range(n):
        roi = comp.camera_roi(i)
        assert (roi[0], roi[1]) == cal.corners[i] and (roi[2], roi[3]) == cal.sizes[i]
    for fi in range(2):                       # two frames through the same handle (tables stay resident)
        frames = [rigs.frame(rig, fi, i) for i in range(n)]
        ref, rmask = P.compose(cal, frames, blender=blender, num_bands=num_bands, weight_type=weight_type, gains=g,
                               output_8u=not out16)
        # 11: fused fast kernels (RGBX pyramid with multi-level launches / streaming feather); 12 / 13: the same with one
        # launch per pyramid level / with the multi-level launches forced; 14: the same with the gather warp stage; 10: fused CV_16S band kernels / one-pixel-per-thread feather; 0: the staged,
        # camera-by-camera path shaped like the reference's feed/blend calls
        for fused in (11, 15, 12, 13, 14, 10, 0):
            comp.set_fused(fused)
            pano, mask = comp.compose(frames)
            assert_same(pano, ref, "%s/%s pano frame %d fused=%s" % (rig, blender, fi, fused))
            assert_same(mask, rmask, "%s/%s mask frame %d fused=%s" % (rig, blender, fi, fused))


@pytest.mark.parametrize("case", [
    dict(rig="mini", blender="multiband"),
    dict(rig="mini", blender="multiband", weight_type=O.CV_16S, seams=True),
    dict(rig="mini", blender="multiband", seams=True, out16=True),
    dict(rig="mini", blender="multiband", num_bands=2, gains=False),
    dict(rig="mini_cyl", blender="feather", gains=False),
    dict(rig="mini_cyl", blender="feather", seams=True),
    dict(rig="mini_cyl", blender="no", gains=False),
    dict(rig="mini", blender="no", seams=True),
    # streaming frame kernel shapes: direct gathers (boxes beyond shared memory), large / tiny boxes, 3 cameras per tile
    dict(rig="mini_cyl_s3", blender="feather", gains=False),
    dict(rig="mini_cyl_s3", blender="no", gains=False),
    dict(rig="mini_cyl_s17", blender="feather", seams=True),
    dict(rig="mini_cyl_s17", blender="multiband"),
    dict(rig="mini_cyl_up", blender="feather", gains=False),
    dict(rig="mini_cyl_n9", blender="feather"),
    dict(rig="mini_cyl_n9", blender="no", gains=False),
    dict(rig="mini_sph_n7", blender="feather", out16=True),
])
def test_compositor_matches_reference_loop(gpu, case):
    _compositor_case(gpu, **case)


def test_shared_divisor_division_selftest(gpu):
    """The 3-FFMA shared-reciprocal quotient used inside the fused kernels equals div.rn.f32."""
    assert gpu.capi.selftest_division(1 << 26, seed=7) == 0
    assert gpu.capi.selftest_division(1 << 24, seed=12345) == 0


def test_compositor_pipelined_slots(gpu):
    """Frames in flight on separate slots give the same panoramas as the synchronous call."""
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras("mini")
    size = (spec["W"], spec["H"])
    comp = gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", gains=spec["gain_values"])
    sets = [[rigs.frame("mini", fi, i) for i in range(5)] for fi in range(4)]
    want = [comp.compose(s)[0].copy() for s in sets]
    comp.set_depth(3)
    outs = [comp.new_output() for _ in sets]
    slots = []
    for k, s in enumerate(sets):
        if k >= 3:
            comp.wait(slots[k - 3])
        slots.append(comp.enqueue(s, outs[k][0], outs[k][1]))
    for sl in slots:
        comp.wait(sl)
    for k in range(len(sets)):
        assert_same(outs[k][0], want[k], "pipelined frame %d" % k)
    assert comp.last_gpu_ms(0) > 0


def test_frame_set_in_one_host_block(gpu):
    """A frame set that is contiguous in host memory (one DMA) gives the same panorama as separate images."""
    from stitchingvideo_b200 import rigs
    for rig, blender in (("mini", "multiband"), ("mini_cyl", "feather")):
        Ks, Rs, spec = rigs.cameras(rig)
        n, size = spec["n_used"], (spec["W"], spec["H"])
        comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=spec["gain_values"])
        frames = [rigs.frame(rig, 3, i) for i in range(n)]
        want, wmask = comp.compose(frames)
        block = np.ascontiguousarray(np.stack(frames))          # (n, H, W, 3): camera i+1 starts where camera i ends
        got, gmask = comp.compose([block[i] for i in range(n)])
        assert_same(got, want, rig + " one-block frame set")
        assert_same(gmask, wmask, rig + " one-block mask")


@pytest.mark.parametrize("rig,blender", [("mini", "multiband"), ("mini_cyl", "feather"), ("mini_cyl", "no")])
def test_compositor_blocks_gain(gpu, rig, blender):
    """BlocksGainCompensator::apply inside the fused frame kernels (the live app's BlockApply, APP64:310-331, 754):
    per-camera block gain maps, resized once with INTER_LINEAR, multiplied in per pixel."""
    from stitchingvideo_b200 import rigs
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    rng = np.random.default_rng(70)
    comp0 = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender)
    gmaps = []
    for i in range(n):
        x, y, w, h = comp0.camera_roi(i)
        gmaps.append(rng.uniform(0.7, 1.4, ((h + 31) // 32, (w + 31) // 32)).astype(np.float32))     # 32x32 blocks (exposure_compensate.cpp:165-170)
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gain_maps=gmaps)
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    frames = [rigs.frame(rig, 5, i) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender=blender, gain_maps=gmaps)
    # every kernel variant, incl. the staged, reference-shaped cross-check path (0: warp -> mul by the resized map -> convertTo ->
    # feed x n -> blend) and the CV_16S band kernels / gather feather kernel (10) and the round-1 streaming kernel (15)
    for fused in ((11, 12, 13, 14, 16, 17, 18, 19, 10, 0) if blender == "multiband" else (11, 15, 10, 0)):
        comp.set_fused(fused)
        pano, mask = comp.compose(frames)
        assert_same(pano, ref, "%s/%s blocks gain (variant %d)" % (rig, blender, fused))
        assert_same(mask, rmask, "%s/%s blocks gain mask (variant %d)" % (rig, blender, fused))


@pytest.mark.parametrize("name,ab", [("fisheye", None), ("stereographic", None), ("compressedRectilinear", (1.5, 1.0)),
                                     ("compressedRectilinearPortrait", (2.0, 1.0)), ("panini", (1.5, 1.0)), ("paniniPortrait", (2.0, 1.0)),
                                     ("mercator", None), ("transverseMercator", None), ("sphericalPortrait", None),
                                     ("cylindricalPortrait", None), ("planePortrait", None)])
def test_remaining_projectors_against_reference_sources(gpu, name, ab):
    """The other projectors of detail/warpers.hpp: host-built maps + the CUDA remap, against the
    reference's own RotationWarperBase<P> code compiled into oracle/_ref."""
    from oracle import ref as RF
    if not RF.available():
        pytest.skip("oracle/_ref not built")
    cls = getattr(gpu, name[0].upper() + name[1:] + "Warper")
    rng = np.random.default_rng(80)
    for trial in range(3):
        W, H = int(rng.integers(120, 260)), int(rng.integers(90, 200))
        K, R = util.random_camera(rng, W, H, yaw=float(rng.uniform(-0.5, 0.5)))
        scale = float(rng.uniform(150, 400))
        w = cls(scale, *ab) if ab else cls(scale)
        rw = RF.Warper(name, scale, *(ab or (1.0, 1.0)))
        assert w.warpRoi((W, H), K, R) == rw.warp_roi((W, H), K, R), "%s warpRoi" % name
        u, v = w.warpPoint((W / 3.0, H / 5.0), K, R)
        ru, rv = rw.warp_point((W / 3.0, H / 5.0), K, R)
        assert (np.float32(u), np.float32(v)) == (np.float32(ru), np.float32(rv)), "%s warpPoint" % name
        roi, xm, ym = w.buildMaps((W, H), K, R)
        rroi, rxm, rym = rw.build_maps((W, H), K, R)
        assert tuple(roi) == tuple(rroi)
        assert_same(xm, rxm, name + " xmap")
        assert_same(ym, rym, name + " ymap")
        img = util.smooth_image(rng, H, W)
        (tl, dst), (rtl, rdst) = w.warp(img, K, R), rw.warp(img, K, R)
        assert tuple(tl) == tuple(rtl)
        assert_same(dst, rdst, name + " warp")


# ---- once-per-calibration steps --------------------------------------------------------------
@pytest.mark.parametrize("n,w,h", [(2, 120, 90), (3, 160, 100), (5, 200, 120), (4, 333, 251)])
def test_gain_compensator_feed(gpu, n, w, h):
    """GainCompensator::feed (exposure_compensate.cpp:76-147): N exact, sums exact-then-rounded on the device vs the
    reference's sequential double sums -> gains to 1e-11 relative (the tolerance of a double sum over <= 1e5 terms)."""
    corners, imgs, masks = util.exposure_scene(n, w, h, seed=10 + n)
    ref = O.gain_feed(corners, imgs, masks)
    c = gpu.GainCompensator()
    c.feed(corners, imgs, masks)
    got = np.array(c.gains())
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=0)
    c2 = gpu.GainCompensator()                                       # deterministic run to run (integer accumulation)
    c2.feed(corners, imgs, masks)
    assert c2.gains() == c.gains()
    # the estimated gains drive apply() like handed-in ones
    img = imgs[1].copy()
    c.apply(1, corners[1], img)
    assert np.array_equal(img, O.gain_apply(imgs[1], ref[1])) or np.float32(ref[1]) != np.float32(got[1])


def test_gain_compensator_feed_disjoint_and_errors(gpu):
    corners, imgs, masks = util.exposure_scene(3, 64, 48, seed=3)
    corners = [(0, 0), (1000, 0), (2000, 500)]                       # no overlaps: every gain is exactly 1
    c = gpu.GainCompensator()
    c.feed(corners, imgs, masks)
    assert c.gains() == [1.0, 1.0, 1.0]
    assert list(O.gain_feed(corners, imgs, masks)) == [1.0, 1.0, 1.0]
    with pytest.raises(gpu.StitchError):                             # CV_Assert(corners.size() == images.size() && ...)
        c.feed(corners[:2], imgs, masks)
    with pytest.raises(gpu.StitchError):
        c.feed(corners, [im[:, :, 0] for im in imgs], masks)


@pytest.mark.parametrize("n,w,h,bl", [(2, 120, 90, 32), (3, 160, 100, 32), (3, 150, 97, 20)])
def test_blocks_gain_compensator_feed(gpu, n, w, h, bl):
    """BlocksGainCompensator::feed (exposure_compensate.cpp:165-222): block lists, one gain solve over all blocks, two
    smoothing passes; float32 gain maps within 2 ulp of the oracle's (the double gains agree to 1e-11)."""
    corners, imgs, masks = util.exposure_scene(n, w, h, seed=20 + n)
    ref = O.blocks_gain_feed(corners, imgs, masks, bl, bl)
    c = gpu.BlocksGainCompensator()
    c.setBlockSize(bl, bl)
    c.feed(corners, imgs, masks)
    got = c.gainMaps()
    assert [m.shape for m in got] == [m.shape for m in ref]
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a, b, rtol=3e-7, atol=0)
    img = imgs[0].copy()                                             # and apply() uses them
    c.apply(0, corners[0], img)
    assert np.abs(img.astype(int) - O.blocks_gain_apply(imgs[0], ref[0]).astype(int)).max() <= 1


@pytest.mark.parametrize("shape,dsize", [((37, 53), (106, 74)), ((40, 64), (32, 20)), ((64, 2), (5, 100)), ((50, 70), (70, 50)),
                                         ((31, 45), (100, 77)), ((1, 9), (20, 3)), ((249, 389), (1555, 998))])
def test_seam_mask_refinement_bit_exact(gpu, shape, dsize):
    """stitcher.cpp:291-294: dilate -> resize(INTER_LINEAR) -> & ; and the two primitives on their own."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = (rng.random(shape) > 0.6).astype(np.uint8) * 255
    a[rng.random(shape) > 0.9] = rng.integers(1, 255)
    mw = (rng.random((dsize[1], dsize[0])) > 0.2).astype(np.uint8) * 255
    from stitchingvideo_b200 import capi
    assert_same(capi.dilate3x3(a), O.dilate3x3(a), "dilate")
    assert_same(capi.resize_linear_8u(a, dsize), O.resize_linear_8u(a, dsize), "resize")
    assert_same(capi.refine_seam_mask(a, mw), O.refine_seam_mask(a, mw), "refined seam mask")


@pytest.mark.parametrize("n,w,h,sharp", [(2, 120, 90, 0.02), (3, 160, 100, 0.05), (4, 90, 70, 0.3)])
def test_feather_create_weight_maps(gpu, n, w, h, sharp):
    """FeatherBlender::createWeightMaps (blenders.cpp:158-186) incl. an all-zero mask (sum < eps -> 1, written back into
    the shared sum) — bit for bit; the oracle is pinned against cv2 in tests/test_oracle_vs_cv2.py."""
    corners, _, masks = util.exposure_scene(n, w, h, seed=30 + n)
    if n == 4:
        masks[2][:] = 0
    oroi, omaps = O.feather_create_weight_maps(masks, corners, sharp)
    roi, maps = gpu.FeatherBlender(sharp).createWeightMaps(masks, corners)
    assert tuple(roi) == tuple(oroi)
    for a, b in zip(maps, omaps):
        assert_same(a, b, "normalised weight map")
    with pytest.raises(gpu.StitchError):
        gpu.FeatherBlender().createWeightMaps(masks, corners[:1])


def test_seam_mask_refinement_random_shapes(gpu):
    """dilate / 8U linear resize / refine over random shapes and scale factors (up, down, exact 2x in one or both
    dimensions, 1-pixel sizes) — bit for bit against the oracle (which test_oracle_vs_cv2 fuzzes against OpenCV)."""
    from stitchingvideo_b200 import capi
    rng = np.random.default_rng(77)
    for k in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 120))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        a[rng.random((h, w)) < 0.5] = 0
        mode = k % 4
        if mode == 0 and h % 2 == 0 and w % 2 == 0:
            ds = (w // 2, h // 2)
        elif mode == 1 and w % 2 == 0:
            ds = (w // 2, int(rng.integers(1, 200)))
        elif mode == 2:
            ds = (int(w * rng.uniform(1.0, 4.0)) + 1, int(h * rng.uniform(1.0, 4.0)) + 1)
        else:
            ds = (int(rng.integers(1, 200)), int(rng.integers(1, 150)))
        mw = (rng.random((ds[1], ds[0])) > 0.3).astype(np.uint8) * 255
        assert_same(capi.dilate3x3(a), O.dilate3x3(a), "dilate %dx%d" % (w, h))
        assert_same(capi.resize_linear_8u(a, ds), O.resize_linear_8u(a, ds), "resize %dx%d -> %dx%d" % (w, h, ds[0], ds[1]))
        assert_same(capi.refine_seam_mask(a, mw), O.refine_seam_mask(a, mw), "refine")
