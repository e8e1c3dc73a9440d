"""CPU: the oracle (oracle/, the C restatement of the reference path) against the committed golden
vectors under tests/golden/ (generated from OpenCV by tests/golden/make_golden.py; the reference
itself ships no golden vectors for this path).  Bit-exact everywhere except the
float-weight multi-band case, where north_star allows +-1 LSB (float pyrDown is build dependent,
SURVEY.md §7)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def same(got, ref, what):
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    if got.dtype == np.float32:
        d = got.view(np.uint32) != np.ascontiguousarray(ref, np.float32).view(np.uint32)
    else:
        d = got != ref
    assert not d.any(), "%s: %d of %d values differ" % (what, int(d.sum()), d.size)


@pytest.mark.parametrize("kind", ["spherical", "cylindrical", "plane"])
def test_build_maps_golden(kind):
    g = load("maps")
    K, R, scale = g[kind + "_K"], g[kind + "_R"], float(g[kind + "_scale"])
    W, H = (int(v) for v in g[kind + "_size"])
    w = O.Warper(kind, scale)
    roi, xm, ym = w.build_maps((W, H), K, R)
    assert tuple(roi) == tuple(int(v) for v in g[kind + "_roi"])
    same(xm, g[kind + "_xmap"], kind + " xmap")
    same(ym, g[kind + "_ymap"], kind + " ymap")
    assert w.warp_roi((W, H), K, R) == tuple(int(v) for v in g[kind + "_warproi"])
    u, v = w.warp_point((W / 3.0, H / 5.0), K, R)
    same(np.array([u, v], np.float32), g[kind + "_pt"], kind + " warpPoint")


@pytest.mark.parametrize("cn", [1, 3])
def test_remap_golden(cn):
    g = load("remap")
    src, xm, ym = g["src%d" % cn], g["xmap%d" % cn], g["ymap%d" % cn]
    for border in range(5):
        for interp in (0, 1):
            got = O.remap(src, xm, ym, interp, border, (7, 9, 11, 0))
            same(got, g["dst%d_b%d_i%d" % (cn, border, interp)], "remap cn=%d border=%d interp=%d" % (cn, border, interp))


@pytest.mark.parametrize("cn", [1, 3])
def test_remap_fixed_point_maps_golden(cn):
    """cv::convertMaps + cv::remap with the CV_16SC2 / CV_16UC1 pair."""
    g = load("remap_fixed")
    src, xm, ym = g["src%d" % cn], g["xmap%d" % cn], g["ymap%d" % cn]
    m1, m2 = O.convert_maps(xm, ym)
    same(m1, g["map1_%d" % cn], "convertMaps map1")
    same(m2, g["map2_%d" % cn], "convertMaps map2")
    n1, _ = O.convert_maps(xm, ym, nn_interpolation=True)
    same(n1, g["nnmap1_%d" % cn], "convertMaps nearest map1")
    for border in range(5):
        for interp in (0, 1):
            same(O.remap_fixed(src, m1, m2, interp, border, (7, 9, 11, 0)), g["dst%d_b%d_i%d" % (cn, border, interp)],
                 "fixed remap cn=%d border=%d interp=%d" % (cn, border, interp))
        same(O.remap_fixed(src, n1, None, O.INTER_NEAREST, border, (7, 9, 11, 0)), g["nndst%d_b%d" % (cn, border)], "fixed nearest")


def test_pyramids_golden():
    g = load("pyr")
    for name in ("s16", "u8", "s16c1"):
        a = g[name]
        same(O.pyr_down(a), g[name + "_down"], name + " pyrDown")
        same(O.pyr_up(a), g[name + "_up"], name + " pyrUp")


def test_gain_resize_distance_golden():
    g = load("misc")
    img = g["img"]
    for i in range(4):
        same(O.gain_apply(img, float(g["gain%d" % i])), g["gain%d_out" % i], "gain %d" % i)
    same(O.resize_linear(g["gmap"], (40, 30)), g["gmap_resized"], "resize(INTER_LINEAR)")
    same(O.distance_l1(g["mask"]), g["dist"], "distanceTransform(L1, 3)")


@pytest.mark.parametrize("case,kind,kw,tol", [
    ("no", O.BLEND_NO, {}, 0),
    ("feather", O.BLEND_FEATHER, {"sharpness": 0.02}, 0),
    ("mb16_5", O.BLEND_MULTI_BAND, {"num_bands": 5, "weight_type": O.CV_16S}, 0),
    ("mb16_2", O.BLEND_MULTI_BAND, {"num_bands": 2, "weight_type": O.CV_16S}, 0),
    ("mb32_5", O.BLEND_MULTI_BAND, {"num_bands": 5, "weight_type": O.CV_32F}, 1),   # +-1 LSB: float-weight normalize
])
def test_blenders_golden(case, kind, kw, tol):
    g = load("blend")
    n = int(g["n"])
    imgs = [g["img%d" % i] for i in range(n)]
    masks = [g["mask%d" % i] for i in range(n)]
    tls = [tuple(int(v) for v in t) for t in g["tls"]]
    b = O.Blender(kind, **kw)
    b.prepare(tls, [(im.shape[1], im.shape[0]) for im in imgs])
    for im, m, tl in zip(imgs, masks, tls):
        b.feed(im, m, tl)
    dst, dmask = b.blend()
    ref, rmask = g[case + "_dst"], g[case + "_mask"]
    assert dst.shape == ref.shape
    if tol == 0:
        same(dst, ref, case + " dst")
        same(dmask, rmask, case + " mask")
    else:
        d = np.abs(dst.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= tol, "%s: max |diff| %d > %d" % (case, int(d.max()), tol)
        same(dmask, rmask, case + " mask")


def test_scalar_semantics():
    """x86 cvtss2si / cvttss2si behaviour the path depends on."""
    L = O.lib()
    assert [L.so_cvround(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5)] == [0, 2, 2, 0, -2]     # half to even
    assert L.so_cvround(3e9) == -2 ** 31 and L.so_cvround(float("nan")) == -2 ** 31        # integer indefinite
    assert L.so_trunc_short(-1.9) == -1 and L.so_trunc_short(1.9) == 1                     # toward zero
    assert L.so_trunc_short(65537.5) == 1                                                  # low 16 bits
    assert L.so_trunc_short(1e20) == 0                                                     # 0x80000000 -> low word 0
    for border, want in ((O.BORDER_REFLECT, [1, 0, 0, 4, 4, 3]), (O.BORDER_REFLECT_101, [2, 1, 0, 4, 3, 2]),
                         (O.BORDER_REPLICATE, [0, 0, 0, 4, 4, 4]), (O.BORDER_WRAP, [3, 4, 0, 4, 0, 1]),
                         (O.BORDER_CONSTANT, [-1, -1, 0, 4, -1, -1])):
        assert [L.so_border_interpolate(p, 5, border) for p in (-2, -1, 0, 4, 5, 6)] == want


def test_sinf_cosf_match_host_libm():
    """The oracle's portable sinf/cosf equal the host libm's bit for bit (the functions OpenCV's
    mapBackward calls, warpers_inl.hpp:256-259,289-291) on a dense sample incl. large arguments."""
    import ctypes as C
    import ctypes.util
    m = C.CDLL(ctypes.util.find_library("m"))
    m.sinf.restype = m.cosf.restype = C.c_float
    m.sinf.argtypes = m.cosf.argtypes = [C.c_float]
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-8, 8, 20000), rng.uniform(-200, 200, 5000), rng.uniform(-1e6, 1e6, 2000),
                         [0.0, -0.0, 1e-20, 1e-5, np.pi / 4, np.pi / 2, np.pi, 1e10, -1e30]]).astype(np.float32)
    s, c = O.sinf(xs), O.cosf(xs)
    hs = np.array([m.sinf(float(v)) for v in xs], np.float32)
    hc = np.array([m.cosf(float(v)) for v in xs], np.float32)
    assert (s.view(np.uint32) == hs.view(np.uint32)).all()
    assert (c.view(np.uint32) == hc.view(np.uint32)).all()


def test_calibration_side_steps_golden():
    """ExposureCompensator::feed (exposure_compensate.cpp:76-147, 165-222) and the seam-mask refinement
    (stitcher.cpp:291-294) against cv2: the integer steps bit for bit, block gain maps bit for bit, scalar gains to
    1e-12 (cv::solve takes a closed form for n <= 3)."""
    from tests import util
    g = load("calib")
    for k, ds in enumerate(((106, 74), (32, 20), (5, 100), (70, 50), (100, 77), (20, 3))):
        a = g["m%d" % k]
        same(O.dilate3x3(a), g["m%d_dilate" % k], "dilate %d" % k)
        same(O.resize_linear_8u(a, ds), g["m%d_resize" % k], "resize %d" % k)
        same(O.refine_seam_mask(a, g["m%d_warped" % k]), g["m%d_refined" % k], "refine %d" % k)
    same(O.sep_filter3(g["gmap"]), g["gmap_smooth"], "sepFilter2D")
    for n, w, h in ((2, 120, 90), (3, 160, 100), (5, 200, 120)):
        corners, imgs, masks = util.exposure_scene(n, w, h, seed=n)
        gains = O.gain_feed(corners, imgs, masks)
        np.testing.assert_allclose(gains, g["gains%d" % n], rtol=1e-12, atol=0)
        for i, m in enumerate(O.blocks_gain_feed(corners, imgs, masks)):
            same(m, g["blocks%d_%d" % (n, i)], "block gain map %d/%d" % (i, n))
