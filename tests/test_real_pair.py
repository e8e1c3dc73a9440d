"""BASELINE.json configs[0] on REAL frames: the reference's own sample images (REL32/output1/img-0.jpg, img-1.jpg), a real
calibration (rotations with pitch and roll, not the synthetic pure-yaw rigs), seam masks from a seam finder, gains from the
exposure compensator — fixture tests/golden/real_pair.npz, made by tests/golden/make_real_pair.py with OpenCV's own
cv::detail compose loop (stitcher.cpp:221-313 shape).

CPU: the oracle reproduces OpenCV's panorama (+-1 LSB, the float-weight tolerance of north_star; masks equal; seam-scale
geometry, gains and refined masks exact).  GPU: the CUDA path reproduces the oracle bit for bit, end to end: seam-scale
warp -> GainCompensator::feed -> dilate/resize -> compositor."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_pair.npz"))
N = int(G["n"])
SCALE, SWA = float(G["scale"]), float(G["seam_aspect"])
KS, RS = [G["K%d" % i] for i in range(N)], [G["R%d" % i] for i in range(N)]
IMGS, SMALL = [G["img%d" % i] for i in range(N)], [G["small%d" % i] for i in range(N)]
SIZE = (IMGS[0].shape[1], IMGS[0].shape[0])


def k_seam(K):
    Ks = K.copy()                                            # stitcher.cpp:173-177: K scaled by seam_work_aspect
    Ks[0, 0] *= np.float32(SWA); Ks[0, 2] *= np.float32(SWA); Ks[1, 1] *= np.float32(SWA); Ks[1, 2] *= np.float32(SWA)
    return Ks


def oracle_seam_stage():
    w = O.Warper("spherical", SCALE * SWA)
    corners, warped, masks = [], [], []
    for im, K, R in zip(SMALL, KS, RS):
        c, d = w.warp(im, k_seam(K), R, O.INTER_LINEAR, O.BORDER_REFLECT)
        _, m = w.warp(np.full(im.shape[:2], 255, np.uint8), k_seam(K), R, O.INTER_NEAREST, O.BORDER_CONSTANT)
        corners.append(c); warped.append(d); masks.append(m)
    return corners, warped, masks


def oracle_panorama():
    cal0 = P.Calibration(SIZE, KS, RS, "spherical", SCALE)
    seams = [O.resize_linear_8u(O.dilate3x3(G["seam_mask%d" % i]), cal0.sizes[i]) for i in range(N)]     # stitcher.cpp:291-292
    cal = P.Calibration(SIZE, KS, RS, "spherical", SCALE, seams)
    return cal, seams, P.compose(cal, IMGS, blender="multiband", num_bands=5, gains=list(G["gains"]))


def test_oracle_reproduces_opencv_on_real_frames():
    corners, warped, masks = oracle_seam_stage()
    assert [tuple(c) for c in corners] == [tuple(c) for c in G["corners_seam"]]
    for i in range(N):
        assert np.array_equal(masks[i], G["warped_mask_seam%d" % i])
    np.testing.assert_allclose(O.gain_feed(corners, warped, masks), G["gains"], rtol=1e-12, atol=0)      # n = 2: cv::solve closed form
    cal, seams, (pano, mask) = oracle_panorama()
    assert [tuple(c) for c in cal.corners] == [tuple(c) for c in G["corners"]]
    assert [tuple(s) for s in cal.sizes] == [tuple(s) for s in G["sizes"]]
    assert pano.shape == G["pano"].shape and np.array_equal(mask, G["pano_mask"])
    d = np.abs(pano.astype(int) - G["pano"].astype(int))
    assert d.max() <= 1, "oracle vs OpenCV: max |diff| %d" % d.max()          # float pyrDown summation order
    assert (d != 0).mean() < 0.02


@pytest.mark.gpu
def test_cuda_path_on_real_frames(gpu):
    from stitchingvideo_b200 import capi
    # seam-estimation scale: warp + exposure compensator feed on the device
    w = gpu.SphericalWarper(SCALE * SWA)
    ocorners, owarped, omasks = oracle_seam_stage()
    corners, warped, masks = [], [], []
    for im, K, R in zip(SMALL, KS, RS):
        c, d = w.warp(im, k_seam(K), R, gpu.INTER_LINEAR, gpu.BORDER_REFLECT)
        _, m = w.warp(np.full(im.shape[:2], 255, np.uint8), k_seam(K), R, gpu.INTER_NEAREST, gpu.BORDER_CONSTANT)
        corners.append(tuple(c)); warped.append(d); masks.append(m)
    assert corners == [tuple(c) for c in ocorners]
    for i in range(N):
        assert np.array_equal(warped[i], owarped[i]) and np.array_equal(masks[i], omasks[i])
    comp = gpu.GainCompensator()
    comp.feed(corners, warped, masks)
    np.testing.assert_allclose(comp.gains(), G["gains"], rtol=1e-11, atol=0)
    # compose scale: refined seam masks, then the fused frame loop
    cal, oseams, (opano, omask) = oracle_panorama()
    seams = [capi.resize_linear_8u(capi.dilate3x3(G["seam_mask%d" % i]), cal.sizes[i]) for i in range(N)]
    for a, b in zip(seams, oseams):
        assert np.array_equal(a, b)
    for fused in (11, 12, 14, 16, 17, 18, 10, 0):
        c = gpu.Compositor(SIZE, KS, RS, warper="spherical", scale=SCALE, blender="multiband", num_bands=5,
                           gains=list(G["gains"]), seam_masks=seams)
        c.set_fused(fused)
        pano, mask = c.compose(IMGS)
        assert np.array_equal(pano, opano) and np.array_equal(mask, omask), "variant %d" % fused
    assert np.abs(pano.astype(int) - G["pano"].astype(int)).max() <= 1 and np.array_equal(mask, G["pano_mask"])
