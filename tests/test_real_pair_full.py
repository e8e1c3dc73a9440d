"""BASELINE.json configs[0] at its stated size: the reference's own 1920x1088 frames (REL32/output1/img-0.jpg, img-1.jpg), real
calibration, SphericalWarper + GainCompensator + MultiBandBlender(5 bands) - fixture tests/golden/real_pair_full.npz
(make_real_pair_full.py).  CPU: the oracle's panorama plus the stored sparse +-1 difference IS OpenCV's panorama (SHA-256).
GPU: every kernel variant of the compositor reproduces the oracle's panorama bit for bit."""
import hashlib
import io
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_pair_full.npz"))
N = int(G["n"])
SCALE = float(G["scale"])
KS, RS = [G["K%d" % i] for i in range(N)], [G["R%d" % i] for i in range(N)]
SIZE = (1920, 1088)
_cache = {}


def frames():
    if "imgs" not in _cache:
        Image = pytest.importorskip("PIL.Image")
        imgs = [np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(G["jpeg%d" % i].tobytes())).convert("RGB"))[:, :, ::-1]) for i in range(N)]
        for i, im in enumerate(imgs):
            if hashlib.sha256(im.tobytes()).digest() != G["img_sha256_%d" % i].tobytes():
                pytest.skip("this JPEG decoder gives other pixels than the one the fixture was made with")
        _cache["imgs"] = imgs
    return _cache["imgs"]


def oracle_panorama():
    if "pano" not in _cache:
        cal0 = P.Calibration(SIZE, KS, RS, "spherical", SCALE)
        seams = [O.resize_linear_8u(O.dilate3x3(G["seam_mask%d" % i]), cal0.sizes[i]) for i in range(N)]          # stitcher.cpp:291-292
        cal = P.Calibration(SIZE, KS, RS, "spherical", SCALE, seams)
        _cache["pano"] = (cal, seams, P.compose(cal, frames(), blender="multiband", num_bands=5, gains=list(G["gains"])))
    return _cache["pano"]


def test_oracle_plus_stored_difference_is_opencvs_panorama():
    cal, seams, (pano, mask) = oracle_panorama()
    assert [tuple(c) for c in cal.corners] == [tuple(c) for c in G["corners"]] and [tuple(s) for s in cal.sizes] == [tuple(s) for s in G["sizes"]]
    assert tuple(pano.shape) == tuple(G["pano_shape"])
    assert hashlib.sha256(mask.tobytes()).digest() == G["pano_mask_sha256"].tobytes()
    assert np.abs(G["diff_value"]).max() <= 1 and G["diff_index"].size < 1e-4 * pano.size      # +-1 LSB on < 0.01 % of the values
    cv = pano.astype(np.int16).reshape(-1)
    cv[G["diff_index"]] += G["diff_value"]
    assert hashlib.sha256(cv.astype(np.uint8).tobytes()).digest() == G["pano_sha256"].tobytes()


@pytest.mark.gpu
def test_cuda_path_on_the_full_size_real_pair(gpu):
    cal, seams, (opano, omask) = oracle_panorama()
    for fused in (11, 12, 14, 16, 17, 18, 10, 0):
        c = gpu.Compositor(SIZE, KS, RS, warper="spherical", scale=SCALE, blender="multiband", num_bands=5, gains=list(G["gains"]), seam_masks=seams)
        c.set_fused(fused)
        pano, mask = c.compose(frames())
        assert np.array_equal(pano, opano) and np.array_equal(mask, omask), "variant %d" % fused
