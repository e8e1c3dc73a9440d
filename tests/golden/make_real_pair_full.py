"""Full-size real-frame fixture for BASELINE.json configs[0]: the reference's own 1920x1088 sample frames
REL32/output1/img-0.jpg + img-1.jpg at their native size (make_real_pair.py halves them to keep its fixture small),
calibrated HERE with OpenCV (ORB, BestOf2Nearest, ray bundle adjustment, wave correction), seams from DpSeamFinder, gains
from GainCompensator, then OpenCV's own compose loop (stitcher.cpp:221-313 shape) with MultiBandBlender(5 bands).

To stay small the fixture holds the two JPEG files' bytes (decoded with PIL here AND in the test; a hash of the decoded
pixels guards against a different decoder), everything the per-frame path consumes, and of OpenCV's 2874x1100-odd panorama
only a SHA-256 plus the sparse, +-1 difference between it and the oracle's panorama at generation time: the test recomputes
the oracle's panorama, adds the stored difference and must land exactly on OpenCV's hash.
    python tests/golden/make_real_pair_full.py          (container with /root/reference and cv2)
"""
import hashlib
import io
import os
import sys

import cv2
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/stitching/OpenCV2.4.11-Stitching/Release/output1/"
cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)
cv2.setRNGSeed(12345)


def decode(jpeg_bytes):
    return np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(jpeg_bytes)).convert("RGB"))[:, :, ::-1])      # BGR like cv2.imread


def main():
    from oracle import oracle as O
    from oracle import pipeline as P
    raw = [open(SRC + "img-%d.jpg" % i, "rb").read() for i in (0, 1)]
    imgs = [decode(b) for b in raw]
    n = len(imgs)
    assert imgs[0].shape == (1088, 1920, 3)
    finder = cv2.ORB_create(4000)
    feats = [cv2.detail.computeImageFeatures2(finder, im) for im in imgs]
    matcher = cv2.detail_BestOf2NearestMatcher(False, 0.3)
    pairwise = matcher.apply2(feats)
    matcher.collectGarbage()
    ok, cams = cv2.detail_HomographyBasedEstimator().apply(feats, pairwise, None)
    assert ok
    for c in cams:
        c.R = c.R.astype(np.float32)
    ba = cv2.detail_BundleAdjusterRay()
    ba.setConfThresh(0.3)
    ok, cams = ba.apply(feats, pairwise, cams)
    assert ok
    rmats = cv2.detail.waveCorrect([np.copy(c.R) for c in cams], cv2.detail.WAVE_CORRECT_HORIZ)
    scale = float(np.median([c.focal for c in cams]))
    Ks = [np.array([[c.focal, 0, c.ppx], [0, c.focal * c.aspect, c.ppy], [0, 0, 1]], np.float32) for c in cams]
    Rs = [np.asarray(r, np.float32) for r in rmats]
    swa = 0.25                                                # seam_work_aspect (stitcher.cpp:165-177)
    small = [cv2.resize(im, None, fx=swa, fy=swa, interpolation=cv2.INTER_LINEAR) for im in imgs]
    wseam = cv2.PyRotationWarper("spherical", scale * swa)
    corners_s, warped_s, masks_s = [], [], []
    for im, K, R in zip(small, Ks, Rs):
        Ks_ = K.copy()
        Ks_[0, 0] *= swa; Ks_[0, 2] *= swa; Ks_[1, 1] *= swa; Ks_[1, 2] *= swa
        c, w = wseam.warp(im, Ks_, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = wseam.warp(np.full(im.shape[:2], 255, np.uint8), Ks_, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners_s.append(c); warped_s.append(w); masks_s.append(m)
    comp = cv2.detail_GainCompensator(1)
    comp.feed(corners_s, [cv2.UMat(w) for w in warped_s], [cv2.UMat(m) for m in masks_s])
    gains = np.array(comp.getMatGains(), np.float64).reshape(-1)
    seam = cv2.detail_DpSeamFinder("COLOR")
    seam_masks = [m.get() for m in seam.find([w.astype(np.float32) for w in warped_s], corners_s, [cv2.UMat(m) for m in masks_s])]
    warper = cv2.PyRotationWarper("spherical", scale)
    blender = cv2.detail_MultiBandBlender(0, 5, cv2.CV_32F)
    corners, sizes, feeds = [], [], []
    for i, (im, K, R) in enumerate(zip(imgs, Ks, Rs)):
        c, w = warper.warp(im, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, mw = warper.warp(np.full(im.shape[:2], 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        w = comp.apply(i, c, w, mw)
        dil = cv2.dilate(seam_masks[i], None)
        sm = cv2.resize(dil, (mw.shape[1], mw.shape[0]), interpolation=cv2.INTER_LINEAR)
        corners.append(c); sizes.append((w.shape[1], w.shape[0])); feeds.append((w.astype(np.int16), sm & mw))
    blender.prepare(cv2.detail.resultRoi(corners=corners, sizes=sizes))
    for (w, m), c in zip(feeds, corners):
        blender.feed(w, m, c)
    res, res_mask = blender.blend(None, None)
    pano = np.clip(res, 0, 255).astype(np.uint8)
    # the oracle's panorama now, and where it differs from OpenCV's (float pyrDown summation order: +-1 LSB)
    size = (1920, 1088)
    cal0 = P.Calibration(size, Ks, Rs, "spherical", scale)
    seams = [O.resize_linear_8u(O.dilate3x3(seam_masks[i]), cal0.sizes[i]) for i in range(n)]
    cal = P.Calibration(size, Ks, Rs, "spherical", scale, seams)
    opano, omask = P.compose(cal, imgs, blender="multiband", num_bands=5, gains=list(gains))
    assert opano.shape == pano.shape and np.array_equal(omask, res_mask)
    diff = pano.astype(np.int16) - opano.astype(np.int16)
    assert np.abs(diff).max() <= 1
    idx = np.flatnonzero(diff)
    out = {"n": np.int32(n), "scale": np.float32(scale), "seam_aspect": np.float32(swa), "gains": gains,
           "pano_shape": np.array(pano.shape, np.int32), "pano_sha256": np.frombuffer(hashlib.sha256(pano.tobytes()).digest(), np.uint8),
           "diff_index": idx.astype(np.int64), "diff_value": diff.reshape(-1)[idx].astype(np.int8),
           "pano_mask_sha256": np.frombuffer(hashlib.sha256(res_mask.tobytes()).digest(), np.uint8),
           "corners": np.array(corners, np.int32), "sizes": np.array(sizes, np.int32)}
    for i in range(n):
        out.update({"jpeg%d" % i: np.frombuffer(raw[i], np.uint8), "img_sha256_%d" % i: np.frombuffer(hashlib.sha256(imgs[i].tobytes()).digest(), np.uint8),
                    "K%d" % i: Ks[i], "R%d" % i: Rs[i], "seam_mask%d" % i: seam_masks[i]})
    path = os.path.join(HERE, "real_pair_full.npz")
    np.savez_compressed(path, **out)
    print("scale", scale, "gains", gains, "pano", pano.shape, "differing values", idx.size, "of", diff.size, "file MB", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    main()
