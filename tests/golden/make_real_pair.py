"""Real-frame fixture for BASELINE.json configs[0] (2-camera 1080p pair, SphericalWarper + MultiBandBlender(5 bands)):
the reference's own sample frames REL32/output1/img-0.jpg + img-1.jpg (neighbours per REL32/test.txt; configs[0]'s
"ruandata/TestRelease" holds no images), halved to 960x544 to keep the fixture small, calibrated HERE
with OpenCV (cv2 4.13: ORB features, BestOf2Nearest matcher, homography estimator, ray bundle adjustment — the host-side
calibration that is out of this path's scope), then run through OpenCV's own cv::detail compose loop
(stitcher.cpp:221-313 shape: seam-scale warp -> GainCompensator::feed -> seam finder -> compose-scale warp -> apply ->
dilate/resize/& -> MultiBandBlender).  Everything the per-frame path consumes and what OpenCV produced from it is stored
in tests/golden/real_pair.npz.  Run from the repo root in the container that has /root/reference:
    python tests/golden/make_real_pair.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/stitching/OpenCV2.4.11-Stitching/Release/output1/"
cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)
cv2.setRNGSeed(12345)


def main():
    imgs = [cv2.resize(cv2.imread(SRC + "img-%d.jpg" % i), (960, 544), interpolation=cv2.INTER_AREA) for i in (0, 1)]
    n = len(imgs)
    finder = cv2.ORB_create(2000)
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

    # ---- seam-estimation scale (stitcher.cpp:165-219 shape): warp, exposure feed, seams
    swa = 0.5
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
    seam_masks = seam.find([w.astype(np.float32) for w in warped_s], corners_s, [cv2.UMat(m) for m in masks_s])
    seam_masks = [m.get() for m in seam_masks]

    # ---- compose scale (stitcher.cpp:221-313)
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
    out = {"n": np.int32(n), "scale": np.float32(scale), "seam_aspect": np.float32(swa), "gains": gains,
           "pano": pano, "pano_mask": res_mask, "corners": np.array(corners, np.int32), "sizes": np.array(sizes, np.int32),
           "corners_seam": np.array(corners_s, np.int32)}
    for i in range(n):
        out.update({"img%d" % i: imgs[i], "small%d" % i: small[i], "K%d" % i: Ks[i], "R%d" % i: Rs[i], "seam_mask%d" % i: seam_masks[i],
                    "warped_mask_seam%d" % i: masks_s[i]})
    np.savez_compressed(os.path.join(HERE, "real_pair.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})
    print("gains", gains, "scale", scale, "pano", pano.shape, "file MB", os.path.getsize(os.path.join(HERE, "real_pair.npz")) / 1e6)


if __name__ == "__main__":
    main()
