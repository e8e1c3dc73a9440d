"""CPU oracle of the per-frame loop of Stitcher::composePanorama (LIB/src/stitcher.cpp:221-313).

TEST INFRASTRUCTURE ONLY (see oracle/stitch_oracle.h): used by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker and the timed CPU arm.
"""
import numpy as np

from . import oracle as O

_BLEND = {"no": O.BLEND_NO, "feather": O.BLEND_FEATHER, "multiband": O.BLEND_MULTI_BAND}


class Calibration:
    """Per-sequence constants: warped corners/sizes, float maps, warped (seam-ANDed) masks."""

    def __init__(self, src_size, Ks, Rs, warper, scale, seam_masks=None, build_maps=None):
        """build_maps: optional callable (src_size, K, R) -> (roi, xmap, ymap) standing in for the warper - the reference's
        own RotationWarperBase<P>::buildMaps from oracle/_ref for the projectors the C restatement does not cover."""
        self.warper = O.Warper(warper, scale) if build_maps is None else None
        self.Ks, self.Rs = Ks, Rs
        self.corners, self.sizes, self.maps, self.masks = [], [], [], []
        ones = np.full((src_size[1], src_size[0]), 255, np.uint8)
        for i, (K, R) in enumerate(zip(Ks, Rs)):
            roi, xmap, ymap = (build_maps or self.warper.build_maps)(src_size, K, R)           # warpers_inl.hpp:62-85
            self.corners.append((roi[0], roi[1]))
            self.sizes.append((xmap.shape[1], xmap.shape[0]))
            self.maps.append((xmap, ymap))
            m = O.remap(ones, xmap, ymap, O.INTER_NEAREST, O.BORDER_CONSTANT)  # stitcher.cpp:278-280
            if seam_masks is not None:
                m = m & seam_masks[i]                                          # stitcher.cpp:294
            self.masks.append(m)


def warp_frame(cal, i, f, gains=None, gain_maps=None, undistort_maps=None):
    """One camera's frame as the blender sees it: [undistort ->] warp -> exposure gain (8UC3)."""
    if undistort_maps is not None:                                             # APP64:741 remap(iimg, img3, mapEye1, mapEye2, INTER_LINEAR)
        f = O.remap_fixed(f, undistort_maps[i][0], undistort_maps[i][1], O.INTER_LINEAR, O.BORDER_CONSTANT)
    xmap, ymap = cal.maps[i]
    warped = O.remap(f, xmap, ymap, O.INTER_LINEAR, O.BORDER_REFLECT)          # stitcher.cpp:275, APP64:752
    if gains is not None:
        warped = O.gain_apply(warped, gains[i])                                # stitcher.cpp:283
    elif gain_maps is not None:
        warped = O.blocks_gain_apply(warped, gain_maps[i])                     # BlocksGainCompensator::apply, exposure_compensate.cpp:225-246; APP64:310-331
    return warped


def crop_geometry(pano_w, pano_h, up, down, left, right):
    """UpdateMat / feedSizeRemap (APP64:702, 153): size of ResultStitch and the row / column offsets of its gather, in the
    float arithmetic of the reference (upblack etc. are floats; conversions to int truncate)."""
    f32 = np.float32
    keep = f32(f32(1) - f32(up)) - f32(down)
    out_w = int(f32(f32(pano_w) - f32(left)) - f32(right))
    out_h = int(f32(pano_h) * keep)
    yy = int(f32(f32(out_h) / keep) * f32(up))
    return out_w, out_h, int(left), yy


def compose_app(cal, frames, gains=None, gain_maps=None, undistort_maps=None, crop=(0.0, 0.0, 0, 0), fill=True):
    """The live app's per-frame composite (APP64:724-770): look-up tables built once by feedSize (APP64:115-148: the LAST
    camera whose mask is non-zero owns a panorama pixel; tables stay zero elsewhere), then per frame feedSizeRemap's
    unconditional gather (APP64:150-177) into the cropped ResultStitch.  -> (8UC3 composite, full-size mask)."""
    roi = O.result_roi(cal.corners, cal.sizes)
    W, H = roi[2], roi[3]
    idx = np.zeros((H, W), np.int64); ly = np.zeros((H, W), np.int64); lx = np.zeros((H, W), np.int64)
    mask = np.zeros((H, W), np.uint8)
    for i, m in enumerate(cal.masks):                                          # feedSize, camera by camera
        dx, dy = cal.corners[i][0] - roi[0], cal.corners[i][1] - roi[1]
        h, w = m.shape
        ys, xs = np.nonzero(m)
        idx[dy + ys, dx + xs] = i; ly[dy + ys, dx + xs] = ys; lx[dy + ys, dx + xs] = xs
        mask[dy:dy + h, dx:dx + w] |= m
    warped = [warp_frame(cal, i, f, gains, gain_maps, undistort_maps).astype(np.int16) for i, f in enumerate(frames)]
    out_w, out_h, xx, yy = crop_geometry(W, H, *crop)
    out = np.zeros((out_h, out_w, 3), np.int16)
    sub = (slice(yy, yy + out_h), slice(xx, xx + out_w))
    for i, wimg in enumerate(warped):                                          # feedSizeRemap: dst = img[idx].at(y, x), no mask test
        sel = idx[sub] == i
        out[sel] = wimg[ly[sub][sel], lx[sub][sel]]
    if not fill:                                                               # Blender::blend semantics instead (blenders.cpp:105-112)
        out[mask[sub] == 0] = 0
    return O.convert_16s_8u(out), mask


def compose(cal, frames, blender="multiband", num_bands=5, weight_type=O.CV_32F, sharpness=0.02, gains=None,
            output_8u=True, use_ref=False, gain_maps=None, undistort_maps=None):
    """One frame set through [undistort ->] warp -> gain -> convertTo(16S) -> feed -> blend -> convertTo(8U).
    use_ref: blend with the reference's own blenders.cpp (oracle/_ref) instead of the oracle's restatement."""
    if use_ref:
        from . import ref as RF
        b = RF.Blender(_BLEND[blender], num_bands, weight_type, sharpness)
    else:
        b = O.Blender(_BLEND[blender], num_bands, weight_type, sharpness)
    b.prepare(cal.corners, cal.sizes)                                          # stitcher.cpp:296-300
    for i, f in enumerate(frames):
        warped = warp_frame(cal, i, f, gains, gain_maps, undistort_maps)
        b.feed(warped.astype(np.int16), cal.masks[i], cal.corners[i])          # stitcher.cpp:285, 303
    dst, dmask = b.blend()                                                     # stitcher.cpp:307
    if output_8u:
        dst = O.convert_16s_8u(dst)                                            # stitcher.cpp:313
    return dst, dmask
