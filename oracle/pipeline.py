"""CPU oracle of the per-frame loop of Stitcher::composePanorama (LIB/src/stitcher.cpp:221-313).

TEST INFRASTRUCTURE ONLY (see oracle/stitch_oracle.h): used by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker and the timed CPU arm.
"""
import numpy as np

from . import oracle as O

_BLEND = {"no": O.BLEND_NO, "feather": O.BLEND_FEATHER, "multiband": O.BLEND_MULTI_BAND}


class Calibration:
    """Per-sequence constants: warped corners/sizes, float maps, warped (seam-ANDed) masks."""

    def __init__(self, src_size, Ks, Rs, warper, scale, seam_masks=None):
        self.warper = O.Warper(warper, scale)
        self.Ks, self.Rs = Ks, Rs
        self.corners, self.sizes, self.maps, self.masks = [], [], [], []
        ones = np.full((src_size[1], src_size[0]), 255, np.uint8)
        for i, (K, R) in enumerate(zip(Ks, Rs)):
            roi, xmap, ymap = self.warper.build_maps(src_size, K, R)           # warpers_inl.hpp:62-85
            self.corners.append((roi[0], roi[1]))
            self.sizes.append((xmap.shape[1], xmap.shape[0]))
            self.maps.append((xmap, ymap))
            m = O.remap(ones, xmap, ymap, O.INTER_NEAREST, O.BORDER_CONSTANT)  # stitcher.cpp:278-280
            if seam_masks is not None:
                m = m & seam_masks[i]                                          # stitcher.cpp:294
            self.masks.append(m)


def compose(cal, frames, blender="multiband", num_bands=5, weight_type=O.CV_32F, sharpness=0.02, gains=None,
            output_8u=True, use_ref=False, gain_maps=None):
    """One frame set through warp -> gain -> convertTo(16S) -> feed -> blend -> convertTo(8U).
    use_ref: blend with the reference's own blenders.cpp (oracle/_ref) instead of the oracle's restatement."""
    if use_ref:
        from . import ref as RF
        b = RF.Blender(_BLEND[blender], num_bands, weight_type, sharpness)
    else:
        b = O.Blender(_BLEND[blender], num_bands, weight_type, sharpness)
    b.prepare(cal.corners, cal.sizes)                                          # stitcher.cpp:296-300
    for i, f in enumerate(frames):
        xmap, ymap = cal.maps[i]
        warped = O.remap(f, xmap, ymap, O.INTER_LINEAR, O.BORDER_REFLECT)      # stitcher.cpp:275
        if gains is not None:
            warped = O.gain_apply(warped, gains[i])                            # stitcher.cpp:283
        elif gain_maps is not None:
            warped = O.blocks_gain_apply(warped, gain_maps[i])                 # BlocksGainCompensator::apply, exposure_compensate.cpp:225-246
        b.feed(warped.astype(np.int16), cal.masks[i], cal.corners[i])          # stitcher.cpp:285, 303
    dst, dmask = b.blend()                                                     # stitcher.cpp:307
    if output_8u:
        dst = O.convert_16s_8u(dst)                                            # stitcher.cpp:313
    return dst, dmask
