"""stitchingvideo_b200 — B200-native (sm_100a CUDA) per-frame compositing path of StitchingVideo's
OpenCV 2.4.11 cv::detail pipeline: warp -> exposure gain -> blend, behind the reference's own
RotationWarper / ExposureCompensator / Blender interfaces (C ABI: include/stitchb200.h).

The package holds only what that path needs: csrc/ (CUDA kernels + C ABI), capi.py (ctypes binding
and the host-side mirror of the reference interfaces), rigs.py (deterministic synthetic camera rigs
of SURVEY.md §8d), sharding.py (frame sharding across GPUs).  No CPU fallback exists.
"""
from .capi import (  # noqa: F401
    BLEND_FEATHER, BLEND_MULTI_BAND, BLEND_NO, BORDER_CONSTANT, BORDER_REFLECT, BORDER_REFLECT_101,
    BORDER_REPLICATE, BORDER_WRAP, COMP_GAIN, COMP_GAIN_BLOCKS, COMP_NO, CV_8U, CV_8UC1, CV_8UC3, CV_16S,
    CV_16SC1, CV_16SC3, CV_32F, CV_32FC1, INTER_LINEAR, INTER_NEAREST, Blender, BlocksGainCompensator,
    Batch, Compositor, CylindricalWarper, DeviceImage, MultiCompositor, ExposureCompensator, FeatherBlender, GainCompensator,
    MultiBandBlender, NoExposureCompensator, PlaneWarper, RotationWarper, SphericalWarper, StitchError,
    CompressedRectilinearPortraitWarper, CompressedRectilinearWarper, CylindricalPortraitWarper, FisheyeWarper, MercatorWarper,
    PaniniPortraitWarper, PaniniWarper, PlanePortraitWarper, SphericalPortraitWarper, StereographicWarper, TransverseMercatorWarper,
    convertMaps, createLaplacePyr, createWeightMap, device_count, kernel_launch_count, lib, normalizeUsingWeightMap, remap,
    restoreImageFromLaplacePyr,
)

__all__ = [n for n in dir() if not n.startswith("_")]
