"""CPU: the C-ABI shared library loads and exports every symbol include/stitchb200.h declares, the
ctypes table in capi.py covers exactly that set, and — with no GPU in this container — every entry
point that needs the device fails loudly with CV_GpuApiCallError instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stitchb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    names = declared_symbols()
    for must in ("sb_warper_build_maps", "sb_warper_warp", "sb_warper_warp_roi", "sb_warper_warp_point", "sb_warper_remap",
                 "sb_comp_apply", "sb_comp_set_gains", "sb_blender_prepare", "sb_blender_prepare_rect", "sb_blender_feed",
                 "sb_blender_blend", "sb_compositor_compose", "sb_last_error"):
        assert must in names
    assert len(names) >= 50


def test_library_exports_every_declared_symbol():
    from stitchingvideo_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "libstitchb200.so missing: run __graft_entry__.build()"
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, "declared in include/stitchb200.h but not exported: %s" % missing
    assert sorted(capi.API) == declared_symbols(), "capi.API and the header disagree"


def test_no_torch_types_in_signatures():
    text = open(HEADER).read()
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_version_and_error_string():
    import stitchingvideo_b200 as sv
    L = sv.lib()
    assert b"sm_100a" in L.sb_version()
    assert isinstance(L.sb_last_error(), bytes)
    assert L.sb_kernel_launch_count() >= 0


def _has_gpu():
    import stitchingvideo_b200 as sv
    return sv.device_count() > 0


@pytest.mark.skipif(_has_gpu(), reason="a CUDA device is visible: the no-device behaviour cannot be observed")
def test_no_cpu_fallback_without_a_device():
    import stitchingvideo_b200 as sv
    K = np.array([[90, 0, 50], [0, 90, 40], [0, 0, 1]], np.float32)
    R = np.eye(3, dtype=np.float32)
    for make in (lambda: sv.SphericalWarper(100.0), lambda: sv.MultiBandBlender(), lambda: sv.FeatherBlender(),
                 lambda: sv.GainCompensator(),
                 lambda: sv.Compositor((64, 48), [K], [R], warper="spherical", scale=90.0, blender="feather"),
                 lambda: sv.remap(np.zeros((4, 4, 3), np.uint8), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32))):
        with pytest.raises(sv.StitchError) as e:
            make()
        assert e.value.code == -217, "expected CV_GpuApiCallError, got %d" % e.value.code
        assert "no CPU fallback" in str(e.value)


def test_argument_errors_mirror_the_reference():
    """Factory errors are raised before any device work (blenders.cpp:60, exposure_compensate.cpp:58, blenders.cpp:198)."""
    import stitchingvideo_b200 as sv
    with pytest.raises(sv.StitchError) as e:
        sv.Blender.createDefault(7)
    assert e.value.code == -5
    with pytest.raises(sv.StitchError) as e:
        sv.ExposureCompensator.createDefault(9)
    assert e.value.code == -5
    with pytest.raises(sv.StitchError) as e:
        sv.MultiBandBlender(False, 5, sv.CV_8U)
    assert e.value.code == -215
