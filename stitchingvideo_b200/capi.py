"""ctypes binding of libstitchb200.so (include/stitchb200.h) and the host-side mirror of the
reference's operator interfaces for the per-frame compositing path:

    detail::RotationWarper       (warpers.hpp:53-72)            -> SphericalWarper / CylindricalWarper / PlaneWarper
    detail::ExposureCompensator  (exposure_compensate.hpp:51-101) -> NoExposureCompensator / GainCompensator / BlocksGainCompensator
    detail::Blender              (blenders.hpp:53-117)           -> Blender / FeatherBlender / MultiBandBlender
    Stitcher::composePanorama's frame loop (stitcher.cpp:221-313) -> Compositor

Same names, argument meaning and error behaviour (cv::Exception -> StitchError carrying the same
status code).  There is NO CPU fallback: if the shared library is missing or no CUDA device is
usable, every operation raises.  numpy arrays stand in for cv::Mat (host), DeviceImage for GpuMat.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# STITCHB200_LIB: an alternative build of the same library (kernel-shape experiments); never a fallback
LIB_PATH = os.environ.get("STITCHB200_LIB") or os.path.join(_HERE, "libstitchb200.so")

# OpenCV-numbered constants
CV_8U, CV_16S, CV_32F = 0, 3, 5
CV_8UC1, CV_8UC3, CV_16SC1, CV_16SC3, CV_32FC1 = 0, 16, 3, 19, 5
INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4
WARP_PLANE, WARP_CYLINDRICAL, WARP_SPHERICAL = 0, 1, 2
(WARP_FISHEYE, WARP_STEREOGRAPHIC, WARP_COMPRESSED_RECTILINEAR, WARP_COMPRESSED_RECTILINEAR_PORTRAIT, WARP_PANINI, WARP_PANINI_PORTRAIT,
 WARP_MERCATOR, WARP_TRANSVERSE_MERCATOR, WARP_SPHERICAL_PORTRAIT, WARP_CYLINDRICAL_PORTRAIT, WARP_PLANE_PORTRAIT) = range(3, 14)
COMP_NO, COMP_GAIN, COMP_GAIN_BLOCKS = 0, 1, 2
BLEND_NO, BLEND_FEATHER, BLEND_MULTI_BAND = 0, 1, 2
SB_OK, SB_ERR_NO_MEM, SB_ERR_BAD_ARG, SB_ERR_ASSERT, SB_ERR_NOT_IMPL, SB_ERR_CUDA = 0, -4, -5, -215, -213, -217


class StitchError(RuntimeError):
    """cv::Exception stand-in: .code is the OpenCV status code the reference would throw."""

    def __init__(self, code, msg):
        super().__init__("stitchb200 error %d: %s" % (code, msg))
        self.code = code


class SbImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("type", C.c_int),
                ("step", C.c_size_t), ("device", C.c_int)]


class SbPoint(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class SbSize(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int)]


class SbRect(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("width", C.c_int), ("height", C.c_int)]


class SbCompositorConfig(C.Structure):
    _fields_ = [("n_cameras", C.c_int), ("src_size", SbSize), ("warper_kind", C.c_int), ("warper_scale", C.c_float),
                ("K", C.POINTER(C.c_float)), ("R", C.POINTER(C.c_float)), ("blender_kind", C.c_int),
                ("num_bands", C.c_int), ("weight_type", C.c_int), ("sharpness", C.c_float), ("comp_kind", C.c_int),
                ("gains", C.POINTER(C.c_double)), ("seam_masks", C.POINTER(SbImage)), ("output_type", C.c_int),
                ("gain_maps", C.POINTER(SbImage)),
                ("warper_a", C.c_float), ("warper_b", C.c_float),
                ("undistort_map1", C.POINTER(SbImage)), ("undistort_map2", C.POINTER(SbImage)),
                ("crop_up", C.c_float), ("crop_down", C.c_float), ("crop_left", C.c_int), ("crop_right", C.c_int), ("crop_app_fill", C.c_int)]


# every symbol include/stitchb200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_F9 = _P(C.c_float)
API = {
    "sb_last_error": (C.c_char_p, []),
    "sb_version": (C.c_char_p, []),
    "sb_kernel_launch_count": (C.c_uint64, []),
    "sb_device_count": (C.c_int, []),
    "sb_selftest_division": (C.c_int, [C.c_int, C.c_ulonglong, C.c_uint, _P(C.c_ulonglong)]),
    "sb_host_alloc": (C.c_int, [_P(C.c_void_p), C.c_size_t]),
    "sb_host_free": (None, [C.c_void_p]),
    "sb_warper_create": (C.c_int, [C.c_int, C.c_float, C.c_int, _P(C.c_void_p)]),
    "sb_warper_destroy": (None, [C.c_void_p]),
    "sb_warper_get_scale": (C.c_float, [C.c_void_p]),
    "sb_warper_set_scale": (C.c_int, [C.c_void_p, C.c_float]),
    "sb_warper_set_translation": (C.c_int, [C.c_void_p, _F9]),
    "sb_warper_set_ab": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "sb_warper_warp_point": (C.c_int, [C.c_void_p, _F9, _F9, _F9, _F9]),
    "sb_warper_warp_roi": (C.c_int, [C.c_void_p, SbSize, _F9, _F9, _P(SbRect)]),
    "sb_warper_build_maps": (C.c_int, [C.c_void_p, SbSize, _F9, _F9, _P(SbImage), _P(SbImage), _P(SbRect)]),
    "sb_warper_warp": (C.c_int, [C.c_void_p, _P(SbImage), _F9, _F9, C.c_int, C.c_int, _P(SbImage), _P(SbPoint)]),
    "sb_warper_remap": (C.c_int, [C.c_void_p, _P(SbImage), C.c_int, C.c_int, _P(SbImage)]),
    "sb_warper_warp_backward": (C.c_int, [C.c_void_p, _P(SbImage), _F9, _F9, C.c_int, C.c_int, SbSize, _P(SbImage)]),
    "sb_remap": (C.c_int, [_P(SbImage), _P(SbImage), _P(SbImage), _P(SbImage), C.c_int, C.c_int, _P(C.c_uint8), C.c_int]),
    "sb_convert_maps": (C.c_int, [_P(SbImage), _P(SbImage), _P(SbImage), _P(SbImage), C.c_int, C.c_int]),
    "sb_comp_create": (C.c_int, [C.c_int, C.c_int, _P(C.c_void_p)]),
    "sb_comp_destroy": (None, [C.c_void_p]),
    "sb_comp_set_gains": (C.c_int, [C.c_void_p, _P(C.c_double), C.c_int]),
    "sb_comp_get_gains": (C.c_int, [C.c_void_p, _P(C.c_double), C.c_int]),
    "sb_comp_set_gain_maps": (C.c_int, [C.c_void_p, _P(SbImage), C.c_int]),
    "sb_comp_apply": (C.c_int, [C.c_void_p, C.c_int, SbPoint, _P(SbImage), _P(SbImage)]),
    "sb_comp_feed": (C.c_int, [C.c_void_p, _P(SbPoint), _P(SbImage), _P(SbImage), C.c_int]),
    "sb_gain_solve": (C.c_int, [C.c_int, C.c_int, _P(C.c_int), _P(C.c_int), _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_double)]),
    "sb_comp_set_block_size": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sb_comp_num_gains": (C.c_int, [C.c_void_p]),
    "sb_comp_gain_map_size": (C.c_int, [C.c_void_p, C.c_int, _P(SbSize)]),
    "sb_comp_get_gain_map": (C.c_int, [C.c_void_p, C.c_int, _P(SbImage)]),
    "sb_calibration_save": (C.c_int, [_P(SbCompositorConfig), C.c_char_p]),
    "sb_calibration_load": (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    "sb_calibration_config": (_P(SbCompositorConfig), [C.c_void_p]),
    "sb_calibration_free": (None, [C.c_void_p]),
    "sb_refine_seam_mask": (C.c_int, [_P(SbImage), _P(SbImage), _P(SbImage), C.c_int]),
    "sb_dilate3x3": (C.c_int, [_P(SbImage), _P(SbImage), C.c_int]),
    "sb_resize_linear_8u": (C.c_int, [_P(SbImage), _P(SbImage), C.c_int]),
    "sb_blender_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P(C.c_void_p)]),
    "sb_blender_destroy": (None, [C.c_void_p]),
    "sb_blender_num_bands": (C.c_int, [C.c_void_p]),
    "sb_blender_set_num_bands": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_blender_sharpness": (C.c_float, [C.c_void_p]),
    "sb_blender_set_sharpness": (C.c_int, [C.c_void_p, C.c_float]),
    "sb_blender_create_weight_maps": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbPoint), C.c_int, _P(SbImage), _P(SbRect)]),
    "sb_blender_prepare": (C.c_int, [C.c_void_p, _P(SbPoint), _P(SbSize), C.c_int]),
    "sb_blender_prepare_rect": (C.c_int, [C.c_void_p, SbRect]),
    "sb_blender_feed": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbImage), SbPoint]),
    "sb_blender_result_size": (C.c_int, [C.c_void_p, _P(SbSize)]),
    "sb_blender_blend": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbImage)]),
    "sb_normalize_using_weight_map": (C.c_int, [_P(SbImage), _P(SbImage), C.c_int]),
    "sb_create_weight_map": (C.c_int, [_P(SbImage), C.c_float, _P(SbImage), C.c_int]),
    "sb_create_laplace_pyr": (C.c_int, [_P(SbImage), C.c_int, _P(SbImage), C.c_int]),
    "sb_restore_image_from_laplace_pyr": (C.c_int, [_P(SbImage), C.c_int, C.c_int]),
    "sb_compositor_create": (C.c_int, [_P(SbCompositorConfig), C.c_int, _P(C.c_void_p)]),
    "sb_compositor_destroy": (None, [C.c_void_p]),
    "sb_compositor_pano_size": (C.c_int, [C.c_void_p, _P(SbSize)]),
    "sb_compositor_camera_roi": (C.c_int, [C.c_void_p, C.c_int, _P(SbRect)]),
    "sb_compositor_compose": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbImage), _P(SbImage)]),
    "sb_compositor_set_depth": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_kernel_plan": (C.c_int, [C.c_void_p]),
    "sb_multi_create": (C.c_int, [_P(SbCompositorConfig), C.c_int, _P(C.c_int), C.c_int, _P(C.c_void_p)]),
    "sb_multi_size": (C.c_int, [C.c_void_p]),
    "sb_multi_handle": (C.c_void_p, [C.c_void_p, C.c_int]),
    "sb_multi_run": (C.c_int, [C.c_void_p, C.c_int, _P(SbImage), _P(SbImage), _P(SbImage)]),
    "sb_multi_destroy": (None, [C.c_void_p]),
    "sb_debug_fs2_trace_dump": (C.c_int, []),
    "sb_compositor_enqueue": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbImage), _P(SbImage), _P(C.c_int)]),
    "sb_compositor_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_batch_create": (C.c_int, [C.c_void_p, C.c_int, _P(SbImage), _P(SbImage), _P(SbImage), _P(C.c_void_p)]),
    "sb_batch_launch": (C.c_int, [C.c_void_p]),
    "sb_batch_wait": (C.c_int, [C.c_void_p]),
    "sb_batch_last_gpu_ms": (C.c_int, [C.c_void_p, _P(C.c_float)]),
    "sb_batch_frames": (C.c_int, [C.c_void_p]),
    "sb_batch_mode": (C.c_int, [C.c_void_p]),
    "sb_batch_destroy": (None, [C.c_void_p]),
    "sb_compositor_last_gpu_ms": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "sb_compositor_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_marked_ms": (C.c_int, [C.c_void_p, _P(C.c_float)]),
    "sb_compositor_profile_frame": (C.c_int, [C.c_void_p, _P(SbImage), C.c_char_p, C.c_size_t]),
    "sb_compositor_num_bands": (C.c_int, [C.c_void_p]),
    "sb_compositor_set_strip": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sb_compositor_strip_peer_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _P(C.c_void_p), _P(C.c_size_t)]),
    "sb_compositor_strip_peer_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sb_compositor_strip_frame_peer": (C.c_int, [C.c_void_p, _P(SbImage)]),
    "sb_compositor_strip_peer_push": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sb_compositor_strip_peer_pull": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sb_compositor_set_strip_halo": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_strip_compose": (C.c_int, [C.c_void_p, _P(SbImage)]),
    "sb_compositor_strip_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(C.c_int), _P(C.c_int)]),
    "sb_compositor_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_compositor_strip_halo_bytes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _P(C.c_size_t), _P(C.c_size_t)]),
    "sb_compositor_strip_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sb_compositor_strip_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sb_compositor_strip_warp": (C.c_int, [C.c_void_p, _P(SbImage)]),
    "sb_compositor_strip_down": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_strip_band": (C.c_int, [C.c_void_p, C.c_int]),
    "sb_compositor_strip_result": (C.c_int, [C.c_void_p, _P(SbImage), _P(SbImage)]),
}

_lib = None


def lib():
    """Load libstitchb200.so.  Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StitchError(SB_ERR_CUDA, "%s is missing: build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in API.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != SB_OK:
        raise StitchError(rc, lib().sb_last_error().decode("utf-8", "replace"))


CV_16U, CV_16UC1, CV_16SC2 = 2, 2, 11      # fixed-point remap maps (cv::convertMaps)
_NP2CV = {np.dtype(np.uint8): CV_8U, np.dtype(np.int16): CV_16S, np.dtype(np.float32): CV_32F, np.dtype(np.uint16): CV_16U}
_CV2NP = {CV_8U: np.uint8, CV_16S: np.int16, CV_32F: np.float32, CV_16U: np.uint16}


def _depth(t):
    return t & 7


def _cn(t):
    return ((t >> 3) & 63) + 1


class DeviceImage:
    """GpuMat stand-in: a device pointer plus geometry.  `owner` keeps the allocation alive."""

    def __init__(self, ptr, rows, cols, cvtype, step, device, owner=None):
        self.ptr, self.rows, self.cols, self.type, self.step, self.device, self.owner = ptr, rows, cols, cvtype, step, device, owner

    @staticmethod
    def from_torch(t):
        import torch
        assert t.is_cuda and t.is_contiguous()
        dt = {torch.uint8: CV_8U, torch.int16: CV_16S, torch.float32: CV_32F}[t.dtype]
        cn = 1 if t.dim() == 2 else t.shape[2]
        return DeviceImage(t.data_ptr(), t.shape[0], t.shape[1], dt + ((cn - 1) << 3), t.stride(0) * t.element_size(),
                           t.device.index, owner=t)

    def sb(self):
        return SbImage(self.ptr, self.rows, self.cols, self.type, self.step, self.device)


def _image(a, output=False):
    """numpy array | DeviceImage -> (SbImage, keepalive).  An input whose pixels are not contiguous is copied; an OUTPUT
    must be written in place, so a view with non-contiguous pixels or a negative row stride is an error."""
    if isinstance(a, DeviceImage):
        return a.sb(), a
    a = np.asarray(a)
    if a.dtype not in _NP2CV:
        raise StitchError(SB_ERR_ASSERT, "unsupported dtype %s" % a.dtype)
    if a.ndim not in (2, 3):
        raise StitchError(SB_ERR_ASSERT, "image must be 2-D or 3-D")
    cn = 1 if a.ndim == 2 else a.shape[2]
    if a.strides[-1] != a.itemsize or (a.ndim == 3 and a.strides[1] != a.itemsize * cn) or a.strides[0] < 0:
        if output:
            raise StitchError(SB_ERR_ASSERT, "output array must have contiguous pixels and a non-negative row stride (got strides %s)" % (a.strides,))
        a = np.ascontiguousarray(a)
    return SbImage(a.ctypes.data, a.shape[0], a.shape[1], _NP2CV[a.dtype] + ((cn - 1) << 3), a.strides[0], -1), a


def _empty(rows, cols, cvtype):
    cn = _cn(cvtype)
    shape = (rows, cols) if cn == 1 else (rows, cols, cn)
    return np.empty(shape, _CV2NP[_depth(cvtype)])


def _f9(m):
    a = np.ascontiguousarray(m, np.float32)
    if a.size != 9:
        raise StitchError(SB_ERR_ASSERT, "K.size() == Size(3, 3) && K.type() == CV_32F")   # warpers.cpp:52-53
    return a.reshape(9)


def _fp(a):
    return a.ctypes.data_as(_F9)


def kernel_launch_count():
    return int(lib().sb_kernel_launch_count())


def selftest_division(n=1 << 26, seed=1, device=0):
    bad = C.c_ulonglong(0)
    _check(lib().sb_selftest_division(device, n, seed, C.byref(bad)))
    return int(bad.value)


def device_count():
    return int(lib().sb_device_count())


# ======================================================================================= warpers
class RotationWarper:
    """detail::RotationWarper (warpers.hpp:53-72) / RotationWarperBase (warpers.hpp:102-125)."""
    KIND = None

    def __init__(self, scale=1.0, device=0):
        self._h = C.c_void_p()
        self.device = device
        _check(lib().sb_warper_create(self.KIND, C.c_float(scale), device, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb_warper_destroy(self._h)
            self._h = None

    def getScale(self):
        return lib().sb_warper_get_scale(self._h)

    def setScale(self, v):
        _check(lib().sb_warper_set_scale(self._h, C.c_float(v)))

    def warpPoint(self, pt, K, R):
        K, R = _f9(K), _f9(R)
        p = np.asarray(pt, np.float32).reshape(2)
        uv = np.zeros(2, np.float32)
        _check(lib().sb_warper_warp_point(self._h, _fp(p), _fp(K), _fp(R), _fp(uv)))
        return float(uv[0]), float(uv[1])

    def warpRoi(self, src_size, K, R):
        K, R = _f9(K), _f9(R)
        r = SbRect()
        _check(lib().sb_warper_warp_roi(self._h, SbSize(*src_size), _fp(K), _fp(R), C.byref(r)))
        return (r.x, r.y, r.width, r.height)

    def buildMaps(self, src_size, K, R):
        """-> (Rect(tl, br) as (x, y, w, h), xmap, ymap); maps are float32 (h+1, w+1)."""
        K, R = _f9(K), _f9(R)
        r = SbRect()
        _check(lib().sb_warper_build_maps(self._h, SbSize(*src_size), _fp(K), _fp(R), None, None, C.byref(r)))
        xmap = np.empty((r.height + 1, r.width + 1), np.float32)
        ymap = np.empty_like(xmap)
        ix, _k1 = _image(xmap)
        iy, _k2 = _image(ymap)
        _check(lib().sb_warper_build_maps(self._h, SbSize(*src_size), _fp(K), _fp(R), C.byref(ix), C.byref(iy), C.byref(r)))
        self._map_shape = xmap.shape
        return (r.x, r.y, r.width, r.height), xmap, ymap

    def warp(self, src, K, R, interp_mode=INTER_LINEAR, border_mode=BORDER_REFLECT):
        """-> (tl, dst)  (warpers_inl.hpp:88-99); dst is (roi.height+1, roi.width+1) of src's type."""
        K, R = _f9(K), _f9(R)
        isrc, keep = _image(src)
        roi = self.warpRoi((isrc.cols, isrc.rows), K, R)
        dst = _empty(roi[3], roi[2], isrc.type)
        idst, _k = _image(dst, output=True)
        tl = SbPoint()
        _check(lib().sb_warper_warp(self._h, C.byref(isrc), _fp(K), _fp(R), interp_mode, border_mode, C.byref(idst), C.byref(tl)))
        return (tl.x, tl.y), dst

    def remap(self, src, interp_mode=INTER_LINEAR, border_mode=BORDER_REFLECT):
        """cv::remap with the maps cached by the last buildMaps (the app's video path, APP64:752)."""
        isrc, keep = _image(src)
        if getattr(self, "_map_shape", None) is None:
            raise StitchError(SB_ERR_ASSERT, "remap: no cached maps; call buildMaps first")
        dst = _empty(self._map_shape[0], self._map_shape[1], isrc.type)
        idst, _k = _image(dst, output=True)
        _check(lib().sb_warper_remap(self._h, C.byref(isrc), interp_mode, border_mode, C.byref(idst)))
        return dst

    def warpBackward(self, src, K, R, interp_mode, border_mode, dst_size):
        K, R = _f9(K), _f9(R)
        isrc, keep = _image(src)
        dst = _empty(dst_size[1], dst_size[0], isrc.type)
        idst, _k = _image(dst, output=True)
        _check(lib().sb_warper_warp_backward(self._h, C.byref(isrc), _fp(K), _fp(R), interp_mode, border_mode,
                                             SbSize(*dst_size), C.byref(idst)))
        return dst


class PlaneWarper(RotationWarper):
    KIND = WARP_PLANE

    def setTranslation(self, T):
        t = np.ascontiguousarray(T, np.float32).reshape(3)
        _check(lib().sb_warper_set_translation(self._h, _fp(t)))


class _ABWarper(RotationWarper):
    """CompressedRectilinear / Panini warpers: (scale, A = 1, B = 1) (warpers.hpp:227-299)."""

    def __init__(self, scale=1.0, A=1.0, B=1.0, device=0):
        super().__init__(scale, device)
        _check(lib().sb_warper_set_ab(self._h, C.c_float(A), C.c_float(B)))


class FisheyeWarper(RotationWarper):
    KIND = WARP_FISHEYE


class StereographicWarper(RotationWarper):
    KIND = WARP_STEREOGRAPHIC


class CompressedRectilinearWarper(_ABWarper):
    KIND = WARP_COMPRESSED_RECTILINEAR


class CompressedRectilinearPortraitWarper(_ABWarper):
    KIND = WARP_COMPRESSED_RECTILINEAR_PORTRAIT


class PaniniWarper(_ABWarper):
    KIND = WARP_PANINI


class PaniniPortraitWarper(_ABWarper):
    KIND = WARP_PANINI_PORTRAIT


class MercatorWarper(RotationWarper):
    KIND = WARP_MERCATOR


class TransverseMercatorWarper(RotationWarper):
    KIND = WARP_TRANSVERSE_MERCATOR


class SphericalPortraitWarper(RotationWarper):
    KIND = WARP_SPHERICAL_PORTRAIT


class CylindricalPortraitWarper(RotationWarper):
    KIND = WARP_CYLINDRICAL_PORTRAIT


class PlanePortraitWarper(RotationWarper):
    KIND = WARP_PLANE_PORTRAIT


class CylindricalWarper(RotationWarper):
    KIND = WARP_CYLINDRICAL


class SphericalWarper(RotationWarper):
    KIND = WARP_SPHERICAL


def convertMaps(xmap, ymap, nninterpolation=False, device=0):
    """cv::convertMaps(xmap, ymap, CV_16SC2): float maps -> (map1 int16 HxWx2, map2 uint16 HxW | None)."""
    ix, k1 = _image(np.ascontiguousarray(xmap, np.float32))
    iy, k2 = _image(np.ascontiguousarray(ymap, np.float32))
    map1 = np.empty((ix.rows, ix.cols, 2), np.int16)
    map2 = None if nninterpolation else np.empty((ix.rows, ix.cols), np.uint16)
    i1, k3 = _image(map1)
    i2 = SbImage(None, 0, 0, CV_16UC1, 0, -1)
    if map2 is not None:
        i2, k4 = _image(map2)
    _check(lib().sb_convert_maps(C.byref(ix), C.byref(iy), C.byref(i1), C.byref(i2), 1 if nninterpolation else 0, device))
    return map1, map2


def remap(src, xmap, ymap, interp_mode=INTER_LINEAR, border_mode=BORDER_CONSTANT, border_value=(0, 0, 0, 0), device=0):
    """cv::remap.  (xmap, ymap): float32 maps, or the fixed-point pair of convertMaps (int16 HxWx2, uint16 HxW | None)."""
    isrc, k0 = _image(src)
    xmap = np.asarray(xmap)
    if xmap.dtype == np.int16:
        ix, k1 = _image(np.ascontiguousarray(xmap))
        if ymap is None:
            iy = SbImage(None, 0, 0, CV_16UC1, 0, -1)
        else:
            iy, k2 = _image(np.ascontiguousarray(ymap, np.uint16))
    else:
        ix, k1 = _image(np.ascontiguousarray(xmap, np.float32))
        iy, k2 = _image(np.ascontiguousarray(ymap, np.float32))
    dst = _empty(ix.rows, ix.cols, isrc.type)
    idst, k3 = _image(dst, output=True)
    bv = (C.c_uint8 * 4)(*border_value)
    _check(lib().sb_remap(C.byref(isrc), C.byref(idst), C.byref(ix), C.byref(iy), interp_mode, border_mode, bv, device))
    return dst


# ======================================================================================= exposure
class ExposureCompensator:
    """detail::ExposureCompensator (exposure_compensate.hpp:51-64).  feed() is calibration: hand its result in
    (setGains / setGainMaps) or let feed() estimate it (overlap statistics on the device)."""
    NO, GAIN, GAIN_BLOCKS = COMP_NO, COMP_GAIN, COMP_GAIN_BLOCKS
    KIND = COMP_NO

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().sb_comp_create(self.KIND, device, C.byref(self._h)))

    @staticmethod
    def createDefault(kind, device=0):
        cls = {COMP_NO: NoExposureCompensator, COMP_GAIN: GainCompensator, COMP_GAIN_BLOCKS: BlocksGainCompensator}.get(kind)
        if cls is None:
            h = C.c_void_p()
            _check(lib().sb_comp_create(kind, device, C.byref(h)))     # raises CV_StsBadArg
        return cls(device)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb_comp_destroy(self._h)
            self._h = None

    def feed(self, corners, images, masks):
        """ExposureCompensator::feed(corners, images, masks) (exposure_compensate.cpp:64-71, 76-147, 165-222):
        images uint8 HxWx3, masks uint8 HxW (level 255), host arrays or DeviceImages."""
        n = len(images)
        if not (len(corners) == n and len(masks) == n):
            raise StitchError(SB_ERR_ASSERT, "corners.size() == images.size() && images.size() == masks.size()")
        im, mk = [_image(a) for a in images], [_image(a) for a in masks]
        ai, am = (SbImage * max(n, 1))(*[x[0] for x in im]), (SbImage * max(n, 1))(*[x[0] for x in mk])
        pts = (SbPoint * max(n, 1))(*[SbPoint(int(c[0]), int(c[1])) for c in corners])
        _check(lib().sb_comp_feed(self._h, pts, ai, am, n))
        self._n = n

    def apply(self, index, corner, image, mask=None):
        """In place on a host uint8 array or a DeviceImage (exposure_compensate.cpp:150-153, 225-246)."""
        img, keep = _image(image)
        if keep is not image and not isinstance(image, DeviceImage):
            raise StitchError(SB_ERR_ASSERT, "apply() works in place: pass a contiguous array")
        _check(lib().sb_comp_apply(self._h, index, SbPoint(*corner), C.byref(img), None))
        return image


class NoExposureCompensator(ExposureCompensator):
    KIND = COMP_NO


class GainCompensator(ExposureCompensator):
    KIND = COMP_GAIN

    def setGains(self, gains):
        g = np.ascontiguousarray(gains, np.float64)
        _check(lib().sb_comp_set_gains(self._h, g.ctypes.data_as(_P(C.c_double)), len(g)))
        self._n = len(g)

    def gains(self):
        g = np.zeros(getattr(self, "_n", 0), np.float64)
        _check(lib().sb_comp_get_gains(self._h, g.ctypes.data_as(_P(C.c_double)), len(g)))
        return list(g)


class BlocksGainCompensator(ExposureCompensator):
    KIND = COMP_GAIN_BLOCKS

    def setGainMaps(self, maps):
        keep = [np.ascontiguousarray(m, np.float32) for m in maps]
        arr = (SbImage * len(keep))(*[_image(m)[0] for m in keep])
        _check(lib().sb_comp_set_gain_maps(self._h, arr, len(keep)))

    def setBlockSize(self, bl_width, bl_height):
        """BlocksGainCompensator(bl_width = 32, bl_height = 32) (exposure_compensate.hpp:92)"""
        _check(lib().sb_comp_set_block_size(self._h, bl_width, bl_height))

    def gainMaps(self):
        """gain_maps_ as estimated by feed(): list of float32 arrays (block grid of every image)"""
        out = []
        for i in range(lib().sb_comp_num_gains(self._h)):
            sz = SbSize()
            _check(lib().sb_comp_gain_map_size(self._h, i, C.byref(sz)))
            m = np.empty((sz.height, sz.width), np.float32)
            im, _ = _image(m)
            _check(lib().sb_comp_get_gain_map(self._h, i, C.byref(im)))
            out.append(m)
        return out


def gain_solve(n, pi, pj, N, Iij, Iji):
    """sb_gain_solve: the normal equations of GainCompensator::feed (exposure_compensate.cpp:128-144) from pair statistics"""
    pi, pj = np.ascontiguousarray(pi, np.int32), np.ascontiguousarray(pj, np.int32)
    N, Iij, Iji = (np.ascontiguousarray(a, np.float64) for a in (N, Iij, Iji))
    g = np.zeros(n, np.float64)
    d = _P(C.c_double)
    _check(lib().sb_gain_solve(n, len(pi), pi.ctypes.data_as(_P(C.c_int)), pj.ctypes.data_as(_P(C.c_int)), N.ctypes.data_as(d),
                               Iij.ctypes.data_as(d), Iji.ctypes.data_as(d), g.ctypes.data_as(d)))
    return g


def refine_seam_mask(seam_mask, mask_warped, device=0):
    """stitcher.cpp:291-294: dilate(seam_mask) -> resize to mask_warped.size() (INTER_LINEAR) -> & mask_warped"""
    a, k0 = _image(np.ascontiguousarray(seam_mask, np.uint8))
    b, k1 = _image(np.ascontiguousarray(mask_warped, np.uint8))
    out = np.empty((b.rows, b.cols), np.uint8)
    o, k2 = _image(out, output=True)
    _check(lib().sb_refine_seam_mask(C.byref(a), C.byref(b), C.byref(o), device))
    return out


def dilate3x3(src, device=0):
    """cv::dilate(src, dst, Mat()) on uint8 HxW"""
    a, k0 = _image(np.ascontiguousarray(src, np.uint8))
    out = np.empty((a.rows, a.cols), np.uint8)
    o, k1 = _image(out, output=True)
    _check(lib().sb_dilate3x3(C.byref(a), C.byref(o), device))
    return out


def resize_linear_8u(src, dsize_wh, device=0):
    """cv::resize(src, dsize, INTER_LINEAR) on uint8 HxW"""
    a, k0 = _image(np.ascontiguousarray(src, np.uint8))
    out = np.empty((dsize_wh[1], dsize_wh[0]), np.uint8)
    o, k1 = _image(out, output=True)
    _check(lib().sb_resize_linear_8u(C.byref(a), C.byref(o), device))
    return out


# ======================================================================================= blenders
class Blender:
    """detail::Blender (blenders.hpp:53-69): puts one image over another."""
    NO, FEATHER, MULTI_BAND = BLEND_NO, BLEND_FEATHER, BLEND_MULTI_BAND
    KIND = BLEND_NO

    def __init__(self, device=0, _num_bands=5, _weight_type=CV_32F, _sharpness=0.02):
        self._h = C.c_void_p()
        self.device = device
        _check(lib().sb_blender_create(self.KIND, _num_bands, _weight_type, C.c_float(_sharpness), device, C.byref(self._h)))

    @staticmethod
    def createDefault(kind, try_gpu=False, device=0):
        cls = {BLEND_NO: Blender, BLEND_FEATHER: FeatherBlender, BLEND_MULTI_BAND: MultiBandBlender}.get(kind)
        if cls is None:
            h = C.c_void_p()
            _check(lib().sb_blender_create(kind, 5, CV_32F, C.c_float(0.02), device, C.byref(h)))   # CV_StsBadArg
        return cls(device=device)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb_blender_destroy(self._h)
            self._h = None

    def prepare(self, corners_or_rect, sizes=None):
        if sizes is None:
            x, y, w, h = corners_or_rect
            _check(lib().sb_blender_prepare_rect(self._h, SbRect(x, y, w, h)))
        else:
            n = len(corners_or_rect)
            if n != len(sizes):
                raise StitchError(SB_ERR_ASSERT, "sizes.size() == corners.size()")     # util.cpp:129
            pts = (SbPoint * n)(*[SbPoint(int(c[0]), int(c[1])) for c in corners_or_rect])
            szs = (SbSize * n)(*[SbSize(int(s[0]), int(s[1])) for s in sizes])
            _check(lib().sb_blender_prepare(self._h, pts, szs, n))

    def feed(self, img, mask, tl):
        i, k0 = _image(img)
        m, k1 = _image(mask)
        _check(lib().sb_blender_feed(self._h, C.byref(i), C.byref(m), SbPoint(int(tl[0]), int(tl[1]))))

    def blend(self):
        """-> (dst int16 HxWx3, dst_mask uint8 HxW)."""
        s = SbSize()
        _check(lib().sb_blender_result_size(self._h, C.byref(s)))
        dst = np.empty((s.height, s.width, 3), np.int16)
        dmask = np.empty((s.height, s.width), np.uint8)
        i, k0 = _image(dst, output=True)
        m, k1 = _image(dmask, output=True)
        _check(lib().sb_blender_blend(self._h, C.byref(i), C.byref(m)))
        return dst, dmask


class FeatherBlender(Blender):
    KIND = BLEND_FEATHER

    def __init__(self, sharpness=0.02, device=0):
        super().__init__(device=device, _sharpness=sharpness)

    def sharpness(self):
        return lib().sb_blender_sharpness(self._h)

    def setSharpness(self, v):
        _check(lib().sb_blender_set_sharpness(self._h, C.c_float(v)))

    def createWeightMaps(self, masks, corners):
        """FeatherBlender::createWeightMaps (blenders.hpp:80-81, blenders.cpp:158-186) -> (dst_roi (x, y, w, h), weight_maps)"""
        n = len(masks)
        if len(corners) != n or n == 0:
            raise StitchError(SB_ERR_ASSERT, "masks.size() == corners.size()")
        mk = [_image(np.ascontiguousarray(m, np.uint8)) for m in masks]
        maps = [np.empty((m[0].rows, m[0].cols), np.float32) for m in mk]
        wm = [_image(w) for w in maps]
        am, aw = (SbImage * n)(*[x[0] for x in mk]), (SbImage * n)(*[x[0] for x in wm])
        pts = (SbPoint * n)(*[SbPoint(int(c[0]), int(c[1])) for c in corners])
        roi = SbRect()
        _check(lib().sb_blender_create_weight_maps(self._h, am, pts, n, aw, C.byref(roi)))
        return (roi.x, roi.y, roi.width, roi.height), maps


class MultiBandBlender(Blender):
    KIND = BLEND_MULTI_BAND

    def __init__(self, try_gpu=False, num_bands=5, weight_type=CV_32F, device=0):
        super().__init__(device=device, _num_bands=num_bands, _weight_type=weight_type)

    def numBands(self):
        return lib().sb_blender_num_bands(self._h)

    def setNumBands(self, v):
        _check(lib().sb_blender_set_num_bands(self._h, int(v)))


def normalizeUsingWeightMap(weight, src, device=0):
    w, k0 = _image(weight)
    s, k1 = _image(src, output=True)      # normalised in place
    _check(lib().sb_normalize_using_weight_map(C.byref(w), C.byref(s), device))
    return src


def createWeightMap(mask, sharpness, device=0):
    m, k0 = _image(mask)
    out = np.empty((m.rows, m.cols), np.float32)
    o, k1 = _image(out, output=True)
    _check(lib().sb_create_weight_map(C.byref(m), C.c_float(sharpness), C.byref(o), device))
    return out


def createLaplacePyr(img, num_levels, device=0):
    i, k0 = _image(img)
    pyr, r, c = [], i.rows, i.cols
    for _ in range(num_levels + 1):
        pyr.append(np.empty((r, c, 3), np.int16))
        r, c = (r + 1) // 2, (c + 1) // 2
    arr = (SbImage * len(pyr))(*[_image(p)[0] for p in pyr])
    _check(lib().sb_create_laplace_pyr(C.byref(i), num_levels, arr, device))
    return pyr


def restoreImageFromLaplacePyr(pyr, device=0):
    pyr = [np.ascontiguousarray(p).copy() for p in pyr]
    arr = (SbImage * len(pyr))(*[_image(p)[0] for p in pyr])
    _check(lib().sb_restore_image_from_laplace_pyr(arr, len(pyr), device))
    return pyr[0]


# ======================================================================================= compositor
def save_calibration(cfg, path):
    """sb_calibration_save of an SbCompositorConfig (or a pointer to one)"""
    _check(lib().sb_calibration_save(cfg if isinstance(cfg, _P(SbCompositorConfig)) else C.byref(cfg), os.fsencode(path)))


def load_calibration(path):
    """sb_calibration_load -> dict of plain numpy values (the file's content; no device needed)"""
    cal = C.c_void_p()
    _check(lib().sb_calibration_load(os.fsencode(path), C.byref(cal)))
    try:
        c = lib().sb_calibration_config(cal).contents
        n = c.n_cameras

        def img(im):
            dt = np.uint8 if im.type == CV_8UC1 else np.float32
            a = np.ctypeslib.as_array(C.cast(im.data, _P(C.c_uint8)), shape=(im.rows, im.step))
            return a[:, :im.cols * np.dtype(dt).itemsize].copy().view(dt)
        return {"n_cameras": n, "src_size": (c.src_size.width, c.src_size.height), "warper_kind": c.warper_kind,
                "warper_scale": c.warper_scale, "K": np.ctypeslib.as_array(c.K, shape=(n, 3, 3)).copy(),
                "R": np.ctypeslib.as_array(c.R, shape=(n, 3, 3)).copy(), "blender_kind": c.blender_kind, "num_bands": c.num_bands,
                "weight_type": c.weight_type, "sharpness": c.sharpness, "comp_kind": c.comp_kind, "output_type": c.output_type,
                "gains": np.ctypeslib.as_array(c.gains, shape=(n,)).copy() if c.gains else None,
                "seam_masks": [img(c.seam_masks[i]) for i in range(n)] if c.seam_masks else None,
                "gain_maps": [img(c.gain_maps[i]) for i in range(n)] if c.gain_maps else None}
    finally:
        lib().sb_calibration_free(cal)


class Batch:
    """sb_batch: a recorded lap of frame sets, replayed with one host call (see Compositor.batch)."""

    def __init__(self, comp, frame_sets, panos, pano_masks=None):
        n, nf = comp.n, len(frame_sets)
        assert nf > 0 and len(panos) == nf and (pano_masks is None or len(pano_masks) == nf)
        self._keep = [comp]
        srcs = (SbImage * (nf * n))()
        for f, frames in enumerate(frame_sets):
            arr, keep = comp._srcs(frames)
            self._keep.append(keep)
            for i in range(n):
                srcs[f * n + i] = arr[i]
        self.panos = (SbImage * nf)()
        for f, pano in enumerate(panos):
            if pano is None:
                self.panos[f] = SbImage(None, 0, 0, comp.output_type, 0, -1)
            else:
                self.panos[f], k = _image(pano, output=True)
                self._keep.append(k)
        self.masks = None
        if pano_masks is not None:
            self.masks = (SbImage * nf)()
            for f, m in enumerate(pano_masks):
                if m is None:
                    self.masks[f] = SbImage(None, 0, 0, CV_8UC1, 0, -1)
                else:
                    self.masks[f], k = _image(m, output=True)
                    self._keep.append(k)
        self._srcs_arr = srcs
        h = C.c_void_p()
        _check(lib().sb_compositor_batch_create(comp._h, nf, srcs, self.panos, self.masks, C.byref(h)))
        self._h = h
        self.n_frames = nf
        self.mode = int(lib().sb_batch_mode(h))       # 1: one persistent launch per lap, 0: stream replay, 2: CUDA graph

    def launch(self):
        _check(_lib.sb_batch_launch(self._h))

    def wait(self):
        _check(_lib.sb_batch_wait(self._h))

    def last_gpu_ms(self):
        ms = C.c_float()
        _check(lib().sb_batch_last_gpu_ms(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001 - interpreter shutdown
            pass


class MultiCompositor:
    """sb_multi: one process driving several GPUs, frame f on devices[f % n] (SURVEY.md §8e throughput mode)."""

    def __init__(self, devices, depth, *args, **kw):
        proto = Compositor.__new__(Compositor)
        proto._h = None
        proto._build_config(*args, **kw)
        self._proto = proto                                   # keeps the config's arrays alive
        self.n = proto.n
        self.output_type = proto._cfg.contents.output_type
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        _check(lib().sb_multi_create(proto._cfg, len(devices), devs, depth, C.byref(self._h)))
        s = SbSize()
        _check(lib().sb_compositor_pano_size(lib().sb_multi_handle(self._h, 0), C.byref(s)))
        self.pano_size = (s.width, s.height)

    def run(self, frame_sets, with_masks=True):
        """frame_sets[f] = the n frames of frame set f (host arrays) -> list of (pano, mask) in frame order."""
        nf, w, h = len(frame_sets), self.pano_size[0], self.pano_size[1]
        srcs = (SbImage * (nf * self.n))()
        keep = []
        for f, frames in enumerate(frame_sets):
            for i, fr in enumerate(frames):
                srcs[f * self.n + i], k = _image(fr)
                keep.append(k)
        panos, masks = [_empty(h, w, self.output_type) for _ in range(nf)], [np.empty((h, w), np.uint8) for _ in range(nf)]
        ip = (SbImage * nf)(*[_image(p, output=True)[0] for p in panos])
        im = (SbImage * nf)(*[_image(m, output=True)[0] for m in masks]) if with_masks else None
        _check(lib().sb_multi_run(self._h, nf, srcs, ip, im))
        return list(zip(panos, masks))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb_multi_destroy(self._h)
            self._h = None


_WARPER_KINDS = {"plane": WARP_PLANE, "cylindrical": WARP_CYLINDRICAL, "spherical": WARP_SPHERICAL, "fisheye": WARP_FISHEYE,
                 "stereographic": WARP_STEREOGRAPHIC, "compressedRectilinear": WARP_COMPRESSED_RECTILINEAR,
                 "compressedRectilinearPortrait": WARP_COMPRESSED_RECTILINEAR_PORTRAIT, "panini": WARP_PANINI, "paniniPortrait": WARP_PANINI_PORTRAIT,
                 "mercator": WARP_MERCATOR, "transverseMercator": WARP_TRANSVERSE_MERCATOR, "sphericalPortrait": WARP_SPHERICAL_PORTRAIT,
                 "cylindricalPortrait": WARP_CYLINDRICAL_PORTRAIT, "planePortrait": WARP_PLANE_PORTRAIT}


class Compositor:
    """The per-frame loop of Stitcher::composePanorama (stitcher.cpp:221-313) with calibration fixed."""

    def __init__(self, src_size, Ks, Rs, warper="spherical", scale=None, blender="multiband", num_bands=5,
                 weight_type=CV_32F, sharpness=0.02, gains=None, seam_masks=None, output_type=CV_8UC3, device=0,
                 gain_maps=None, warper_ab=None, undistort_maps=None, crop=None, crop_app_fill=False):
        self._build_config(src_size, Ks, Rs, warper, scale, blender, num_bands, weight_type, sharpness, gains, seam_masks, output_type,
                           device, gain_maps, warper_ab, undistort_maps, crop, crop_app_fill)
        self._create(device)

    def _build_config(self, src_size, Ks, Rs, warper="spherical", scale=None, blender="multiband", num_bands=5,
                      weight_type=CV_32F, sharpness=0.02, gains=None, seam_masks=None, output_type=CV_8UC3, device=0,
                      gain_maps=None, warper_ab=None, undistort_maps=None, crop=None, crop_app_fill=False):
        """warper: "plane" / "cylindrical" / "spherical" or any WARP_* kind (warper_ab = (a, b) for the compressed-rectilinear /
        Panini projectors); undistort_maps: per camera the (map1 CV_16SC2, map2 CV_16UC1) pair of initUndistortRectifyMap
        (the app's fisheye front end, APP64:201-238); crop = (up, down, left, right): the live app's crop margins
        (fractions of the height, pixels; APP64:47, 150-177, 702), crop_app_fill its unconditional gather for uncovered pixels."""
        n = len(Ks)
        self.n = n
        self.device = device
        self.src_size = tuple(src_size)
        K = np.ascontiguousarray(np.stack([_f9(k) for k in Ks]), np.float32)
        R = np.ascontiguousarray(np.stack([_f9(r) for r in Rs]), np.float32)
        cfg = SbCompositorConfig()
        cfg.n_cameras = n
        cfg.src_size = SbSize(*self.src_size)
        cfg.warper_kind = _WARPER_KINDS.get(warper, warper)
        cfg.warper_scale = float(scale)
        if warper_ab is not None:
            cfg.warper_a, cfg.warper_b = float(warper_ab[0]), float(warper_ab[1])
        ukeep = None
        if undistort_maps is not None:
            u1 = [np.ascontiguousarray(m[0], np.int16) for m in undistort_maps]
            u2 = [np.ascontiguousarray(m[1], np.uint16) for m in undistort_maps]
            a1 = (SbImage * n)(*[_image(m)[0] for m in u1])
            a2 = (SbImage * n)(*[_image(m)[0] for m in u2])
            cfg.undistort_map1, cfg.undistort_map2 = a1, a2
            ukeep = (u1, u2, a1, a2)
        if crop is not None:
            cfg.crop_up, cfg.crop_down, cfg.crop_left, cfg.crop_right = float(crop[0]), float(crop[1]), int(crop[2]), int(crop[3])
        cfg.crop_app_fill = int(bool(crop_app_fill))
        cfg.K = K.ctypes.data_as(_F9)
        cfg.R = R.ctypes.data_as(_F9)
        cfg.blender_kind = {"no": BLEND_NO, "feather": BLEND_FEATHER, "multiband": BLEND_MULTI_BAND}.get(blender, blender)
        cfg.num_bands = num_bands
        cfg.weight_type = weight_type
        cfg.sharpness = sharpness
        g = gkeep = garr = arr = None
        if gains is not None:
            g = np.ascontiguousarray(gains, np.float64)
            cfg.comp_kind = COMP_GAIN
            cfg.gains = g.ctypes.data_as(_P(C.c_double))
        elif gain_maps is not None:          # BlocksGainCompensator: one float32 block gain map per camera
            gkeep = [np.ascontiguousarray(m, np.float32) for m in gain_maps]
            garr = (SbImage * n)(*[_image(m)[0] for m in gkeep])
            cfg.comp_kind = COMP_GAIN_BLOCKS
            cfg.gain_maps = garr
        else:
            cfg.comp_kind = COMP_NO
        keep = None
        if seam_masks is not None:
            keep = [np.ascontiguousarray(m, np.uint8) for m in seam_masks]
            arr = (SbImage * n)(*[_image(m)[0] for m in keep])
            cfg.seam_masks = arr
        cfg.output_type = output_type
        self._cal = None
        self._cfg, self._cfg_keep = C.pointer(cfg), (K, R, g, keep, gkeep, garr, arr, ukeep)      # (the config points into these)

    def _create(self, device):
        cfg = self._cfg.contents
        self.n, self.device, self.src_size = cfg.n_cameras, device, (cfg.src_size.width, cfg.src_size.height)
        self.output_type = cfg.output_type
        self._h = C.c_void_p()
        _check(lib().sb_compositor_create(self._cfg, device, C.byref(self._h)))
        s = SbSize()
        _check(lib().sb_compositor_pano_size(self._h, C.byref(s)))
        self.pano_size = (s.width, s.height)

    def save_calibration(self, path):
        """Everything the calibration handed to this compositor (K, R, scale, blender settings, gains / gain maps,
        seam masks) into one checksummed file (sb_calibration_save)."""
        save_calibration(self._cfg, path)

    @classmethod
    def from_calibration(cls, path, device=0):
        """Resume from a file written by save_calibration: the device tables are rebuilt, bit-identical."""
        self = cls.__new__(cls)
        self._h = None
        self._cal = C.c_void_p()
        _check(lib().sb_calibration_load(os.fsencode(path), C.byref(self._cal)))
        self._cfg, self._cfg_keep = lib().sb_calibration_config(self._cal), None
        self._create(device)
        return self

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb_compositor_destroy(self._h)
            self._h = None
        if getattr(self, "_cal", None) and _lib is not None:
            _lib.sb_calibration_free(self._cal)
            self._cal = None

    def camera_roi(self, i):
        r = SbRect()
        _check(lib().sb_compositor_camera_roi(self._h, i, C.byref(r)))
        return (r.x, r.y, r.width, r.height)

    def _srcs(self, frames):
        if len(frames) != self.n:
            raise StitchError(SB_ERR_ASSERT, "expected %d frames" % self.n)
        pairs = [_image(f) for f in frames]
        return (SbImage * self.n)(*[p[0] for p in pairs]), pairs

    def new_output(self):
        w, h = self.pano_size
        return _empty(h, w, self.output_type), np.empty((h, w), np.uint8)

    def compose(self, frames, pano=None, pano_mask=None):
        """frames: n uint8 (H, W, 3) arrays or DeviceImages -> (pano, pano_mask)."""
        if pano is None:
            pano, pano_mask = self.new_output()
        arr, keep = self._srcs(frames)
        ip, k0 = _image(pano, output=True)
        im = None
        if pano_mask is not None:
            im, k1 = _image(pano_mask, output=True)
        _check(lib().sb_compositor_compose(self._h, arr, C.byref(ip), C.byref(im) if im is not None else None))
        return pano, pano_mask

    def set_fused(self, fused):
        """True/1: fused kernels; False/0: staged feed/blend-shaped path; 10/11: fused with kernel variant 0/1."""
        _check(lib().sb_compositor_set_fused(self._h, int(fused)))

    def kernel_plan(self):
        """2 / 1 / 0: which generation of the frame kernels this calibration runs on (sb_compositor_kernel_plan)."""
        return int(lib().sb_compositor_kernel_plan(self._h))

    def set_depth(self, depth):
        _check(lib().sb_compositor_set_depth(self._h, depth))

    def enqueue(self, frames, pano, pano_mask=None):
        """pano=None: the panorama stays in the slot's device buffer (returned SbImage is lent)."""
        arr, keep = self._srcs(frames)
        if pano is None:
            ip = SbImage(None, 0, 0, self.output_type, 0, -1)
            self.last_lent = ip
        else:
            ip, k0 = _image(pano, output=True)
        im = None
        if pano_mask is not None:
            im, k1 = _image(pano_mask, output=True)
        slot = C.c_int(-1)
        _check(lib().sb_compositor_enqueue(self._h, arr, C.byref(ip), C.byref(im) if im is not None else None, C.byref(slot)))
        return slot.value

    def prepare_call(self, frames, pano, pano_mask=None):
        """The argument block of enqueue(frames, pano, pano_mask), built once for buffers that are reused frame after
        frame (a video loop cycles through a few input / output buffers): enqueue_prepared then costs one C call."""
        arr, keep = self._srcs(frames)
        if pano is None:
            ip, k0 = SbImage(None, 0, 0, self.output_type, 0, -1), None
        else:
            ip, k0 = _image(pano, output=True)
        im, k1 = _image(pano_mask, output=True) if pano_mask is not None else (None, None)
        return (arr, C.byref(ip), C.byref(im) if im is not None else None, C.c_int(-1), ip if pano is None else None, (keep, ip, im, k0, k1))

    def enqueue_prepared(self, call):
        lent = call[4]
        if lent is not None:                         # pano=None: the library lends the slot's buffer through this struct every time
            lent.data, lent.rows, lent.cols, lent.step, lent.device = None, 0, 0, 0, -1
            self.last_lent = lent
        _check(_lib.sb_compositor_enqueue(self._h, call[0], call[1], call[2], C.byref(call[3])))
        return call[3].value

    def wait(self, slot):
        _check(lib().sb_compositor_wait(self._h, slot))

    def batch(self, frame_sets, panos, pano_masks=None):
        """One lap of a buffer ring as a CUDA graph (sb_compositor_batch_create): frame_sets[f] = the n source images of
        frame f (numpy arrays or DeviceImage), panos[f] = the output array of frame f or None (stays in the slot's device
        buffer), pano_masks likewise or None for no masks at all.  -> Batch (launch / wait / last_gpu_ms)."""
        return Batch(self, frame_sets, panos, pano_masks)

    def mark(self, which):
        _check(lib().sb_compositor_mark(self._h, which))

    def marked_ms(self):
        ms = C.c_float()
        _check(lib().sb_compositor_marked_ms(self._h, C.byref(ms)))
        return ms.value

    def profile_frame(self, frames):
        """-> list of {"name", "ms", "bytes"} per kernel launch of one frame (CUDA events on the stream)."""
        import json
        arr, keep = self._srcs(frames)
        buf = C.create_string_buffer(1 << 16)
        _check(lib().sb_compositor_profile_frame(self._h, arr, buf, len(buf)))
        return json.loads(buf.value.decode())

    def last_gpu_ms(self, slot=0):
        ms = C.c_float()
        _check(lib().sb_compositor_last_gpu_ms(self._h, slot, C.byref(ms)))
        return ms.value


    # ---- latency ("strip") mode: see stitchingvideo_b200/strips.py for the driver ----
    @property
    def num_bands(self):
        return int(lib().sb_compositor_num_bands(self._h))

    def set_strip(self, rank, world):
        _check(lib().sb_compositor_set_strip(self._h, rank, world))

    def set_strip_halo(self, recompute):
        _check(lib().sb_compositor_set_strip_halo(self._h, 1 if recompute else 0))

    def strip_compose(self, frames):
        """Recompute-halo mode: every stage of one frame on this rank's columns, no exchange; asynchronous."""
        arr, keep = self._srcs(frames)
        _check(lib().sb_compositor_strip_compose(self._h, arr))

    def strip_peer_export(self, side):
        """This rank's receive area of `side` for the peer-memory halo exchange -> (64-byte CUDA IPC handle, device pointer)."""
        h, ptr, n = C.create_string_buffer(64), C.c_void_p(), C.c_size_t()
        _check(lib().sb_compositor_strip_peer_export(self._h, side, h, C.byref(ptr), C.byref(n)))
        return h.raw, ptr.value

    def strip_peer_connect(self, side, ipc_handle=None, same_process_ptr=None):
        """Map the neighbour's receive area (of ITS opposite side): its IPC handle, or its pointer inside one process."""
        hb = C.create_string_buffer(ipc_handle, 64) if ipc_handle is not None else None
        _check(lib().sb_compositor_strip_peer_connect(self._h, side, hb, C.c_void_p(same_process_ptr) if same_process_ptr is not None else None))

    def strip_peer_push(self, what, level):
        _check(lib().sb_compositor_strip_peer_push(self._h, what, level))

    def strip_peer_pull(self, what, level):
        _check(lib().sb_compositor_strip_peer_pull(self._h, what, level))

    def strip_frame_peer(self, frames):
        """Exchange-halo mode with peer-memory exchange: every stage and exchange of one frame, ONE call; asynchronous."""
        arr, keep = self._srcs(frames)
        _check(lib().sb_compositor_strip_frame_peer(self._h, arr))

    def strip_range(self, rank, world):
        x0, x1 = C.c_int(), C.c_int()
        _check(lib().sb_compositor_strip_range(self._h, rank, world, C.byref(x0), C.byref(x1)))
        return x0.value, x1.value

    def set_stream(self, cuda_stream):
        _check(lib().sb_compositor_set_stream(self._h, C.c_void_p(cuda_stream)))

    def strip_halo_bytes(self, what, level, side):
        a, b = C.c_size_t(), C.c_size_t()
        _check(lib().sb_compositor_strip_halo_bytes(self._h, what, level, side, C.byref(a), C.byref(b)))
        return a.value, b.value

    def strip_pack(self, what, level, side, device_ptr):
        _check(lib().sb_compositor_strip_pack(self._h, what, level, side, C.c_void_p(device_ptr)))

    def strip_unpack(self, what, level, side, device_ptr):
        _check(lib().sb_compositor_strip_unpack(self._h, what, level, side, C.c_void_p(device_ptr)))

    def strip_warp(self, frames):
        arr, keep = self._srcs(frames)
        _check(lib().sb_compositor_strip_warp(self._h, arr))

    def strip_down(self, level):
        _check(lib().sb_compositor_strip_down(self._h, level))

    def strip_band(self, level):
        _check(lib().sb_compositor_strip_band(self._h, level))

    def strip_result(self, rank, world, pano=None, pano_mask=None):
        """-> (strip, strip_mask) host arrays holding this rank's columns of the panorama; with `pano` / `pano_mask` (full
        panorama-sized arrays) the columns are written straight into them and the returned arrays are views."""
        x0, x1 = self.strip_range(rank, world)
        strip = _empty(self.pano_size[1], x1 - x0, self.output_type) if pano is None else pano[:, x0:x1]
        smask = np.empty((self.pano_size[1], x1 - x0), np.uint8) if pano_mask is None else pano_mask[:, x0:x1]
        i0, k0 = _image(strip, output=True)
        i1, k1 = _image(smask, output=True)
        _check(lib().sb_compositor_strip_result(self._h, C.byref(i0), C.byref(i1)))
        return strip, smask
