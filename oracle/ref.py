"""ctypes binding of oracle/_ref/libstitch_ref.so: the reference's OWN blenders.cpp / warpers.cpp / util.cpp
compiled where they lie against the OpenCV stand-in of oracle/ref_shim (primitives = the oracle's restatement of
OpenCV 2.4.11).  TEST INFRASTRUCTURE ONLY, like everything under oracle/."""
import ctypes as C
import os

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libstitch_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        O.lib()                                             # libstitch_oracle.so provides the primitives
        L = C.CDLL(LIB_PATH)
        L.ref_version.restype = C.c_char_p
        L.ref_blender_create.restype = C.c_void_p
        L.ref_blender_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_blender_destroy.argtypes = [C.c_void_p]
        L.ref_blender_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_blender_roi.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_blender_feed.argtypes = [C.c_void_p, C.POINTER(O.SoMat), C.POINTER(O.SoMat), C.c_int, C.c_int]
        L.ref_blender_blend.argtypes = [C.c_void_p, C.POINTER(O.SoMat), C.POINTER(O.SoMat)]
        L.ref_create_weight_map.argtypes = [C.POINTER(O.SoMat), C.c_float, C.POINTER(O.SoMat)]
        L.ref_set_ab.argtypes = [C.c_float, C.c_float]
        L.ref_warp_roi.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_warp_point.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_build_maps.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                     C.POINTER(O.SoMat), C.POINTER(O.SoMat)]
        L.ref_warp.argtypes = [C.c_int, C.c_float, C.POINTER(O.SoMat), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                               C.POINTER(C.c_int), C.POINTER(O.SoMat)]
        L.ref_warp_backward.argtypes = [C.c_int, C.c_float, C.POINTER(O.SoMat), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(O.SoMat)]
        L.ref_plane_warp_roi_t.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_plane_warp_point_t.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_plane_build_maps_t.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                             C.POINTER(O.SoMat), C.POINTER(O.SoMat)]
        L.ref_plane_warp_t.argtypes = [C.c_float, C.POINTER(O.SoMat), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(O.SoMat)]
        L.ref_gain_feed.argtypes = [C.c_int, C.c_void_p, C.POINTER(O.SoMat), C.POINTER(O.SoMat), C.c_void_p]
        L.ref_blocks_gain_feed.argtypes = [C.c_int, C.c_void_p, C.POINTER(O.SoMat), C.POINTER(O.SoMat), C.c_int, C.c_int,
                                           C.POINTER(O.SoMat), C.POINTER(O.SoMat)]
        L.ref_gain_apply.argtypes = [C.POINTER(O.SoMat), C.c_double]
        L.ref_feather_create_weight_maps.argtypes = [C.c_int, C.POINTER(O.SoMat), C.c_void_p, C.c_float, C.POINTER(O.SoMat), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("reference %s failed rc=%d" % (what, rc))


_KIND = {"plane": 0, "cylindrical": 1, "spherical": 2, "fisheye": 3, "stereographic": 4, "compressedRectilinear": 5,
         "compressedRectilinearPortrait": 6, "panini": 7, "paniniPortrait": 8, "mercator": 9, "transverseMercator": 10,
         "sphericalPortrait": 11, "cylindricalPortrait": 12, "planePortrait": 13}


def _f9(m):
    return np.ascontiguousarray(m, np.float32).reshape(9)


class Warper:
    """detail::RotationWarper through the reference's RotationWarperBase<P> (warpers_inl.hpp / warpers.cpp)."""

    def __init__(self, kind, scale, a=1.0, b=1.0):
        self.kind, self.scale, self.a, self.b = _KIND.get(kind, kind), float(scale), float(a), float(b)

    def _ab(self):
        lib().ref_set_ab(C.c_float(self.a), C.c_float(self.b))

    def warp_roi(self, src_size, K, R):
        self._ab()
        K, R, roi = _f9(K), _f9(R), (C.c_int * 4)()
        _chk(lib().ref_warp_roi(self.kind, self.scale, src_size[0], src_size[1], K.ctypes.data, R.ctypes.data, roi), "warpRoi")
        return tuple(roi)

    def warp_point(self, pt, K, R):
        self._ab()
        K, R = _f9(K), _f9(R)
        p, uv = np.asarray(pt, np.float32), np.zeros(2, np.float32)
        _chk(lib().ref_warp_point(self.kind, self.scale, p.ctypes.data, K.ctypes.data, R.ctypes.data, uv.ctypes.data), "warpPoint")
        return float(uv[0]), float(uv[1])

    def build_maps(self, src_size, K, R):
        x, y, w, h = self.warp_roi(src_size, K, R)          # Rect(tl, br + 1): the maps are h x w
        K, R, roi = _f9(K), _f9(R), (C.c_int * 4)()
        xmap, ymap = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
        mx, my = O.mat(xmap), O.mat(ymap)
        _chk(lib().ref_build_maps(self.kind, self.scale, src_size[0], src_size[1], K.ctypes.data, R.ctypes.data, roi,
                                  C.byref(mx), C.byref(my)), "buildMaps")
        return tuple(roi), xmap, ymap

    def warp(self, src, K, R, interp=O.INTER_LINEAR, border=O.BORDER_REFLECT):
        src = np.ascontiguousarray(src)
        x, y, w, h = self.warp_roi((src.shape[1], src.shape[0]), K, R)
        K, R, tl = _f9(K), _f9(R), (C.c_int * 2)()
        dst = np.empty((h, w) + src.shape[2:], np.uint8)
        ms, md = O.mat(src), O.mat(dst)
        _chk(lib().ref_warp(self.kind, self.scale, C.byref(ms), K.ctypes.data, R.ctypes.data, interp, border, tl, C.byref(md)), "warp")
        return (tl[0], tl[1]), dst


    def warp_backward(self, src, K, R, dst_size, interp=O.INTER_LINEAR, border=O.BORDER_REFLECT):
        """RotationWarperBase<P>::warpBackward (warpers_inl.hpp:102-128) -> dst of dst_size (w, h)."""
        self._ab()
        src = np.ascontiguousarray(src)
        K, R = _f9(K), _f9(R)
        dst = np.empty((dst_size[1], dst_size[0]) + src.shape[2:], np.uint8)
        ms, md = O.mat(src), O.mat(dst)
        _chk(lib().ref_warp_backward(self.kind, self.scale, C.byref(ms), K.ctypes.data, R.ctypes.data, interp, border,
                                     dst_size[0], dst_size[1], C.byref(md)), "warpBackward")
        return dst


class PlaneWarperT:
    """detail::PlaneWarper's overloads with a translation T (warpers.cpp:81-137) through the reference's own code."""

    def __init__(self, scale, T):
        self.scale, self.T = float(scale), np.ascontiguousarray(T, np.float32).reshape(3)

    def warp_roi(self, src_size, K, R):
        K, R, roi = _f9(K), _f9(R), (C.c_int * 4)()
        _chk(lib().ref_plane_warp_roi_t(self.scale, src_size[0], src_size[1], K.ctypes.data, R.ctypes.data, self.T.ctypes.data, roi), "warpRoi(T)")
        return tuple(roi)

    def warp_point(self, pt, K, R):
        K, R = _f9(K), _f9(R)
        p, uv = np.asarray(pt, np.float32), np.zeros(2, np.float32)
        _chk(lib().ref_plane_warp_point_t(self.scale, p.ctypes.data, K.ctypes.data, R.ctypes.data, self.T.ctypes.data, uv.ctypes.data), "warpPoint(T)")
        return float(uv[0]), float(uv[1])

    def build_maps(self, src_size, K, R):
        x, y, w, h = self.warp_roi(src_size, K, R)
        K, R, roi = _f9(K), _f9(R), (C.c_int * 4)()
        xmap, ymap = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
        mx, my = O.mat(xmap), O.mat(ymap)
        _chk(lib().ref_plane_build_maps_t(self.scale, src_size[0], src_size[1], K.ctypes.data, R.ctypes.data, self.T.ctypes.data, roi,
                                          C.byref(mx), C.byref(my)), "buildMaps(T)")
        return tuple(roi), xmap, ymap

    def warp(self, src, K, R, interp=O.INTER_LINEAR, border=O.BORDER_REFLECT):
        src = np.ascontiguousarray(src)
        x, y, w, h = self.warp_roi((src.shape[1], src.shape[0]), K, R)
        K, R, tl = _f9(K), _f9(R), (C.c_int * 2)()
        dst = np.empty((h, w) + src.shape[2:], np.uint8)
        ms, md = O.mat(src), O.mat(dst)
        _chk(lib().ref_plane_warp_t(self.scale, C.byref(ms), K.ctypes.data, R.ctypes.data, self.T.ctypes.data, interp, border, tl, C.byref(md)), "warp(T)")
        return (tl[0], tl[1]), dst


class Blender:
    """detail::Blender / FeatherBlender / MultiBandBlender: the reference's blenders.cpp itself."""

    def __init__(self, kind=O.BLEND_MULTI_BAND, num_bands=5, weight_type=O.CV_32F, sharpness=0.02):
        self.kind, self.num_bands = kind, num_bands
        self.h = lib().ref_blender_create(kind, num_bands, weight_type, C.c_float(sharpness))
        if not self.h:
            raise ValueError("unsupported blender configuration")

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_blender_destroy(self.h)
            self.h = None

    def prepare(self, corners, sizes):
        c = np.ascontiguousarray(corners, np.int32)
        s = np.ascontiguousarray(sizes, np.int32)
        _chk(lib().ref_blender_prepare(self.h, c.ctypes.data, s.ctypes.data, len(c)), "prepare")
        roi = (C.c_int * 4)()
        lib().ref_blender_roi(self.h, roi)
        self.roi = tuple(roi)

    def feed(self, img, mask, tl):
        img, mask = np.ascontiguousarray(img), np.ascontiguousarray(mask)
        mi, mm = O.mat(img), O.mat(mask)
        _chk(lib().ref_blender_feed(self.h, C.byref(mi), C.byref(mm), int(tl[0]), int(tl[1])), "feed")

    def blend(self):
        w, h = self.roi[2], self.roi[3]
        dst, dmask = np.empty((h, w, 3), np.int16), np.empty((h, w), np.uint8)
        md, mm = O.mat(dst), O.mat(dmask)
        _chk(lib().ref_blender_blend(self.h, C.byref(md), C.byref(mm)), "blend")
        return dst, dmask


def create_weight_map(mask, sharpness):
    mask = np.ascontiguousarray(mask, np.uint8)
    out = np.empty(mask.shape, np.float32)
    mm, mo = O.mat(mask), O.mat(out)
    _chk(lib().ref_create_weight_map(C.byref(mm), C.c_float(sharpness), C.byref(mo)), "createWeightMap")
    return out


def create_laplace_pyr(img, num_levels):
    img = np.ascontiguousarray(img)
    pyr, r, c = [], img.shape[0], img.shape[1]
    for _ in range(num_levels + 1):
        pyr.append(np.zeros((r, c, 3), np.int16))
        r, c = (r + 1) // 2, (c + 1) // 2
    mats = (O.SoMat * (num_levels + 1))(*[O.mat(p) for p in pyr])
    mi = O.mat(img)
    _chk(lib().ref_create_laplace_pyr(C.byref(mi), num_levels, mats), "createLaplacePyr")
    return pyr


def restore_from_laplace_pyr(pyr):
    pyr = [np.ascontiguousarray(p).copy() for p in pyr]
    mats = (O.SoMat * len(pyr))(*[O.mat(p) for p in pyr])
    _chk(lib().ref_restore_image_from_laplace_pyr(mats, len(pyr)), "restoreImageFromLaplacePyr")
    return pyr[0]


def normalize_using_weight_map(weight, src):
    src = np.ascontiguousarray(src).copy()
    mw, ms = O.mat(np.ascontiguousarray(weight)), O.mat(src)
    _chk(lib().ref_normalize_using_weight_map(C.byref(mw), C.byref(ms)), "normalizeUsingWeightMap")
    return src


# ---- ExposureCompensator through the reference's own exposure_compensate.cpp ---------------------------------------
def gain_feed(corners, images, masks):
    """GainCompensator::feed(corners, images, masks) + gains() (exposure_compensate.cpp:64-162)"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    g = np.zeros(n, np.float64)
    _chk(lib().ref_gain_feed(n, cxy.ctypes.data_as(C.c_void_p), O._mat_array(images), O._mat_array(masks), g.ctypes.data_as(C.c_void_p)), "GainCompensator::feed")
    return g


def blocks_gain_feed(corners, images, masks, bl_width=32, bl_height=32, apply_first=False):
    """BlocksGainCompensator(bl_width, bl_height)::feed -> gain_maps_ (and, optionally, images[0] after apply(0, ...))"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    maps = [np.zeros(((a.shape[0] + bl_height - 1) // bl_height, (a.shape[1] + bl_width - 1) // bl_width), np.float32) for a in images]
    out0 = np.zeros_like(images[0]) if apply_first else None
    m0 = O.mat(out0) if apply_first else None
    _chk(lib().ref_blocks_gain_feed(n, cxy.ctypes.data_as(C.c_void_p), O._mat_array(images), O._mat_array(masks), bl_width, bl_height,
                                    O._mat_array(maps), C.byref(m0) if apply_first else None), "BlocksGainCompensator::feed")
    return (maps, out0) if apply_first else maps


def gain_apply(img, gain):
    """GainCompensator::apply with gains_(0, 0) = gain (exposure_compensate.cpp:150-153)"""
    out = np.ascontiguousarray(img).copy()
    m = O.mat(out)
    _chk(lib().ref_gain_apply(C.byref(m), C.c_double(gain)), "GainCompensator::apply")
    return out


def feather_create_weight_maps(masks, corners, sharpness=0.02):
    """FeatherBlender(sharpness)::createWeightMaps (blenders.cpp:158-186) -> (dst_roi, weight maps)"""
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    maps = [np.zeros(m.shape, np.float32) for m in masks]
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    roi = (C.c_int * 4)()
    _chk(lib().ref_feather_create_weight_maps(len(masks), O._mat_array(masks), cxy.ctypes.data_as(C.c_void_p), C.c_float(sharpness),
                                              O._mat_array(maps), roi), "FeatherBlender::createWeightMaps")
    return tuple(roi), maps
