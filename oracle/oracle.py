"""ctypes binding of the CPU oracle (oracle/libstitch_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under stitchingvideo_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstitch_oracle.so")

CV_8U, CV_16S, CV_32F = 0, 3, 5
CV_8UC1, CV_8UC3, CV_16SC1, CV_16SC3, CV_32FC1 = 0, 16, 3, 19, 5
INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4
WARP_PLANE, WARP_CYLINDRICAL, WARP_SPHERICAL = 0, 1, 2
BLEND_NO, BLEND_FEATHER, BLEND_MULTI_BAND = 0, 1, 2


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("so_prims.c", "so_stitch.c", "stitch_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libstitch_oracle.so"])
    return _LIB_PATH


class SoMat(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("type", C.c_int),
                ("step", C.c_size_t)]


class SoProjector(C.Structure):
    _fields_ = [("kind", C.c_int), ("scale", C.c_float), ("k", C.c_float * 9), ("rinv", C.c_float * 9),
                ("r_kinv", C.c_float * 9), ("k_rinv", C.c_float * 9), ("t", C.c_float * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.so_cvround.restype = C.c_int
        L.so_cvround.argtypes = [C.c_float]
        L.so_trunc_short.restype = C.c_short
        L.so_trunc_short.argtypes = [C.c_float]
        L.so_sinf.restype = C.c_float
        L.so_sinf.argtypes = [C.c_float]
        L.so_cosf.restype = C.c_float
        L.so_cosf.argtypes = [C.c_float]
        L.so_border_interpolate.restype = C.c_int
        L.so_version.restype = C.c_char_p
        L.so_blender_create.restype = C.c_void_p
        L.so_blender_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        for n in ("so_blender_destroy", "so_blender_prepare", "so_blender_prepare_rect", "so_blender_feed",
                  "so_blender_result_size", "so_blender_num_bands_effective", "so_blender_blend"):
            getattr(L, n).argtypes = None
        L.so_blender_destroy.argtypes = [C.c_void_p]
        L.so_blender_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.so_blender_prepare_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.so_blender_feed.argtypes = [C.c_void_p, C.POINTER(SoMat), C.POINTER(SoMat), C.c_int, C.c_int]
        L.so_blender_result_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.so_blender_num_bands_effective.argtypes = [C.c_void_p]
        L.so_blender_blend.argtypes = [C.c_void_p, C.POINTER(SoMat), C.POINTER(SoMat)]
        L.so_gain_apply.argtypes = [C.POINTER(SoMat), C.c_double]
        L.so_scale_8u.argtypes = [C.POINTER(SoMat), C.c_double]
        L.so_create_weight_map.argtypes = [C.POINTER(SoMat), C.c_float, C.POINTER(SoMat)]
        _lib = L
    return _lib


_NP2CV = {np.dtype(np.uint8): CV_8U, np.dtype(np.int16): CV_16S, np.dtype(np.float32): CV_32F}


def mat(a):
    """Wrap a C-contiguous-rows numpy array (H,W) or (H,W,C) as so_mat (no copy)."""
    if a.dtype == np.uint16:                                # CV_16UC1: the fractional part of a fixed-point map pair
        assert a.ndim == 2 and a.strides[-1] == 2
        m = SoMat(a.ctypes.data, a.shape[0], a.shape[1], 2, a.strides[0])
        m._keep = a
        return m
    assert a.dtype in _NP2CV, a.dtype
    assert a.ndim in (2, 3)
    cn = 1 if a.ndim == 2 else a.shape[2]
    assert a.strides[-1] == a.itemsize and (a.ndim == 2 or a.strides[1] == a.itemsize * cn)
    m = SoMat(a.ctypes.data, a.shape[0], a.shape[1], _NP2CV[a.dtype] + ((cn - 1) << 3), a.strides[0])
    m._keep = a
    return m


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("oracle %s failed rc=%d" % (what, rc))


# ------------------------------------------------------------------ scalar helpers
def sinf(x):
    L = lib()
    return np.array([L.so_sinf(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32).reshape(np.shape(x))


def cosf(x):
    L = lib()
    return np.array([L.so_cosf(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32).reshape(np.shape(x))


# ------------------------------------------------------------------ warpers
class Warper:
    """RotationWarper-shaped wrapper (warpers.hpp:53-72) over the oracle."""

    def __init__(self, kind, scale):
        self.kind = {"plane": WARP_PLANE, "cylindrical": WARP_CYLINDRICAL, "spherical": WARP_SPHERICAL}.get(kind, kind)
        self.scale = float(scale)
        self.p = SoProjector()

    def _set(self, K, R, T=None):
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        Tp = None if T is None else np.ascontiguousarray(T, np.float32).reshape(3).ctypes.data_as(C.c_void_p)
        lib().so_projector_set(C.byref(self.p), self.kind, C.c_float(self.scale), K.ctypes.data_as(C.c_void_p),
                               R.ctypes.data_as(C.c_void_p), Tp)

    def projector(self, K, R, T=None):
        self._set(K, R, T)
        return {n: np.array(getattr(self.p, n), np.float32) for n in ("k", "rinv", "r_kinv", "k_rinv", "t")}

    def warp_point(self, pt, K, R, T=None):
        self._set(K, R, T)
        u, v = C.c_float(), C.c_float()
        lib().so_map_forward(C.byref(self.p), C.c_float(pt[0]), C.c_float(pt[1]), C.byref(u), C.byref(v))
        return u.value, v.value

    def _roi(self, src_size):
        tl, br = (C.c_int * 2)(), (C.c_int * 2)()
        lib().so_detect_result_roi(C.byref(self.p), int(src_size[0]), int(src_size[1]), tl, br)
        return (tl[0], tl[1]), (br[0], br[1])

    def warp_roi(self, src_size, K, R, T=None):
        """Rect(dst_tl, Point(dst_br.x+1, dst_br.y+1)) -> (x, y, w, h)  (warpers_inl.hpp:131-139)."""
        self._set(K, R, T)
        tl, br = self._roi(src_size)
        return (tl[0], tl[1], br[0] + 1 - tl[0], br[1] + 1 - tl[1])

    def build_maps(self, src_size, K, R, T=None):
        """-> (Rect(tl, br) as (x, y, w, h) with w = br.x - tl.x, xmap, ymap); maps are (h+1, w+1)."""
        self._set(K, R, T)
        tl, br = self._roi(src_size)
        h, w = br[1] - tl[1] + 1, br[0] - tl[0] + 1
        xmap = np.empty((h, w), np.float32)
        ymap = np.empty((h, w), np.float32)
        mx, my = mat(xmap), mat(ymap)
        lib().so_build_maps(C.byref(self.p), (C.c_int * 2)(*tl), (C.c_int * 2)(*br), C.byref(mx), C.byref(my))
        return (tl[0], tl[1], br[0] - tl[0], br[1] - tl[1]), xmap, ymap

    def warp(self, src, K, R, interp=INTER_LINEAR, border=BORDER_REFLECT, T=None):
        """-> (tl, dst)  (warpers_inl.hpp:88-99)."""
        roi, xmap, ymap = self.build_maps((src.shape[1], src.shape[0]), K, R, T)
        return (roi[0], roi[1]), remap(src, xmap, ymap, interp, border)


def remap(src, xmap, ymap, interp=INTER_LINEAR, border=BORDER_REFLECT, border_value=(0, 0, 0, 0)):
    src = np.ascontiguousarray(src)
    dst = np.empty(xmap.shape + src.shape[2:], np.uint8)
    bv = (C.c_uint8 * 4)(*border_value)
    ms, md, mx, my = mat(src), mat(dst), mat(np.ascontiguousarray(xmap)), mat(np.ascontiguousarray(ymap))
    _chk(lib().so_remap(C.byref(ms), C.byref(md), C.byref(mx), C.byref(my), interp, border, bv), "remap")
    return dst


def convert_maps(xmap, ymap, nn_interpolation=False):
    """cv::convertMaps(xmap, ymap, CV_16SC2) -> (map1 int16 HxWx2, map2 uint16 HxW | None)."""
    xmap, ymap = np.ascontiguousarray(xmap, np.float32), np.ascontiguousarray(ymap, np.float32)
    m1 = np.empty(xmap.shape + (2,), np.int16)
    m2 = None if nn_interpolation else np.empty(xmap.shape, np.uint16)
    mx, my, a1 = mat(xmap), mat(ymap), mat(m1)
    a2 = None if m2 is None else mat(m2)
    _chk(lib().so_convert_maps(C.byref(mx), C.byref(my), C.byref(a1), C.byref(a2) if a2 is not None else None, int(nn_interpolation)), "convertMaps")
    return m1, m2


def remap_fixed(src, map1, map2, interp=INTER_LINEAR, border=BORDER_REFLECT, border_value=(0, 0, 0, 0)):
    """cv::remap with a CV_16SC2 / CV_16UC1 fixed-point map pair (map2 may be None)."""
    src = np.ascontiguousarray(src)
    dst = np.empty(map1.shape[:2] + src.shape[2:], np.uint8)
    bv = (C.c_uint8 * 4)(*border_value)
    ms, md, m1 = mat(src), mat(dst), mat(np.ascontiguousarray(map1))
    m2 = None if map2 is None else mat(np.ascontiguousarray(map2))
    _chk(lib().so_remap(C.byref(ms), C.byref(md), C.byref(m1), C.byref(m2) if m2 is not None else None, interp, border, bv), "remap")
    return dst


def copy_make_border(src, top, bottom, left, right, border):
    src = np.ascontiguousarray(src)
    dst = np.empty((src.shape[0] + top + bottom, src.shape[1] + left + right) + src.shape[2:], src.dtype)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_copy_make_border(C.byref(ms), C.byref(md), top, bottom, left, right, border), "copyMakeBorder")
    return dst


def pyr_down(src):
    src = np.ascontiguousarray(src)
    dst = np.empty(((src.shape[0] + 1) // 2, (src.shape[1] + 1) // 2) + src.shape[2:], src.dtype)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_pyr_down(C.byref(ms), C.byref(md)), "pyrDown")
    return dst


def pyr_up(src):
    src = np.ascontiguousarray(src)
    dst = np.empty((src.shape[0] * 2, src.shape[1] * 2) + src.shape[2:], src.dtype)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_pyr_up(C.byref(ms), C.byref(md)), "pyrUp")
    return dst


def gain_apply(img, gain):
    out = np.ascontiguousarray(img).copy()
    m = mat(out)
    _chk(lib().so_gain_apply(C.byref(m), C.c_double(gain)), "gain_apply")
    return out


def blocks_gain_apply(img, gain_map):
    out = np.ascontiguousarray(img).copy()
    m, g = mat(out), mat(np.ascontiguousarray(gain_map, np.float32))
    _chk(lib().so_blocks_gain_apply(C.byref(m), C.byref(g)), "blocks_gain_apply")
    return out


def resize_linear(src, dsize_wh):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.empty((dsize_wh[1], dsize_wh[0]), np.float32)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_resize_linear_32f(C.byref(ms), C.byref(md)), "resize")
    return dst



# ---- once-per-calibration steps (so_calib.c) -------------------------------------------------
def _mat_array(arrs):
    ms = [mat(a) for a in arrs]
    arr = (SoMat * len(ms))(*ms)
    arr._keep = ms
    return arr


def gain_overlap_stats(corners, images, masks, mask_vals=None):
    """exposure_compensate.cpp:93-126 -> (N int32 n x n, I float64 n x n)"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    vals = (C.c_ubyte * n)(*([255] * n if mask_vals is None else mask_vals))
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    N, I = np.zeros((n, n), np.int32), np.zeros((n, n), np.float64)
    lib().so_gain_overlap_stats(n, cxy.ctypes.data_as(C.c_void_p), _mat_array(images), _mat_array(masks), vals,
                                N.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p))
    return N, I


def gain_solve(N, I):
    """exposure_compensate.cpp:128-144 on dense N (int32 n x n) and I (float64 n x n) -> gains"""
    N, I = np.ascontiguousarray(N, np.int32), np.ascontiguousarray(I, np.float64)
    g = np.zeros(N.shape[0], np.float64)
    _chk(lib().so_gain_solve(N.shape[0], N.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p)), "gain_solve")
    return g


def gain_feed(corners, images, masks, mask_vals=None):
    """GainCompensator::feed (exposure_compensate.cpp:76-147) -> gains (float64)"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    vals = (C.c_ubyte * n)(*([255] * n if mask_vals is None else mask_vals))
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    gains = np.zeros(n, np.float64)
    _chk(lib().so_gain_feed(n, cxy.ctypes.data_as(C.c_void_p), _mat_array(images), _mat_array(masks), vals,
                            gains.ctypes.data_as(C.c_void_p)), "gain_feed")
    return gains


def blocks_gain_feed(corners, images, masks, bl_width=32, bl_height=32):
    """BlocksGainCompensator::feed (exposure_compensate.cpp:165-222) -> list of float32 block gain maps"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    vals = (C.c_ubyte * n)(*([255] * n))
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    maps = [np.zeros(((a.shape[0] + bl_height - 1) // bl_height, (a.shape[1] + bl_width - 1) // bl_width), np.float32) for a in images]
    _chk(lib().so_blocks_gain_feed(n, cxy.ctypes.data_as(C.c_void_p), _mat_array(images), _mat_array(masks), vals,
                                   bl_width, bl_height, _mat_array(maps)), "blocks_gain_feed")
    return maps


def feather_create_weight_maps(masks, corners, sharpness=0.02):
    """FeatherBlender::createWeightMaps (blenders.cpp:158-186) -> (dst_roi (x, y, w, h), list of float32 weight maps)"""
    masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
    maps = [np.zeros(m.shape, np.float32) for m in masks]
    cxy = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    roi = (C.c_int * 4)()
    _chk(lib().so_feather_create_weight_maps(len(masks), _mat_array(masks), cxy.ctypes.data_as(C.c_void_p), C.c_float(sharpness),
                                             _mat_array(maps), roi), "createWeightMaps")
    return tuple(roi), maps


def sep_filter3(src, k0=0.5, k1=0.25):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.empty_like(src)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_sep_filter3_f32(C.byref(ms), C.byref(md), C.c_float(k0), C.c_float(k1)), "sepFilter2D")
    return dst


def dilate3x3(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_dilate3x3_8u(C.byref(ms), C.byref(md)), "dilate")
    return dst


def resize_linear_8u(src, dsize_wh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dsize_wh[1], dsize_wh[0]), np.uint8)
    ms, md = mat(src), mat(dst)
    _chk(lib().so_resize_linear_8u(C.byref(ms), C.byref(md)), "resize 8u")
    return dst


def refine_seam_mask(seam_mask, mask_warped):
    """stitcher.cpp:291-294"""
    seam_mask, mask_warped = np.ascontiguousarray(seam_mask, np.uint8), np.ascontiguousarray(mask_warped, np.uint8)
    dst = np.empty_like(mask_warped)
    a, b, d = mat(seam_mask), mat(mask_warped), mat(dst)
    _chk(lib().so_refine_seam_mask(C.byref(a), C.byref(b), C.byref(d)), "refine_seam_mask")
    return dst

def distance_l1(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    dst = np.empty(mask.shape, np.float32)
    ms, md = mat(mask), mat(dst)
    _chk(lib().so_distance_l1_3x3(C.byref(ms), C.byref(md)), "distanceTransform")
    return dst


def create_weight_map(mask, sharpness):
    mask = np.ascontiguousarray(mask, np.uint8)
    dst = np.empty(mask.shape, np.float32)
    ms, md = mat(mask), mat(dst)
    _chk(lib().so_create_weight_map(C.byref(ms), C.c_float(sharpness), C.byref(md)), "createWeightMap")
    return dst


def create_laplace_pyr(img, num_levels):
    img = np.ascontiguousarray(img)
    pyr, r, c = [], img.shape[0], img.shape[1]
    for _ in range(num_levels + 1):
        pyr.append(np.zeros((r, c, 3), np.int16))
        r, c = (r + 1) // 2, (c + 1) // 2
    mats = (SoMat * (num_levels + 1))(*[mat(p) for p in pyr])
    mi = mat(img)
    _chk(lib().so_create_laplace_pyr(C.byref(mi), num_levels, mats), "createLaplacePyr")
    return pyr


def restore_from_laplace_pyr(pyr):
    pyr = [np.ascontiguousarray(p).copy() for p in pyr]
    mats = (SoMat * len(pyr))(*[mat(p) for p in pyr])
    _chk(lib().so_restore_image_from_laplace_pyr(mats, len(pyr)), "restoreImageFromLaplacePyr")
    return pyr[0]


def convert_16s_8u(a):
    return np.clip(a, 0, 255).astype(np.uint8)


def result_roi(corners, sizes):
    c = np.ascontiguousarray(corners, np.int32)
    s = np.ascontiguousarray(sizes, np.int32)
    roi = (C.c_int * 4)()
    lib().so_result_roi(c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(c), roi)
    return tuple(roi)


class Blender:
    """detail::Blender-shaped wrapper (blenders.hpp:53-117) over the oracle."""

    def __init__(self, kind=BLEND_MULTI_BAND, num_bands=5, weight_type=CV_32F, sharpness=0.02):
        self.h = lib().so_blender_create(kind, num_bands, weight_type, C.c_float(sharpness))
        if not self.h:
            raise ValueError("unsupported blender configuration")

    def __del__(self):
        if getattr(self, "h", None):
            lib().so_blender_destroy(self.h)
            self.h = None

    def prepare(self, corners, sizes):
        c = np.ascontiguousarray(corners, np.int32)
        s = np.ascontiguousarray(sizes, np.int32)
        _chk(lib().so_blender_prepare(self.h, c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(c)),
             "prepare")

    def prepare_rect(self, x, y, w, h):
        _chk(lib().so_blender_prepare_rect(self.h, x, y, w, h), "prepare")

    def num_bands(self):
        return lib().so_blender_num_bands_effective(self.h)

    def feed(self, img, mask, tl):
        img = np.ascontiguousarray(img)
        mask = np.ascontiguousarray(mask)
        mi, mm = mat(img), mat(mask)
        _chk(lib().so_blender_feed(self.h, C.byref(mi), C.byref(mm), int(tl[0]), int(tl[1])), "feed")

    def blend(self):
        w, h = C.c_int(), C.c_int()
        lib().so_blender_result_size(self.h, C.byref(w), C.byref(h))
        dst = np.empty((h.value, w.value, 3), np.int16)
        dmask = np.empty((h.value, w.value), np.uint8)
        md, mm = mat(dst), mat(dmask)
        _chk(lib().so_blender_blend(self.h, C.byref(md), C.byref(mm)), "blend")
        return dst, dmask
