// oracle/_ref — the reference's OWN sources of the compositing path, compiled where they lie
// (/root/reference/.../src/{blenders,warpers,util}.cpp + the detail/*.hpp headers they include), unmodified,
// against the OpenCV stand-in of include/opencv2 (primitives = the oracle's restatement of OpenCV 2.4.11).
// TEST INFRASTRUCTURE: tests/test_ref_shim.py compares the oracle (oracle/so_stitch.c) with this library.
// exposure_compensate.cpp is built too (feed() of GainCompensator / BlocksGainCompensator and both apply()).
//
// precomp.hpp pulls in every OpenCV module (features2d, calib3d, ...); its include guard is defined here so that
// only the headers the three sources really need are seen.
#define __OPENCV_STITCHING_PRECOMP_H__
#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <queue>
#include <set>
#include <sstream>
#include <utility>
#include <vector>

#include "opencv2/core/core.hpp"
#define private public                                      // read BlocksGainCompensator::gain_maps_ (no accessor in 2.4.11); the sources stay untouched
#include "opencv2/stitching/detail/exposure_compensate.hpp"
#undef private
#include "opencv2/stitching/detail/blenders.hpp"
#include "opencv2/stitching/detail/util.hpp"
#include "opencv2/stitching/detail/warpers.hpp"

#include REF_SRC(blenders.cpp)
#include REF_SRC(util.cpp)
#include REF_SRC(warpers.cpp)
#include REF_SRC(exposure_compensate.cpp)

// ------------------------------------------------------------------------------------ C API (ctypes)
using cv::Mat;
static Mat wrap(const so_mat *m) { return Mat(m->rows, m->cols, m->type, m->data, m->step); }
static int copy_out(const Mat &src, so_mat *dst)
{
    if (src.rows != dst->rows || src.cols != dst->cols || src.type() != dst->type) return -1;
    for (int y = 0; y < src.rows; ++y)
        std::memcpy((char *)dst->data + (size_t)y * dst->step, src.ptr<unsigned char>(y), (size_t)src.cols * src.elemSize());
    return 0;
}
static float g_a = 1.f, g_b = 1.f;      // (A, B) of the CompressedRectilinear / Panini warpers
static cv::Ptr<cv::detail::RotationWarper> make_warper(int kind, float scale)
{
    switch (kind) {      // kind numbers = SB_WARP_* of include/stitchb200.h
    case 0: return new cv::detail::PlaneWarper(scale);
    case 1: return new cv::detail::CylindricalWarper(scale);
    case 2: return new cv::detail::SphericalWarper(scale);
    case 3: return new cv::detail::FisheyeWarper(scale);
    case 4: return new cv::detail::StereographicWarper(scale);
    case 5: return new cv::detail::CompressedRectilinearWarper(scale, g_a, g_b);
    case 6: return new cv::detail::CompressedRectilinearPortraitWarper(scale, g_a, g_b);
    case 7: return new cv::detail::PaniniWarper(scale, g_a, g_b);
    case 8: return new cv::detail::PaniniPortraitWarper(scale, g_a, g_b);
    case 9: return new cv::detail::MercatorWarper(scale);
    case 10: return new cv::detail::TransverseMercatorWarper(scale);
    case 11: return new cv::detail::SphericalPortraitWarper(scale);
    case 12: return new cv::detail::CylindricalPortraitWarper(scale);
    case 13: return new cv::detail::PlanePortraitWarper(scale);
    }
    CV_Error(CV_StsBadArg, "unknown warper kind");
    return cv::Ptr<cv::detail::RotationWarper>();
}
static Mat mat3(const float *m)
{
    Mat r(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) r.at<float>(i / 3, i % 3) = m[i];
    return r;
}

struct ref_blender {
    cv::Ptr<cv::detail::Blender> b;
    cv::Rect roi;
};

extern "C" {

void ref_set_ab(float a, float b) { g_a = a; g_b = b; }
const char *ref_version(void) { return "reference sources (blenders.cpp, warpers.cpp, util.cpp) on the OpenCV shim"; }

ref_blender *ref_blender_create(int kind, int num_bands, int weight_type, float sharpness)
{
    try {
        ref_blender *h = new ref_blender;
        if (kind == cv::detail::Blender::MULTI_BAND) h->b = new cv::detail::MultiBandBlender(false, num_bands, weight_type);
        else if (kind == cv::detail::Blender::FEATHER) h->b = new cv::detail::FeatherBlender(sharpness);
        else h->b = cv::detail::Blender::createDefault(kind, false);
        return h;
    } catch (const cv::Exception &) { return nullptr; }
}
void ref_blender_destroy(ref_blender *h) { delete h; }
int ref_blender_prepare(ref_blender *h, const int *corners_xy, const int *sizes_wh, int n)
{
    try {
        std::vector<cv::Point> c;
        std::vector<cv::Size> s;
        for (int i = 0; i < n; ++i) { c.push_back(cv::Point(corners_xy[2 * i], corners_xy[2 * i + 1])); s.push_back(cv::Size(sizes_wh[2 * i], sizes_wh[2 * i + 1])); }
        h->roi = cv::detail::resultRoi(c, s);              // util.cpp:127-140
        h->b->prepare(c, s);                               // blenders.cpp:65-68
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
void ref_blender_roi(const ref_blender *h, int roi_xywh[4]) { roi_xywh[0] = h->roi.x; roi_xywh[1] = h->roi.y; roi_xywh[2] = h->roi.width; roi_xywh[3] = h->roi.height; }
int ref_blender_feed(ref_blender *h, const so_mat *img, const so_mat *mask, int tl_x, int tl_y)
{
    try { h->b->feed(wrap(img), wrap(mask), cv::Point(tl_x, tl_y)); return 0; } catch (const cv::Exception &e) { return e.code; }
}
int ref_blender_blend(ref_blender *h, so_mat *dst, so_mat *dst_mask)
{
    try {
        Mat d, m;
        h->b->blend(d, m);
        return copy_out(d, dst) | copy_out(m, dst_mask);
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_create_laplace_pyr(const so_mat *img, int num_levels, so_mat *pyr)
{
    try {
        std::vector<Mat> p;
        cv::detail::createLaplacePyr(wrap(img).clone(), num_levels, p);
        int rc = 0;
        for (int i = 0; i <= num_levels; ++i) rc |= copy_out(p[i], &pyr[i]);
        return rc;
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_restore_image_from_laplace_pyr(so_mat *pyr, int n)
{
    try {
        std::vector<Mat> p;
        for (int i = 0; i < n; ++i) p.push_back(wrap(&pyr[i]).clone());
        cv::detail::restoreImageFromLaplacePyr(p);
        return copy_out(p[0], &pyr[0]);
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_normalize_using_weight_map(const so_mat *weight, so_mat *src)
{
    try { Mat s = wrap(src); cv::detail::normalizeUsingWeightMap(wrap(weight), s); return 0; } catch (const cv::Exception &e) { return e.code; }
}
int ref_create_weight_map(const so_mat *mask, float sharpness, so_mat *weight)
{
    try { Mat w; cv::detail::createWeightMap(wrap(mask), sharpness, w); return copy_out(w, weight); } catch (const cv::Exception &e) { return e.code; }
}

// RotationWarper (warpers.hpp:53-72) through the reference's own RotationWarperBase<P> / projectors
int ref_warp_roi(int kind, float scale, int src_w, int src_h, const float K[9], const float R[9], int roi_xywh[4])
{
    try {
        cv::Rect r = make_warper(kind, scale)->warpRoi(cv::Size(src_w, src_h), mat3(K), mat3(R));
        roi_xywh[0] = r.x; roi_xywh[1] = r.y; roi_xywh[2] = r.width; roi_xywh[3] = r.height;
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_warp_point(int kind, float scale, const float pt[2], const float K[9], const float R[9], float uv[2])
{
    try {
        cv::Point2f p = make_warper(kind, scale)->warpPoint(cv::Point2f(pt[0], pt[1]), mat3(K), mat3(R));
        uv[0] = p.x; uv[1] = p.y;
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
// xmap / ymap: (roi.height + 1) x (roi.width + 1) CV_32F with roi from ref_warp_roi minus one (Rect(tl, br))
int ref_build_maps(int kind, float scale, int src_w, int src_h, const float K[9], const float R[9], int roi_xywh[4], so_mat *xmap, so_mat *ymap)
{
    try {
        Mat xm, ym;
        cv::Rect r = make_warper(kind, scale)->buildMaps(cv::Size(src_w, src_h), mat3(K), mat3(R), xm, ym);
        roi_xywh[0] = r.x; roi_xywh[1] = r.y; roi_xywh[2] = r.width; roi_xywh[3] = r.height;
        return copy_out(xm, xmap) | copy_out(ym, ymap);
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_warp(int kind, float scale, const so_mat *src, const float K[9], const float R[9], int interp, int border, int tl[2], so_mat *dst)
{
    try {
        Mat d;
        cv::Point p = make_warper(kind, scale)->warp(wrap(src), mat3(K), mat3(R), interp, border, d);
        tl[0] = p.x; tl[1] = p.y;
        return copy_out(d, dst);
    } catch (const cv::Exception &e) { return e.code; }
}

// RotationWarperBase<P>::warpBackward (warpers_inl.hpp:102-128): dst is dst_h x dst_w of src's type
int ref_warp_backward(int kind, float scale, const so_mat *src, const float K[9], const float R[9], int interp, int border, int dst_w, int dst_h, so_mat *dst)
{
    try {
        Mat d;
        make_warper(kind, scale)->warpBackward(wrap(src), mat3(K), mat3(R), interp, border, cv::Size(dst_w, dst_h), d);
        return copy_out(d, dst);
    } catch (const cv::Exception &e) { return e.code; }
}
// PlaneWarper's overloads with a translation T (warpers.cpp:81-137)
static Mat mat31(const float *t)
{
    Mat r(3, 1, CV_32F);
    for (int i = 0; i < 3; ++i) r.at<float>(i, 0) = t[i];
    return r;
}
int ref_plane_warp_roi_t(float scale, int src_w, int src_h, const float K[9], const float R[9], const float T[3], int roi_xywh[4])
{
    try {
        cv::detail::PlaneWarper w(scale);
        cv::Rect r = w.warpRoi(cv::Size(src_w, src_h), mat3(K), mat3(R), mat31(T));
        roi_xywh[0] = r.x; roi_xywh[1] = r.y; roi_xywh[2] = r.width; roi_xywh[3] = r.height;
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_plane_warp_point_t(float scale, const float pt[2], const float K[9], const float R[9], const float T[3], float uv[2])
{
    try {
        cv::detail::PlaneWarper w(scale);
        cv::Point2f p = w.warpPoint(cv::Point2f(pt[0], pt[1]), mat3(K), mat3(R), mat31(T));
        uv[0] = p.x; uv[1] = p.y;
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_plane_build_maps_t(float scale, int src_w, int src_h, const float K[9], const float R[9], const float T[3], int roi_xywh[4], so_mat *xmap, so_mat *ymap)
{
    try {
        cv::detail::PlaneWarper w(scale);
        Mat xm, ym;
        cv::Rect r = w.buildMaps(cv::Size(src_w, src_h), mat3(K), mat3(R), mat31(T), xm, ym);
        roi_xywh[0] = r.x; roi_xywh[1] = r.y; roi_xywh[2] = r.width; roi_xywh[3] = r.height;
        return copy_out(xm, xmap) | copy_out(ym, ymap);
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_plane_warp_t(float scale, const so_mat *src, const float K[9], const float R[9], const float T[3], int interp, int border, int tl[2], so_mat *dst)
{
    try {
        cv::detail::PlaneWarper w(scale);
        Mat d;
        cv::Point p = w.warp(wrap(src), mat3(K), mat3(R), mat31(T), interp, border, d);
        tl[0] = p.x; tl[1] = p.y;
        return copy_out(d, dst);
    } catch (const cv::Exception &e) { return e.code; }
}

// FeatherBlender::createWeightMaps (blenders.cpp:158-186); weight_maps[i]: CV_32FC1 of masks[i]'s size
int ref_feather_create_weight_maps(int n, const so_mat *masks, const int *corners_xy, float sharpness, so_mat *weight_maps, int roi_xywh[4])
{
    try {
        std::vector<Mat> mk, wm;
        std::vector<cv::Point> c;
        for (int i = 0; i < n; ++i) { mk.push_back(wrap(&masks[i])); c.push_back(cv::Point(corners_xy[2 * i], corners_xy[2 * i + 1])); }
        cv::detail::FeatherBlender fb(sharpness);
        const cv::Rect r = fb.createWeightMaps(mk, c, wm);
        roi_xywh[0] = r.x; roi_xywh[1] = r.y; roi_xywh[2] = r.width; roi_xywh[3] = r.height;
        int rc = 0;
        for (int i = 0; i < n; ++i) rc |= copy_out(wm[i], &weight_maps[i]);
        return rc;
    } catch (const cv::Exception &e) { return e.code; }
}

// ---- ExposureCompensator (exposure_compensate.cpp): feed / gains / apply through the reference's own classes
static void feed_args(int n, const int *corners_xy, const so_mat *images, const so_mat *masks,
                      std::vector<cv::Point> &c, std::vector<Mat> &im, std::vector<Mat> &mk)
{
    for (int i = 0; i < n; ++i) { c.push_back(cv::Point(corners_xy[2 * i], corners_xy[2 * i + 1])); im.push_back(wrap(&images[i])); mk.push_back(wrap(&masks[i])); }
}
int ref_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, double *gains)
{
    try {
        cv::detail::stitchingLogLevel() = 2;               // silence the LOGLN timing lines
        std::vector<cv::Point> c; std::vector<Mat> im, mk;
        feed_args(n, corners_xy, images, masks, c, im, mk);
        cv::detail::GainCompensator comp;
        static_cast<cv::detail::ExposureCompensator &>(comp).feed(c, im, mk);      // the (corners, images, masks) overload, level 255
        std::vector<double> g = comp.gains();
        for (int i = 0; i < n; ++i) gains[i] = g[i];
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}
// gain_maps[i]: CV_32FC1 of the block grid size; image0_out (may be null): images[0] after apply(0, ...)
int ref_blocks_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, int bl_width, int bl_height,
                         so_mat *gain_maps, so_mat *image0_out)
{
    try {
        cv::detail::stitchingLogLevel() = 2;
        std::vector<cv::Point> c; std::vector<Mat> im, mk;
        feed_args(n, corners_xy, images, masks, c, im, mk);
        cv::detail::BlocksGainCompensator comp(bl_width, bl_height);
        static_cast<cv::detail::ExposureCompensator &>(comp).feed(c, im, mk);
        int rc = 0;
        for (int i = 0; i < n; ++i) rc |= copy_out(comp.gain_maps_[i], &gain_maps[i]);
        if (image0_out) {
            Mat img = im[0].clone();
            comp.apply(0, c[0], img, mk[0]);
            rc |= copy_out(img, image0_out);
        }
        return rc;
    } catch (const cv::Exception &e) { return e.code; }
}
int ref_gain_apply(so_mat *image, double gain)
{
    try {
        cv::detail::GainCompensator comp;
        comp.gains_.create(1, 1);
        comp.gains_(0, 0) = gain;
        Mat img = wrap(image);
        comp.apply(0, cv::Point(0, 0), img, Mat());
        return 0;
    } catch (const cv::Exception &e) { return e.code; }
}

}  // extern "C"
