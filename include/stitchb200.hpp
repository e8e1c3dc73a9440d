// stitchb200.hpp — C++ host side above the C ABI (stitchb200.h): header-only adapter classes with the
// names, argument meaning and error behaviour of the reference's operator interfaces for the per-frame
// compositing path:
//
//   cv::detail::RotationWarper      INC/detail/warpers.hpp:53-72            -> sb200::RotationWarper
//     SphericalWarper / CylindricalWarper / PlaneWarper (warpers.hpp:135-161,330-364)
//   cv::WarperCreator               INC/warpers.hpp:50-167                  -> sb200::WarperCreator
//   cv::detail::ExposureCompensator INC/detail/exposure_compensate.hpp:51-101 -> sb200::ExposureCompensator
//   cv::detail::Blender             INC/detail/blenders.hpp:53-117           -> sb200::Blender / FeatherBlender / MultiBandBlender
//   the frame loop of Stitcher::composePanorama (LIB/src/stitcher.cpp:221-313) -> sb200::Compositor
//
// OpenCV headers are not required: sb200::Mat is a non-owning (or malloc-owning) view with cv::Mat's
// {data, rows, cols, type, step} fields, so `sb200::Mat(m.rows, m.cols, m.type(), m.data, m.step)` wraps
// a cv::Mat without a copy (INTEGRATION.md shows the cv::detail subclasses built on these classes).
// Errors are thrown as sb200::Exception carrying the cv::Exception status code.  There is no CPU
// fallback: without a CUDA device every operation throws CV_GpuApiCallError (-217).
#ifndef STITCHB200_HPP
#define STITCHB200_HPP

#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "stitchb200.h"

namespace sb200 {

struct Point { int x = 0, y = 0; Point() {} Point(int x_, int y_) : x(x_), y(y_) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
    Point tl() const { return Point(x, y); }
    Point br() const { return Point(x + width, y + height); }
};

class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc)
{
    if (rc != SB_OK) throw Exception(rc, sb_last_error());
}

inline size_t elem_size(int type)
{
    static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 0};
    return depth_bytes[type & 7] * (size_t)(((type >> 3) & 63) + 1);
}

// cv::Mat / gpu::GpuMat stand-in.  device < 0: host memory; device >= 0: CUDA device pointer.
class Mat {
public:
    void *data = nullptr;
    int rows = 0, cols = 0, type_ = SB_8UC1;
    size_t step = 0;
    int device = -1;

    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void *d, size_t s = 0, int dev = -1)
        : data(d), rows(r), cols(c), type_(t), step(s ? s : (size_t)c * elem_size(t)), device(dev) {}
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    // Mat::create: (re)allocate host storage when the geometry changes
    void create(int r, int c, int t)
    {
        if (data && r == rows && c == cols && t == type_) return;   // like cv::Mat::create: keep matching storage (also wrapped / device)
        rows = r; cols = c; type_ = t; device = -1;
        step = (size_t)c * elem_size(t);
        owner_.reset(static_cast<unsigned char *>(std::malloc(step * (size_t)(r > 0 ? r : 1))), std::free);
        if (!owner_) throw Exception(SB_ERR_NO_MEM, "Mat::create: out of memory");
        data = owner_.get();
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    template <typename T> T *ptr(int y = 0) { return reinterpret_cast<T *>(static_cast<unsigned char *>(data) + (size_t)y * step); }
    template <typename T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(static_cast<const unsigned char *>(data) + (size_t)y * step); }
    sb_image c() const
    {
        sb_image i;
        i.data = data; i.rows = rows; i.cols = cols; i.type = type_; i.step = step; i.device = device;
        return i;
    }

private:
    std::shared_ptr<unsigned char> owner_;
};

// ------------------------------------------------------------------------------------ warpers
class RotationWarper {
public:
    virtual ~RotationWarper() { sb_warper_destroy(h_); }
    RotationWarper(const RotationWarper &) = delete;
    RotationWarper &operator=(const RotationWarper &) = delete;

    // warpers.hpp:58 — K, R: row-major 3x3 CV_32F (warpers.cpp:52-54)
    virtual Point2f warpPoint(const Point2f &pt, const float K[9], const float R[9])
    {
        const float p[2] = {pt.x, pt.y};
        float uv[2];
        check(sb_warper_warp_point(h_, p, K, R, uv));
        return Point2f(uv[0], uv[1]);
    }
    // warpers.hpp:60 — returns Rect(dst_tl, dst_br); maps are (height+1) x (width+1) CV_32F
    virtual Rect buildMaps(Size src_size, const float K[9], const float R[9], Mat &xmap, Mat &ymap)
    {
        sb_rect r;
        sb_size s = {src_size.width, src_size.height};
        check(sb_warper_build_maps(h_, s, K, R, nullptr, nullptr, &r));
        xmap.create(r.height + 1, r.width + 1, SB_32FC1);
        ymap.create(r.height + 1, r.width + 1, SB_32FC1);
        sb_image ix = xmap.c(), iy = ymap.c();
        check(sb_warper_build_maps(h_, s, K, R, &ix, &iy, &r));
        return Rect(r.x, r.y, r.width, r.height);
    }
    // warpers.hpp:62 — returns dst_roi.tl(); dst is allocated like dst.create(roi.height + 1, roi.width + 1, src.type())
    virtual Point warp(const Mat &src, const float K[9], const float R[9], int interp_mode, int border_mode, Mat &dst)
    {
        const Rect roi = warpRoi(src.size(), K, R);
        dst.create(roi.height, roi.width, src.type());
        sb_image is = src.c(), id = dst.c();
        sb_point tl;
        check(sb_warper_warp(h_, &is, K, R, interp_mode, border_mode, &id, &tl));
        return Point(tl.x, tl.y);
    }
    // the cached-map video path of the app (APP64:188-198, 752)
    virtual void remap(const Mat &src, int interp_mode, int border_mode, Mat &dst)
    {
        sb_image is = src.c(), id = dst.c();
        check(sb_warper_remap(h_, &is, interp_mode, border_mode, &id));
    }
    // warpers.hpp:65
    virtual void warpBackward(const Mat &src, const float K[9], const float R[9], int interp_mode, int border_mode, Size dst_size, Mat &dst)
    {
        dst.create(dst_size, src.type());
        sb_image is = src.c(), id = dst.c();
        sb_size s = {dst_size.width, dst_size.height};
        check(sb_warper_warp_backward(h_, &is, K, R, interp_mode, border_mode, s, &id));
    }
    // warpers.hpp:68 — Rect(dst_tl, Point(dst_br.x + 1, dst_br.y + 1))
    virtual Rect warpRoi(Size src_size, const float K[9], const float R[9])
    {
        sb_rect r;
        sb_size s = {src_size.width, src_size.height};
        check(sb_warper_warp_roi(h_, s, K, R, &r));
        return Rect(r.x, r.y, r.width, r.height);
    }
    virtual float getScale() const { return sb_warper_get_scale(h_); }     // warpers.hpp:70
    virtual void setScale(float v) { check(sb_warper_set_scale(h_, v)); }  // warpers.hpp:71
    sb_warper *handle() const { return h_; }

protected:
    RotationWarper(int kind, float scale, int device) { check(sb_warper_create(kind, scale, device, &h_)); }
    sb_warper *h_ = nullptr;
};

class PlaneWarper : public RotationWarper {
public:
    explicit PlaneWarper(float scale = 1.f, int device = 0) : RotationWarper(SB_WARP_PLANE, scale, device) {}
    void setTranslation(const float T[3]) { check(sb_warper_set_translation(h_, T)); }   // the T overloads, warpers.cpp:81-137
};
class CylindricalWarper : public RotationWarper {
public:
    explicit CylindricalWarper(float scale, int device = 0) : RotationWarper(SB_WARP_CYLINDRICAL, scale, device) {}
};
class SphericalWarper : public RotationWarper {
public:
    explicit SphericalWarper(float scale, int device = 0) : RotationWarper(SB_WARP_SPHERICAL, scale, device) {}
};

// the remaining projectors of detail/warpers.hpp:190-503: maps are built on the host (libm), the remap runs on the device
#define SB200_SIMPLE_WARPER(NAME, KIND)                                                                  \
    class NAME : public RotationWarper {                                                                  \
    public:                                                                                               \
        explicit NAME(float scale, int device = 0) : RotationWarper(KIND, scale, device) {}               \
    }
#define SB200_AB_WARPER(NAME, KIND)                                                                      \
    class NAME : public RotationWarper {                                                                  \
    public:                                                                                               \
        explicit NAME(float scale, float A = 1.f, float B = 1.f, int device = 0) : RotationWarper(KIND, scale, device) \
        {                                                                                                 \
            check(sb_warper_set_ab(h_, A, B));                                                            \
        }                                                                                                 \
    }
SB200_SIMPLE_WARPER(FisheyeWarper, SB_WARP_FISHEYE);
SB200_SIMPLE_WARPER(StereographicWarper, SB_WARP_STEREOGRAPHIC);
SB200_AB_WARPER(CompressedRectilinearWarper, SB_WARP_COMPRESSED_RECTILINEAR);
SB200_AB_WARPER(CompressedRectilinearPortraitWarper, SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT);
SB200_AB_WARPER(PaniniWarper, SB_WARP_PANINI);
SB200_AB_WARPER(PaniniPortraitWarper, SB_WARP_PANINI_PORTRAIT);
SB200_SIMPLE_WARPER(MercatorWarper, SB_WARP_MERCATOR);
SB200_SIMPLE_WARPER(TransverseMercatorWarper, SB_WARP_TRANSVERSE_MERCATOR);
SB200_SIMPLE_WARPER(SphericalPortraitWarper, SB_WARP_SPHERICAL_PORTRAIT);
SB200_SIMPLE_WARPER(CylindricalPortraitWarper, SB_WARP_CYLINDRICAL_PORTRAIT);
SB200_SIMPLE_WARPER(PlanePortraitWarper, SB_WARP_PLANE_PORTRAIT);
#undef SB200_SIMPLE_WARPER
#undef SB200_AB_WARPER

// cv::WarperCreator and its subclasses (INC/warpers.hpp:50-83)
struct WarperCreator {
    virtual ~WarperCreator() {}
    virtual std::unique_ptr<RotationWarper> create(float scale) const = 0;
};
struct PlaneWarperCreator : WarperCreator {
    std::unique_ptr<RotationWarper> create(float scale) const override { return std::unique_ptr<RotationWarper>(new PlaneWarper(scale)); }
};
struct CylindricalWarperCreator : WarperCreator {
    std::unique_ptr<RotationWarper> create(float scale) const override { return std::unique_ptr<RotationWarper>(new CylindricalWarper(scale)); }
};
struct SphericalWarperCreator : WarperCreator {
    std::unique_ptr<RotationWarper> create(float scale) const override { return std::unique_ptr<RotationWarper>(new SphericalWarper(scale)); }
};

// ------------------------------------------------------------------------------------ exposure
class ExposureCompensator {
public:
    enum { NO = SB_COMP_NO, GAIN = SB_COMP_GAIN, GAIN_BLOCKS = SB_COMP_GAIN_BLOCKS };   // exposure_compensate.hpp:56
    virtual ~ExposureCompensator() { sb_comp_destroy(h_); }
    ExposureCompensator(const ExposureCompensator &) = delete;
    ExposureCompensator &operator=(const ExposureCompensator &) = delete;
    // exposure_compensate.cpp:51-61; an unknown type throws CV_StsBadArg
    static std::unique_ptr<ExposureCompensator> createDefault(int type, int device = 0)
    {
        return std::unique_ptr<ExposureCompensator>(new ExposureCompensator(type, device));
    }
    // feed() (gain estimation, exposure_compensate.cpp:64-147,165-222) is calibration, once per sequence: either hand
    // its result in with setGains / setGainMaps, or call feed — the overlap statistics are reduced on the device.
    void feed(const std::vector<Point> &corners, const std::vector<Mat> &images, const std::vector<Mat> &masks)
    {
        if (corners.size() != images.size() || images.size() != masks.size())
            throw Exception(SB_ERR_ASSERT, "corners.size() == images.size() && images.size() == masks.size()");
        std::vector<sb_point> c;
        std::vector<sb_image> im, mk;
        for (size_t i = 0; i < images.size(); ++i) { c.push_back(sb_point{corners[i].x, corners[i].y}); im.push_back(images[i].c()); mk.push_back(masks[i].c()); }
        check(sb_comp_feed(h_, c.data(), im.data(), mk.data(), (int)images.size()));
        n_ = (int)images.size();
    }
    void setGains(const std::vector<double> &g) { check(sb_comp_set_gains(h_, g.data(), (int)g.size())); n_ = (int)g.size(); }
    std::vector<double> gains() const
    {
        std::vector<double> g((size_t)n_);
        check(sb_comp_get_gains(h_, g.data(), n_));
        return g;
    }
    void setGainMaps(const std::vector<Mat> &maps)
    {
        std::vector<sb_image> v;
        for (const Mat &m : maps) v.push_back(m.c());
        check(sb_comp_set_gain_maps(h_, v.data(), (int)v.size()));
    }
    // exposure_compensate.hpp:63 — in place on CV_8UC3
    virtual void apply(int index, Point corner, Mat &image, const Mat &mask)
    {
        sb_image ii = image.c(), im = mask.c();
        sb_point c = {corner.x, corner.y};
        check(sb_comp_apply(h_, index, c, &ii, mask.empty() ? nullptr : &im));
    }

protected:
    ExposureCompensator(int kind, int device) { check(sb_comp_create(kind, device, &h_)); }
    sb_comp *h_ = nullptr;
    int n_ = 0;
};
class NoExposureCompensator : public ExposureCompensator {
public:
    explicit NoExposureCompensator(int device = 0) : ExposureCompensator(SB_COMP_NO, device) {}
};
class GainCompensator : public ExposureCompensator {
public:
    explicit GainCompensator(int device = 0) : ExposureCompensator(SB_COMP_GAIN, device) {}
};
class BlocksGainCompensator : public ExposureCompensator {
public:
    // BlocksGainCompensator(int bl_width = 32, int bl_height = 32) (exposure_compensate.hpp:92)
    explicit BlocksGainCompensator(int bl_width = 32, int bl_height = 32, int device = 0) : ExposureCompensator(SB_COMP_GAIN_BLOCKS, device)
    {
        check(sb_comp_set_block_size(h_, bl_width, bl_height));
    }
};

// ------------------------------------------------------------------------------------ blenders
class Blender {
public:
    enum { NO = SB_BLEND_NO, FEATHER = SB_BLEND_FEATHER, MULTI_BAND = SB_BLEND_MULTI_BAND };   // blenders.hpp:58
    explicit Blender(int device = 0) : Blender(SB_BLEND_NO, 5, SB_32F, 0.02f, device) {}
    virtual ~Blender() { sb_blender_destroy(h_); }
    Blender(const Blender &) = delete;
    Blender &operator=(const Blender &) = delete;
    // blenders.cpp:52-62; an unknown type throws CV_StsBadArg
    static std::unique_ptr<Blender> createDefault(int type, bool /*try_gpu*/ = false, int device = 0)
    {
        return std::unique_ptr<Blender>(new Blender(type, 5, SB_32F, 0.02f, device));
    }
    // blenders.cpp:65-68
    void prepare(const std::vector<Point> &corners, const std::vector<Size> &sizes)
    {
        if (corners.size() != sizes.size()) throw Exception(SB_ERR_ASSERT, "sizes.size() == corners.size()");   // util.cpp:129
        std::vector<sb_point> p;
        std::vector<sb_size> s;
        for (const Point &c : corners) p.push_back(sb_point{c.x, c.y});
        for (const Size &z : sizes) s.push_back(sb_size{z.width, z.height});
        check(sb_blender_prepare(h_, p.data(), s.data(), (int)p.size()));
    }
    virtual void prepare(Rect dst_roi) { check(sb_blender_prepare_rect(h_, sb_rect{dst_roi.x, dst_roi.y, dst_roi.width, dst_roi.height})); }
    // blenders.hpp:65 — img CV_16SC3 (multi-band also CV_8UC3), mask CV_8U
    virtual void feed(const Mat &img, const Mat &mask, Point tl)
    {
        sb_image ii = img.c(), im = mask.c();
        check(sb_blender_feed(h_, &ii, &im, sb_point{tl.x, tl.y}));
    }
    // blenders.hpp:66 — like the reference, prepare() must be called again after blend()
    virtual void blend(Mat &dst, Mat &dst_mask)
    {
        sb_size s;
        check(sb_blender_result_size(h_, &s));
        dst.create(s.height, s.width, SB_16SC3);
        dst_mask.create(s.height, s.width, SB_8UC1);
        sb_image id = dst.c(), im = dst_mask.c();
        check(sb_blender_blend(h_, &id, &im));
    }
    sb_blender *handle() const { return h_; }

protected:
    Blender(int kind, int num_bands, int weight_type, float sharpness, int device)
    {
        check(sb_blender_create(kind, num_bands, weight_type, sharpness, device, &h_));
    }
    sb_blender *h_ = nullptr;
};
class FeatherBlender : public Blender {
public:
    explicit FeatherBlender(float sharpness = 0.02f, int device = 0) : Blender(SB_BLEND_FEATHER, 5, SB_32F, sharpness, device) {}
    float sharpness() const { return sb_blender_sharpness(h_); }
    void setSharpness(float v) { check(sb_blender_set_sharpness(h_, v)); }
    // blenders.hpp:80-81: weight maps for a fixed set of source images by their masks and top-left corners
    Rect createWeightMaps(const std::vector<Mat> &masks, const std::vector<Point> &corners, std::vector<Mat> &weight_maps)
    {
        if (masks.size() != corners.size() || masks.empty()) throw Exception(SB_ERR_ASSERT, "masks.size() == corners.size()");
        std::vector<sb_image> m, w;
        std::vector<sb_point> c;
        weight_maps.resize(masks.size());
        for (size_t i = 0; i < masks.size(); ++i) {
            weight_maps[i].create(masks[i].rows, masks[i].cols, SB_32FC1);
            m.push_back(masks[i].c()); w.push_back(weight_maps[i].c()); c.push_back(sb_point{corners[i].x, corners[i].y});
        }
        sb_rect r;
        check(sb_blender_create_weight_maps(h_, m.data(), c.data(), (int)masks.size(), w.data(), &r));
        return Rect(r.x, r.y, r.width, r.height);
    }
};
class MultiBandBlender : public Blender {
public:
    explicit MultiBandBlender(int /*try_gpu*/ = false, int num_bands = 5, int weight_type = SB_32F, int device = 0)
        : Blender(SB_BLEND_MULTI_BAND, num_bands, weight_type, 0.02f, device) {}
    int numBands() const { return sb_blender_num_bands(h_); }
    void setNumBands(int v) { check(sb_blender_set_num_bands(h_, v)); }
};

inline void normalizeUsingWeightMap(const Mat &weight, Mat &src, int device = 0)   // blenders.hpp:122
{
    sb_image w = weight.c(), s = src.c();
    check(sb_normalize_using_weight_map(&w, &s, device));
}
inline void createWeightMap(const Mat &mask, float sharpness, Mat &weight, int device = 0)   // blenders.hpp:124
{
    weight.create(mask.rows, mask.cols, SB_32FC1);
    sb_image m = mask.c(), w = weight.c();
    check(sb_create_weight_map(&m, sharpness, &w, device));
}
inline void createLaplacePyr(const Mat &img, int num_levels, std::vector<Mat> &pyr, int device = 0)   // blenders.hpp:126
{
    pyr.resize((size_t)num_levels + 1);
    std::vector<sb_image> v;
    int r = img.rows, c = img.cols;
    for (int l = 0; l <= num_levels; ++l) {
        pyr[(size_t)l].create(r, c, SB_16SC3);
        v.push_back(pyr[(size_t)l].c());
        r = (r + 1) / 2; c = (c + 1) / 2;
    }
    sb_image i = img.c();
    check(sb_create_laplace_pyr(&i, num_levels, v.data(), device));
}
inline void restoreImageFromLaplacePyr(std::vector<Mat> &pyr, int device = 0)   // blenders.hpp:130
{
    std::vector<sb_image> v;
    for (Mat &m : pyr) v.push_back(m.c());
    check(sb_restore_image_from_laplace_pyr(v.data(), (int)v.size(), device));
}

// ------------------------------------------------------------------------------------ frame loop
// Stitcher::composePanorama's per-image loop (stitcher.cpp:221-313) / the app's StitchingAll
// (APP64:724-770) with calibration fixed: one call per frame set.
class Compositor {
public:
    struct Config {
        Size src_size;
        int warper_kind = SB_WARP_SPHERICAL;
        float warper_scale = 1.f;
        std::vector<float> K, R;              // n x 9 each, row-major
        int blender_kind = SB_BLEND_MULTI_BAND;
        int num_bands = 5;
        int weight_type = SB_32F;
        float sharpness = 0.02f;
        std::vector<double> gains;            // empty: no exposure compensation
        std::vector<Mat> gain_maps;           // BlocksGainCompensator: one CV_32FC1 block gain map per camera (instead of gains)
        std::vector<Mat> seam_masks;          // empty: none
        int output_type = SB_8UC3;
    };
    explicit Compositor(const Config &cfg, int device = 0)
    {
        if (cfg.K.size() != cfg.R.size() || cfg.K.size() % 9 != 0 || cfg.K.empty())
            throw Exception(SB_ERR_ASSERT, "K and R must hold n x 9 floats");
        sb_compositor_config c;
        std::memset(&c, 0, sizeof c);
        c.n_cameras = (int)(cfg.K.size() / 9);
        c.src_size = sb_size{cfg.src_size.width, cfg.src_size.height};
        c.warper_kind = cfg.warper_kind; c.warper_scale = cfg.warper_scale;
        c.K = cfg.K.data(); c.R = cfg.R.data();
        c.blender_kind = cfg.blender_kind; c.num_bands = cfg.num_bands; c.weight_type = cfg.weight_type; c.sharpness = cfg.sharpness;
        std::vector<sb_image> gm;
        for (const Mat &m : cfg.gain_maps) gm.push_back(m.c());
        c.comp_kind = !cfg.gains.empty() ? SB_COMP_GAIN : !gm.empty() ? SB_COMP_GAIN_BLOCKS : SB_COMP_NO;
        c.gains = cfg.gains.empty() ? nullptr : cfg.gains.data();
        c.gain_maps = gm.empty() ? nullptr : gm.data();
        std::vector<sb_image> sm;
        for (const Mat &m : cfg.seam_masks) sm.push_back(m.c());
        c.seam_masks = sm.empty() ? nullptr : sm.data();
        c.output_type = cfg.output_type;
        n_ = c.n_cameras;
        out_type_ = cfg.output_type;
        check(sb_compositor_create(&c, device, &h_));
    }
    ~Compositor() { sb_compositor_destroy(h_); }
    Compositor(const Compositor &) = delete;
    Compositor &operator=(const Compositor &) = delete;
    Size panoSize() const
    {
        sb_size s;
        check(sb_compositor_pano_size(h_, &s));
        return Size(s.width, s.height);
    }
    Rect cameraRoi(int i) const
    {
        sb_rect r;
        check(sb_compositor_camera_roi(h_, i, &r));
        return Rect(r.x, r.y, r.width, r.height);
    }
    void compose(const std::vector<Mat> &frames, Mat &pano, Mat &pano_mask)
    {
        if ((int)frames.size() != n_) throw Exception(SB_ERR_ASSERT, "one frame per camera expected");
        const Size ps = panoSize();
        pano.create(ps.height, ps.width, out_type_);
        pano_mask.create(ps.height, ps.width, SB_8UC1);
        std::vector<sb_image> v;
        for (const Mat &m : frames) v.push_back(m.c());
        sb_image ip = pano.c(), im = pano_mask.c();
        check(sb_compositor_compose(h_, v.data(), &ip, &im));
    }
    sb_compositor *handle() const { return h_; }

private:
    sb_compositor *h_ = nullptr;
    int n_ = 0, out_type_ = SB_8UC3;
};

}  // namespace sb200
#endif
