// Minimal stand-in for the OpenCV 2.4.11 core + imgproc API surface that the reference's OWN stitching
// sources (LIB/src/blenders.cpp, warpers.cpp, util.cpp and their headers) touch on the compositing path.
// TEST INFRASTRUCTURE (oracle/_ref): it lets those sources be compiled where they lie, unmodified, so that
// the reference's own logic (ROI alignment, gaps, feed/accumulate loops, normalisation, mask handling,
// projector maths) runs for real.  Every PRIMITIVE (pyrDown, pyrUp, remap, copyMakeBorder, add/subtract,
// convertTo, distanceTransform, 3x3 inv/gemm ...) is delegated to the oracle's C restatement of OpenCV
// 2.4.11 (oracle/so_prims.c, pinned against cv2 by tests/golden), because OpenCV's sources are not in
// /root/reference.  Written from the published OpenCV 2.4 API documentation; no OpenCV code is copied.
#ifndef REF_SHIM_OPENCV_CORE_HPP
#define REF_SHIM_OPENCV_CORE_HPP

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "stitch_oracle.h"

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_PI 3.1415926535897932384626433832795

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_MAKETYPE(d, cn) (CV_MAT_DEPTH(d) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC3 CV_MAKETYPE(CV_16S, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

enum { CV_StsNoMem = -4, CV_StsBadArg = -5, CV_StsNotImplemented = -213, CV_StsAssert = -215 };
enum { CV_DIST_L1 = 1, CV_DIST_L2 = 2 };

namespace cv {

typedef unsigned char uchar;
typedef long long int64;
inline int64 getTickCount() { return 0; }                   // (only feeds the reference's LOGLN timing lines)
inline double getTickFrequency() { return 1.0; }

class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
#define CV_Error(code, msg) throw ::cv::Exception(code, std::string(msg))
#define CV_Assert(expr) do { if (!(expr)) throw ::cv::Exception(CV_StsAssert, #expr); } while (0)

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
template <typename T> inline Point3_<T> operator-(const Point3_<T> &a, const Point3_<T> &b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_ &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_ &o) const { return !(*this == o); }
    T area() const { return width * height; }
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T> &a, const Point_<T> &b) : x(std::min(a.x, b.x)), y(std::min(a.y, b.y)),
        width(std::max(a.x, b.x) - std::min(a.x, b.x)), height(std::max(a.y, b.y) - std::min(a.y, b.y)) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    T area() const { return width * height; }
};
typedef Rect_<int> Rect;
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
};
struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

// cv::Ptr: reference-counted pointer with an implicit constructor from a raw pointer
template <typename T> class Ptr {
public:
    Ptr() {}
    Ptr(T *p) : p_(p) {}
    template <typename U> Ptr(const Ptr<U> &o) : p_(o.shared()) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
    operator T *() const { return p_.get(); }
    bool empty() const { return !p_; }
    void release() { p_.reset(); }
    T *obj() const { return p_.get(); }
    const std::shared_ptr<T> &shared() const { return p_; }
private:
    std::shared_ptr<T> p_;
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2 };

// ---------------------------------------------------------------------------------------------- Mat
class Mat {
public:
    int flags_type, rows, cols;
    size_t step;
    uchar *data;

    Mat() : flags_type(0), rows(0), cols(0), step(0), data(0) {}
    Mat(int r, int c, int t) : flags_type(0), rows(0), cols(0), step(0), data(0) { create(r, c, t); }
    Mat(Size s, int t) : flags_type(0), rows(0), cols(0), step(0), data(0) { create(s.height, s.width, t); }
    // user-allocated data (no ownership)
    Mat(int r, int c, int t, void *d, size_t s = 0)
        : flags_type(t), rows(r), cols(c), step(s ? s : (size_t)c * esz(t)), data(static_cast<uchar *>(d)) {}

    static size_t esz(int t)
    {
        static const size_t b[8] = {1, 1, 2, 2, 4, 4, 8, 0};
        return b[CV_MAT_DEPTH(t)] * CV_MAT_CN(t);
    }
    int type() const { return flags_type; }
    int depth() const { return CV_MAT_DEPTH(flags_type); }
    int channels() const { return CV_MAT_CN(flags_type); }
    size_t elemSize() const { return esz(flags_type); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

    void create(int r, int c, int t)
    {
        if (data && r == rows && c == cols && t == flags_type) return;
        rows = r; cols = c; flags_type = t; step = (size_t)c * esz(t);
        owner_.reset(static_cast<uchar *>(std::calloc((size_t)r * c + 1, esz(t))), std::free);
        data = owner_.get();
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { owner_.reset(); data = 0; rows = cols = 0; step = 0; }

    template <typename T> T *ptr(int y = 0) { return reinterpret_cast<T *>(data + (size_t)y * step); }
    template <typename T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(data + (size_t)y * step); }
    template <typename T> T &at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T &at(int y, int x) const { return ptr<T>(y)[x]; }

    Mat operator()(const Rect &r) const
    {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat operator()(Range rr, Range cr) const { return (*this)(Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start)); }

    so_mat so() const
    {
        so_mat s;
        s.data = data; s.rows = rows; s.cols = cols; s.type = flags_type; s.step = step;
        return s;
    }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat &dst) const
    {
        Mat out;                                            // (dst may alias *this)
        out.create(rows, cols, flags_type);
        for (int y = 0; y < rows; ++y) std::memcpy(out.ptr<uchar>(y), ptr<uchar>(y), (size_t)cols * elemSize());
        dst = out;
    }
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat &setTo(const Scalar &s, const Mat &mask = Mat());
    Mat &setTo(double v, const Mat &mask = Mat()) { return setTo(Scalar::all(v), mask); }
    Mat &operator+=(const Mat &o);
    Mat &operator*=(double v)                               // image *= gain: convertTo(image, -1, gain) on 8U (exposure_compensate.cpp:152)
    {
        CV_Assert(depth() == CV_8U);
        so_mat s = so();
        CV_Assert(so_scale_8u(&s, v) == 0);
        return *this;
    }
    Mat t() const;
    Mat inv() const;
    Mat reshape(int cn, int new_rows) const
    {
        CV_Assert(cn == 0 && isContinuous() && new_rows > 0 && (rows * cols) % new_rows == 0);
        Mat m(*this);
        m.rows = new_rows; m.cols = rows * cols / new_rows; m.step = (size_t)m.cols * elemSize();
        return m;
    }
    static Mat eye(int r, int c, int t)
    {
        CV_Assert(t == CV_32F);
        Mat m(r, c, t);
        for (int i = 0; i < std::min(r, c); ++i) m.at<float>(i, i) = 1.f;
        return m;
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    // (declared for util_inl.hpp's sqr(const Mat&), which the compositing path never calls)
    double dot(const Mat &o) const
    {
        CV_Assert(type() == CV_32F && o.type() == CV_32F && rows == o.rows && cols == o.cols);
        double s = 0;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) s += (double)at<float>(y, x) * o.at<float>(y, x);
        return s;
    }

private:
    std::shared_ptr<uchar> owner_;
};

template <typename T> struct DataType_;
template <> struct DataType_<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType_<short> { enum { type = CV_16SC1 }; };
template <> struct DataType_<int> { enum { type = CV_32SC1 }; };
template <> struct DataType_<float> { enum { type = CV_32FC1 }; };
template <> struct DataType_<double> { enum { type = CV_64FC1 }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataType_<T>::type) {}
    Mat_(const Mat &m) : Mat(m) { CV_Assert(m.empty() || m.elemSize() == sizeof(T)); }
    Mat_ &operator=(const Mat &m) { CV_Assert(m.empty() || m.elemSize() == sizeof(T)); Mat::operator=(m); return *this; }
    void create(int r, int c) { Mat::create(r, c, DataType_<T>::type); }
    void create(Size s) { Mat::create(s.height, s.width, DataType_<T>::type); }
    T &operator()(int y, int x) { return this->template at<T>(y, x); }
    const T &operator()(int y, int x) const { return this->template at<T>(y, x); }
};

inline const Mat &noArray() { static const Mat m; return m; }

// ---- element-wise helpers used by the reference ------------------------------------------------
inline float mat_get(const Mat &m, int y, int x)
{
    switch (m.depth()) {
    case CV_8U: return m.ptr<uchar>(y)[x];
    case CV_16S: return m.ptr<short>(y)[x];
    case CV_32F: return m.ptr<float>(y)[x];
    default: CV_Error(CV_StsNotImplemented, "shim: depth");
    }
}
enum { CMP_EQ = 0, CMP_GT = 1, CMP_LT = 3, CMP_NE = 5 };
inline Mat compare_scalar(const Mat &a, double v, int op)         // cv::compare: 255 where true, CV_8U
{
    CV_Assert(a.channels() == 1);
    Mat d(a.rows, a.cols, CV_8U);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            const double e = mat_get(a, y, x);
            const bool r = op == CMP_EQ ? e == v : op == CMP_GT ? e > v : op == CMP_LT ? e < v : e != v;
            d.ptr<uchar>(y)[x] = r ? 255 : 0;
        }
    return d;
}
inline Mat operator==(const Mat &a, double v) { return compare_scalar(a, v, CMP_EQ); }
inline Mat operator!=(const Mat &a, double v) { return compare_scalar(a, v, CMP_NE); }
inline Mat operator>(const Mat &a, double v) { return compare_scalar(a, v, CMP_GT); }
inline Mat operator<(const Mat &a, double v) { return compare_scalar(a, v, CMP_LT); }

// (submask1 == v) & (submask2 == w): bitwise AND of two CV_8U masks (exposure_compensate.cpp:104)
inline Mat operator&(const Mat &a, const Mat &b)
{
    CV_Assert(a.type() == CV_8U && b.type() == CV_8U && a.rows == b.rows && a.cols == b.cols);
    Mat d(a.rows, a.cols, CV_8U);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) d.ptr<uchar>(y)[x] = a.ptr<uchar>(y)[x] & b.ptr<uchar>(y)[x];
    return d;
}
inline int countNonZero(const Mat &m)
{
    CV_Assert(m.type() == CV_8U);
    int n = 0;
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols; ++x) n += m.ptr<uchar>(y)[x] != 0;
    return n;
}
// cv::solve(A, b, x) with the default DECOMP_LU on CV_64F: the oracle's restatement of OpenCV's LU (so_calib.c)
inline bool solve(const Mat &A, const Mat &b, Mat &x)
{
    CV_Assert(A.type() == CV_64FC1 && b.type() == CV_64FC1 && A.rows == A.cols && b.rows == A.rows && b.cols == 1);
    Mat a = A.clone(), r = b.clone();
    const bool ok = so_solve_lu(A.rows, a.ptr<double>(), r.ptr<double>()) != 0;
    x = r;
    return ok;
}
template <typename T> inline T saturate_cast(float v);
template <> inline uchar saturate_cast<uchar>(float v) { const int i = so_cvround(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }

// MatExpr `weight * sharpness` (CV_32F): evaluated by convertTo-style scaling, float(src * alpha)
inline Mat operator*(const Mat &a, double s)
{
    CV_Assert(a.type() == CV_32F);
    Mat d(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) d.ptr<float>(y)[x] = (float)((double)a.ptr<float>(y)[x] * s);
    return d;
}
inline Mat operator*(const Mat &a, float s) { return a * (double)s; }
// matrix product of two 3x3 CV_32F matrices (the len == 3 float fast path of cv::gemm)
inline Mat operator*(const Mat &a, const Mat &b)
{
    CV_Assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == 3 && a.cols == 3 && b.rows == 3 && b.cols == 3);
    float A[9], B[9], D[9];
    for (int i = 0; i < 9; ++i) { A[i] = a.at<float>(i / 3, i % 3); B[i] = b.at<float>(i / 3, i % 3); }
    so_mul3x3_f32(A, B, D);
    Mat d(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) d.at<float>(i / 3, i % 3) = D[i];
    return d;
}
inline Mat Mat::t() const
{
    CV_Assert(type() == CV_32F);
    Mat d(cols, rows, CV_32F);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) d.at<float>(x, y) = at<float>(y, x);
    return d;
}
inline Mat Mat::inv() const
{
    CV_Assert(type() == CV_32F && rows == 3 && cols == 3);
    float S[9], D[9];
    for (int i = 0; i < 9; ++i) S[i] = at<float>(i / 3, i % 3);
    so_inv3x3_f32(S, D);
    Mat d(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) d.at<float>(i / 3, i % 3) = D[i];
    return d;
}
inline Mat &Mat::setTo(const Scalar &s, const Mat &mask)
{
    CV_Assert(mask.empty() || (mask.type() == CV_8U && mask.rows == rows && mask.cols == cols));
    const int cn = channels();
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            if (!mask.empty() && !mask.ptr<uchar>(y)[x]) continue;
            for (int c = 0; c < cn; ++c) {
                const double v = s.val[c];
                switch (depth()) {
                case CV_8U: ptr<uchar>(y)[x * cn + c] = (uchar)v; break;
                case CV_16S: ptr<short>(y)[x * cn + c] = (short)v; break;
                case CV_32F: ptr<float>(y)[x * cn + c] = (float)v; break;
                case CV_32S: ptr<int>(y)[x * cn + c] = (int)v; break;
                case CV_64F: ptr<double>(y)[x * cn + c] = v; break;
                default: CV_Error(CV_StsNotImplemented, "shim: setTo depth");
                }
            }
        }
    return *this;
}
inline Mat &Mat::operator+=(const Mat &o)
{
    CV_Assert(type() == CV_32F && o.type() == CV_32F && rows == o.rows && cols == o.cols);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) ptr<float>(y)[x] += o.ptr<float>(y)[x];
    return *this;
}
inline void Mat::convertTo(Mat &dst, int rtype, double alpha, double beta) const
{
    const int dd = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype), cn = channels();
    Mat out(rows, cols, CV_MAKETYPE(dd, cn));
    CV_Assert(beta == 0);
    if (depth() == CV_8U && dd == CV_16S && alpha == 1) {
        so_mat s = so(), d = out.so();
        so_convert_8u_16s(&s, &d);
    } else if (depth() == CV_8U && dd == CV_32F) {
        // cvtScale_<uchar, float, float>: float(src) * (float)alpha
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols * cn; ++x) out.ptr<float>(y)[x] = (float)ptr<uchar>(y)[x] * (float)alpha;
    } else if (depth() == dd && alpha == 1) {
        copyTo(out);
    } else
        CV_Error(CV_StsNotImplemented, "shim: convertTo combination");
    dst = out;
}

inline void divide(const Mat &a, const Mat &b, Mat &dst)
{
    CV_Assert(a.type() == CV_32F && b.type() == CV_32F);
    Mat out(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            const float d = b.ptr<float>(y)[x];
            out.ptr<float>(y)[x] = d != 0 ? a.ptr<float>(y)[x] / d : 0.f;
        }
    dst = out;
}
inline void subtract(const Mat &a, const Mat &b, Mat &dst, const Mat &mask = noArray(), int dtype = -1)
{
    CV_Assert(mask.empty());
    so_mat sa = a.so(), sb = b.so();
    if (a.depth() == CV_16S && (dtype < 0 || dtype == CV_16S)) {
        Mat out(a.rows, a.cols, a.type());
        so_mat so_ = out.so();
        CV_Assert(so_subtract_16s(&sa, &sb, &so_) == 0);
        dst = out;
    } else if (a.depth() == CV_8U && dtype == CV_16S) {
        Mat out(a.rows, a.cols, CV_MAKETYPE(CV_16S, a.channels()));
        so_mat so_ = out.so();
        CV_Assert(so_subtract_8u_to_16s(&sa, &sb, &so_) == 0);
        dst = out;
    } else
        CV_Error(CV_StsNotImplemented, "shim: subtract combination");
}
inline void add(const Mat &a, const Mat &b, Mat &dst)
{
    CV_Assert(a.depth() == CV_16S);
    Mat out(a.rows, a.cols, a.type());
    so_mat sa = a.so(), sb = b.so(), so_ = out.so();
    CV_Assert(so_add_16s(&sa, &sb, &so_) == 0);
    dst = out;
}
// add(src, scalar, dst, mask): saturating 16S, elements outside the mask keep dst's previous value
inline void add(const Mat &a, int v, Mat &dst, const Mat &mask)
{
    CV_Assert(a.type() == CV_16SC1 && mask.type() == CV_8U && dst.data == a.data);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x)
            if (mask.ptr<uchar>(y)[x]) {
                const int s = a.ptr<short>(y)[x] + v;
                dst.ptr<short>(y)[x] = (short)std::min(std::max(s, -32768), 32767);
            }
}

// ---- imgproc ----------------------------------------------------------------------------------
inline void pyrDown(const Mat &src, Mat &dst, const Size & = Size())
{
    Mat out((src.rows + 1) / 2, (src.cols + 1) / 2, src.type());
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_pyr_down(&s, &d) == 0);
    dst = out;
}
inline void pyrUp(const Mat &src, Mat &dst, const Size &dstsize = Size())
{
    CV_Assert(dstsize.width == 0 || (dstsize.width == src.cols * 2 && dstsize.height == src.rows * 2));
    Mat out(src.rows * 2, src.cols * 2, src.type());
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_pyr_up(&s, &d) == 0);
    dst = out;
}
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int border, const Scalar & = Scalar())
{
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_copy_make_border(&s, &d, top, bottom, left, right, border) == 0);
    dst = out;
}
inline void remap(const Mat &src, Mat &dst, const Mat &xmap, const Mat &ymap, int interp, int border = BORDER_CONSTANT, const Scalar &bv = Scalar())
{
    Mat out(xmap.rows, xmap.cols, src.type());
    const uint8_t b[4] = {(uint8_t)bv.val[0], (uint8_t)bv.val[1], (uint8_t)bv.val[2], (uint8_t)bv.val[3]};
    so_mat s = src.so(), d = out.so(), mx = xmap.so(), my = ymap.so();
    CV_Assert(so_remap(&s, &d, &mx, &my, interp, border, b) == 0);
    dst = out;
}
inline void distanceTransform(const Mat &src, Mat &dst, int distanceType, int maskSize)
{
    CV_Assert(distanceType == CV_DIST_L1 && maskSize == 3);
    Mat out(src.rows, src.cols, CV_32F);
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_distance_l1_3x3(&s, &d) == 0);
    dst = out;
}
inline double threshold(const Mat &src, Mat &dst, double thresh, double, int type)
{
    CV_Assert(type == THRESH_TRUNC && src.type() == CV_32F);
    Mat out(src.rows, src.cols, CV_32F);
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            const float v = src.ptr<float>(y)[x];
            out.ptr<float>(y)[x] = v > (float)thresh ? (float)thresh : v;
        }
    dst = out;
    return thresh;
}
// cv::sepFilter2D with a symmetric 3-tap row/column kernel on CV_32F, BORDER_DEFAULT (exposure_compensate.cpp:217-218)
inline void sepFilter2D(const Mat &src, Mat &dst, int ddepth, const Mat &kx, const Mat &ky)
{
    CV_Assert(src.type() == CV_32F && ddepth == CV_32F && kx.type() == CV_32F && kx.rows == 1 && kx.cols == 3 && ky.data == kx.data);
    CV_Assert(kx.at<float>(0, 0) == kx.at<float>(0, 2));
    Mat out(src.rows, src.cols, CV_32F);
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_sep_filter3_f32(&s, &d, kx.at<float>(0, 1), kx.at<float>(0, 0)) == 0);
    dst = out;
}
// cv::resize INTER_LINEAR on CV_32FC1 (exposure_compensate.cpp:233)
inline void resize(const Mat &src, Mat &dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR)
{
    CV_Assert(src.type() == CV_32F && interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0);
    Mat out(dsize.height, dsize.width, CV_32F);
    so_mat s = src.so(), d = out.so();
    CV_Assert(so_resize_linear_32f(&s, &d) == 0);
    dst = out;
}
template <typename T> inline T randu() { return (T)std::rand(); }

// gpu::GpuMat appears in the signatures of the *WarperGpu classes (compiled out: no HAVE_OPENCV_GPU)
namespace gpu {
class GpuMat {
public:
    void upload(const Mat &) {}
    void download(Mat &) const {}
};
}  // namespace gpu

}  // namespace cv
#endif
