// host_projector.cpp — host half of the RotationWarper: camera-parameter products, forward
// projection and destination-ROI detection.  Strict float32, one rounding per operation
// (compiled with -ffp-contract=off), libm atan2f/acosf/sqrtf as the reference's own host code.
#include <algorithm>
#include <cmath>
#include <limits>

#include "sb_host_projector.h"

namespace sb {

namespace {
const float kPiF = static_cast<float>(3.1415926535897932384626433832795);

// cv::invert of a 3x3 CV_32F matrix (OpenCV 2.4.11 lapack.cpp: cofactors and determinant in
// double, result stored as float; a singular matrix gives zeros)
void invert3x3(const float s[9], float d[9])
{
    auto S = [&](int r, int c) { return static_cast<double>(s[r * 3 + c]); };
    double det = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
                 S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    if (det == 0.) {
        std::fill(d, d + 9, 0.f);
        return;
    }
    const double id = 1. / det;
    d[0] = static_cast<float>((S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * id);
    d[1] = static_cast<float>((S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * id);
    d[2] = static_cast<float>((S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * id);
    d[3] = static_cast<float>((S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * id);
    d[4] = static_cast<float>((S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * id);
    d[5] = static_cast<float>((S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * id);
    d[6] = static_cast<float>((S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * id);
    d[7] = static_cast<float>((S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * id);
    d[8] = static_cast<float>((S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * id);
}

// Mat * Mat for 3x3 CV_32F (gemm small-matrix path: float products summed left to right)
void matmul3x3(const float a[9], const float b[9], float d[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            d[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

struct Extent {
    float tl_u = std::numeric_limits<float>::max(), tl_v = std::numeric_limits<float>::max();
    float br_u = -std::numeric_limits<float>::max(), br_v = -std::numeric_limits<float>::max();
    void add(float u, float v)
    {
        tl_u = std::min(tl_u, u); tl_v = std::min(tl_v, v);
        br_u = std::max(br_u, u); br_v = std::max(br_v, v);
    }
};
}  // namespace

void projector_set(ProjParams &p, int kind, float scale, const float K[9], const float R[9], const float T[3])
{
    p.kind = kind;
    p.scale = scale;
    std::copy(K, K + 9, p.k);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) p.rinv[i * 3 + j] = R[j * 3 + i];
    float kinv[9];
    invert3x3(K, kinv);
    matmul3x3(R, kinv, p.r_kinv);
    matmul3x3(K, p.rinv, p.k_rinv);
    for (int i = 0; i < 3; ++i) p.t[i] = T ? T[i] : 0.f;
}

void projector_map_forward(const ProjParams &p, float x, float y, float *u, float *v)
{
    const float *m = p.r_kinv;
    float x_ = m[0] * x + m[1] * y + m[2];
    float y_ = m[3] * x + m[4] * y + m[5];
    float z_ = m[6] * x + m[7] * y + m[8];
    switch (p.kind) {
    case SB_WARP_PLANE:
        x_ = p.t[0] + x_ / z_ * (1 - p.t[2]);
        y_ = p.t[1] + y_ / z_ * (1 - p.t[2]);
        *u = p.scale * x_;
        *v = p.scale * y_;
        break;
    case SB_WARP_SPHERICAL: {
        *u = p.scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = p.scale * (kPiF - acosf(w == w ? w : 0));
        break;
    }
    default:
        *u = p.scale * atan2f(x_, z_);
        *v = p.scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

void projector_detect_result_roi(const ProjParams &p, int src_w, int src_h, sb_point *tl, sb_point *br)
{
    Extent e;
    float u, v;
    if (p.kind == SB_WARP_PLANE) {
        const float xs[2] = {0.f, static_cast<float>(src_w - 1)}, ys[2] = {0.f, static_cast<float>(src_h - 1)};
        for (float x : xs)
            for (float y : ys) {
                projector_map_forward(p, x, y, &u, &v);
                e.add(u, v);
            }
    } else {
        for (float x = 0; x < src_w; ++x) {
            projector_map_forward(p, x, 0, &u, &v); e.add(u, v);
            projector_map_forward(p, x, static_cast<float>(src_h - 1), &u, &v); e.add(u, v);
        }
        for (int y = 0; y < src_h; ++y) {
            projector_map_forward(p, 0, static_cast<float>(y), &u, &v); e.add(u, v);
            projector_map_forward(p, static_cast<float>(src_w - 1), static_cast<float>(y), &u, &v); e.add(u, v);
        }
    }
    tl->x = static_cast<int>(e.tl_u); tl->y = static_cast<int>(e.tl_v);
    br->x = static_cast<int>(e.br_u); br->y = static_cast<int>(e.br_v);
    if (p.kind != SB_WARP_SPHERICAL) return;

    // a pole inside the image extends the ROI to v = pi*scale / v = 0 (warpers.cpp:180-206)
    Extent s;
    s.tl_u = static_cast<float>(tl->x); s.tl_v = static_cast<float>(tl->y);
    s.br_u = static_cast<float>(br->x); s.br_v = static_cast<float>(br->y);
    for (int pole = 0; pole < 2; ++pole) {
        float x = p.rinv[1], y = pole == 0 ? p.rinv[4] : -p.rinv[4], z = p.rinv[7];
        if (y > 0.f) {
            float x_ = (p.k[0] * x + p.k[1] * y) / z + p.k[2];
            float y_ = p.k[4] * y / z + p.k[5];
            if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h)
                s.add(0.f, pole == 0 ? static_cast<float>(3.1415926535897932384626433832795 * p.scale) : 0.f);
        }
    }
    tl->x = static_cast<int>(s.tl_u); tl->y = static_cast<int>(s.tl_v);
    br->x = static_cast<int>(s.br_u); br->y = static_cast<int>(s.br_v);
}

}  // namespace sb
