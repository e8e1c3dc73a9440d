// host_projector.cpp — host half of the RotationWarper: camera-parameter products, forward
// projection and destination-ROI detection.  Strict float32, one rounding per operation
// (compiled with -ffp-contract=off), libm atan2f/acosf/sqrtf as the reference's own host code.
#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>

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
    p.a = p.b = 1.f;
}

void projector_map_forward(const ProjParams &p, float x, float y, float *u, float *v)
{
    const float *m = p.r_kinv;
    const float scale = p.scale, a = p.a, b = p.b;
    float x_ = m[0] * x + m[1] * y + m[2];
    float y_ = m[3] * x + m[4] * y + m[5];
    float z_ = m[6] * x + m[7] * y + m[8];
    switch (p.kind) {
    case SB_WARP_PLANE:
        x_ = p.t[0] + x_ / z_ * (1 - p.t[2]);
        y_ = p.t[1] + y_ / z_ * (1 - p.t[2]);
        *u = scale * x_;
        *v = scale * y_;
        break;
    case SB_WARP_SPHERICAL: {
        *u = scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = scale * (kPiF - acosf(w == w ? w : 0));
        break;
    }
    case SB_WARP_CYLINDRICAL:
        *u = scale * atan2f(x_, z_);
        *v = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
        break;
    case SB_WARP_FISHEYE: {                                  // warpers_inl.hpp:302-313
        float u_ = atan2f(x_, z_);
        float v_ = kPiF - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        *u = scale * v_ * cosf(u_);
        *v = scale * v_ * sinf(u_);
        break;
    }
    case SB_WARP_STEREOGRAPHIC: {                            // :339-353 (cos / sin of the double overloads, as there)
        float u_ = atan2f(x_, z_);
        float v_ = kPiF - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float r = sinf(v_) / (1 - cosf(v_));
        *u = static_cast<float>(scale * r * std::cos(static_cast<double>(u_)));
        *v = static_cast<float>(scale * r * std::sin(static_cast<double>(u_)));
        break;
    }
    case SB_WARP_COMPRESSED_RECTILINEAR:                     // :380-392
    case SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT: {          // :419-431: x and y swap roles, u changes sign
        if (p.kind == SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT) std::swap(x_, y_);
        float u_ = atan2f(x_, z_);
        float v_ = asinf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        *u = p.kind == SB_WARP_COMPRESSED_RECTILINEAR ? scale * a * tanf(u_ / a) : -scale * a * tanf(u_ / a);
        *v = scale * b * tanf(v_) / cosf(u_);
        break;
    }
    case SB_WARP_PANINI:                                     // :458-476
    case SB_WARP_PANINI_PORTRAIT: {                          // :508-526
        if (p.kind == SB_WARP_PANINI_PORTRAIT) std::swap(x_, y_);
        float u_ = atan2f(x_, z_);
        float v_ = asinf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float tg = a * tanf(u_ / a);
        *u = p.kind == SB_WARP_PANINI ? scale * tg : -scale * tg;
        float sinu = sinf(u_);
        if (std::fabs(sinu) < 1E-7) *v = scale * b * tanf(v_);
        else *v = scale * b * tg * tanf(v_) / sinu;
        break;
    }
    case SB_WARP_MERCATOR: {                                 // :559-571
        float u_ = atan2f(x_, z_);
        float v_ = asinf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        *u = scale * u_;
        *v = scale * logf(tanf(static_cast<float>(3.1415926535897932384626433832795 / 4) + v_ / 2));
        break;
    }
    case SB_WARP_TRANSVERSE_MERCATOR: {                      // :597-611
        float u_ = atan2f(x_, z_);
        float v_ = asinf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float B = cosf(v_) * sinf(u_);
        *u = scale / 2 * logf((1 + B) / (1 - B));
        *v = scale * atan2f(tanf(v_), cosf(u_));
        break;
    }
    case SB_WARP_SPHERICAL_PORTRAIT: {                       // :637-653: x and y swap roles, u changes sign
        std::swap(x_, y_);
        float uu = scale * atan2f(x_, z_);
        float vv = scale * (kPiF - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_)));
        *u = -uu;
        *v = vv;
        break;
    }
    case SB_WARP_CYLINDRICAL_PORTRAIT: {                     // :683-699
        std::swap(x_, y_);
        float uu = scale * atan2f(x_, z_);
        float vv = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
        *u = -uu;
        *v = vv;
        break;
    }
    default: {                                               // SB_WARP_PLANE_PORTRAIT :727-745
        std::swap(x_, y_);
        x_ = p.t[0] + x_ / z_ * (1 - p.t[2]);
        y_ = p.t[1] + y_ / z_ * (1 - p.t[2]);
        float uu = scale * x_, vv = scale * y_;
        *u = -uu;
        *v = vv;
    }
    }
}

namespace {
// x, y = K * R^T * ray, perspective division when the ray points forward (every rotation projector's tail)
inline void project_ray(const ProjParams &p, float x_, float y_, float z_, float *x, float *y)
{
    const float *m = p.k_rinv;
    float xx = m[0] * x_ + m[1] * y_ + m[2] * z_;
    float yy = m[3] * x_ + m[4] * y_ + m[5] * z_;
    float z = m[6] * x_ + m[7] * y_ + m[8] * z_;
    if (z > 0) { xx /= z; yy /= z; }
    else xx = yy = -1;
    *x = xx; *y = yy;
}
}  // namespace

void projector_map_backward(const ProjParams &p, float u, float v, float *x, float *y)
{
    const float scale = p.scale, a = p.a, b = p.b;
    switch (p.kind) {
    case SB_WARP_PLANE: {                                    // warpers_inl.hpp:222-234
        u = u / scale - p.t[0];
        v = v / scale - p.t[1];
        const float *m = p.k_rinv;
        float xx = m[0] * u + m[1] * v + m[2] * (1 - p.t[2]);
        float yy = m[3] * u + m[4] * v + m[5] * (1 - p.t[2]);
        float z = m[6] * u + m[7] * v + m[8] * (1 - p.t[2]);
        *x = xx / z; *y = yy / z;
        return;
    }
    case SB_WARP_PLANE_PORTRAIT: {                           // :748-759
        float uu = -u, vv = v;
        uu = uu / scale - p.t[0];
        vv = vv / scale - p.t[1];
        const float *m = p.k_rinv;
        float xx = m[0] * vv + m[1] * uu + m[2] * (1 - p.t[2]);
        float yy = m[3] * vv + m[4] * uu + m[5] * (1 - p.t[2]);
        float z = m[6] * vv + m[7] * uu + m[8] * (1 - p.t[2]);
        *x = xx / z; *y = yy / z;
        return;
    }
    case SB_WARP_SPHERICAL: {                                // :250-268
        u /= scale; v /= scale;
        float sinv = sinf(kPiF - v);
        project_ray(p, sinv * sinf(u), cosf(kPiF - v), sinv * cosf(u), x, y);
        return;
    }
    case SB_WARP_CYLINDRICAL:                                // :283-300
        u /= scale; v /= scale;
        project_ray(p, sinf(u), v, cosf(u), x, y);
        return;
    case SB_WARP_FISHEYE: {                                  // :316-336
        u /= scale; v /= scale;
        float u_ = atan2f(v, u);
        float v_ = sqrtf(u * u + v * v);
        float sinv = sinf(kPiF - v_);
        project_ray(p, sinv * sinf(u_), cosf(kPiF - v_), sinv * cosf(u_), x, y);
        return;
    }
    case SB_WARP_STEREOGRAPHIC: {                            // :356-377
        u /= scale; v /= scale;
        float u_ = atan2f(v, u);
        float r = sqrtf(u * u + v * v);
        float v_ = 2 * atanf(1.f / r);
        float sinv = sinf(kPiF - v_);
        project_ray(p, sinv * sinf(u_), cosf(kPiF - v_), sinv * cosf(u_), x, y);
        return;
    }
    case SB_WARP_COMPRESSED_RECTILINEAR:                     // :395-416
    case SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT: {          // :434-455
        const bool portrait = p.kind == SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT;
        u /= portrait ? -scale : scale;
        v /= scale;
        float aatg = a * atanf(u / a);
        float u_ = aatg;
        float v_ = atanf(v * cosf(aatg) / b);
        float cosv = cosf(v_);
        float s_ = cosv * sinf(u_), t_ = sinf(v_), z_ = cosv * cosf(u_);
        if (portrait) project_ray(p, t_, s_, z_, x, y);
        else project_ray(p, s_, t_, z_, x, y);
        return;
    }
    case SB_WARP_PANINI:                                     // :479-505
    case SB_WARP_PANINI_PORTRAIT: {                          // :529-556
        const bool portrait = p.kind == SB_WARP_PANINI_PORTRAIT;
        u /= portrait ? -scale : scale;
        v /= scale;
        float lamda = a * atanf(u / a);
        float u_ = lamda;
        float v_;
        if (std::fabs(lamda) > 1E-7) v_ = atanf(v * sinf(lamda) / (b * a * tanf(lamda / a)));
        else v_ = atanf(v / b);
        float cosv = cosf(v_);
        float s_ = cosv * sinf(u_), t_ = sinf(v_), z_ = cosv * cosf(u_);
        if (portrait) project_ray(p, t_, s_, z_, x, y);
        else project_ray(p, s_, t_, z_, x, y);
        return;
    }
    case SB_WARP_MERCATOR: {                                 // :574-594
        u /= scale; v /= scale;
        float v_ = atanf(sinhf(v));
        float u_ = u;
        float cosv = cosf(v_);
        project_ray(p, cosv * sinf(u_), sinf(v_), cosv * cosf(u_), x, y);
        return;
    }
    case SB_WARP_TRANSVERSE_MERCATOR: {                      // :614-634 (cos of the double overload, as there)
        u /= scale; v /= scale;
        float v_ = asinf(sinf(v) / coshf(u));
        float u_ = atan2f(sinhf(u), static_cast<float>(std::cos(static_cast<double>(v))));
        float cosv = cosf(v_);
        project_ray(p, cosv * sinf(u_), sinf(v_), cosv * cosf(u_), x, y);
        return;
    }
    case SB_WARP_SPHERICAL_PORTRAIT: {                       // :656-680
        float uu = -u, vv = v;
        uu /= scale; vv /= scale;
        float sinv = sinf(kPiF - vv);
        float x0 = sinv * sinf(uu), y0 = cosf(kPiF - vv), z_ = sinv * cosf(uu);
        project_ray(p, y0, x0, z_, x, y);
        return;
    }
    default: {                                               // SB_WARP_CYLINDRICAL_PORTRAIT :702-724
        float uu = -u, vv = v;
        uu /= scale; vv /= scale;
        float x0 = sinf(uu), y0 = vv, z_ = cosf(uu);
        project_ray(p, y0, x0, z_, x, y);
    }
    }
}

void projector_build_maps_host(const ProjParams &p, sb_point tl, sb_point br, float *xmap, float *ymap)
{
    const int w = br.x - tl.x + 1;
    for (int v = tl.y; v <= br.y; ++v)
        for (int u = tl.x; u <= br.x; ++u) {
            const size_t i = (size_t)(v - tl.y) * w + (u - tl.x);
            projector_map_backward(p, static_cast<float>(u), static_cast<float>(v), &xmap[i], &ymap[i]);
        }
}

void projector_detect_result_roi(const ProjParams &p, int src_w, int src_h, sb_point *tl, sb_point *br)
{
    Extent e;
    float u, v;
    const bool by_border = p.kind == SB_WARP_SPHERICAL || p.kind == SB_WARP_CYLINDRICAL || p.kind == SB_WARP_SPHERICAL_PORTRAIT ||
                           p.kind == SB_WARP_CYLINDRICAL_PORTRAIT || p.kind == SB_WARP_PLANE_PORTRAIT;
    if (p.kind == SB_WARP_PLANE) {
        const float xs[2] = {0.f, static_cast<float>(src_w - 1)}, ys[2] = {0.f, static_cast<float>(src_h - 1)};
        for (float x : xs)
            for (float y : ys) {
                projector_map_forward(p, x, y, &u, &v);
                e.add(u, v);
            }
    } else if (!by_border) {      // RotationWarperBase<P>::detectResultRoi: every source pixel (warpers_inl.hpp:142-166)
        for (int y = 0; y < src_h; ++y)
            for (int x = 0; x < src_w; ++x) {
                projector_map_forward(p, static_cast<float>(x), static_cast<float>(y), &u, &v);
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
    if (p.kind != SB_WARP_SPHERICAL && p.kind != SB_WARP_SPHERICAL_PORTRAIT) return;

    // a pole inside the image extends the ROI to v = pi*scale / v = 0 (warpers.cpp:180-206; portrait: :389-430,
    // where the pole direction is the first column of R^T instead of the second)
    const int c0 = p.kind == SB_WARP_SPHERICAL ? 1 : 0;
    Extent s;
    s.tl_u = static_cast<float>(tl->x); s.tl_v = static_cast<float>(tl->y);
    s.br_u = static_cast<float>(br->x); s.br_v = static_cast<float>(br->y);
    for (int pole = 0; pole < 2; ++pole) {
        float x = p.rinv[c0], y = pole == 0 ? p.rinv[c0 + 3] : -p.rinv[c0 + 3], z = p.rinv[c0 + 6];
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
