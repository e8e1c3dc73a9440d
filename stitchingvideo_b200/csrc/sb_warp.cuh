// sb_warp.cuh — projector arithmetic shared by the warp kernels (warpers_inl.hpp:206-300).
#pragma once
#include "sb_device.cuh"
#include "sb_kernels.h"

namespace sb {
using namespace sbd;

#define SB_PI_F 3.14159274101257324219f /* static_cast<float>(CV_PI) */

struct Mat3 { float m[9]; };

// x,y,z = M * (a,b,c), each product and sum rounded separately, left to right (C evaluation order)
__device__ __forceinline__ void mul3(const float *m, float a, float b, float c, float &x, float &y, float &z)
{
    x = __fadd_rn(__fadd_rn(__fmul_rn(m[0], a), __fmul_rn(m[1], b)), __fmul_rn(m[2], c));
    y = __fadd_rn(__fadd_rn(__fmul_rn(m[3], a), __fmul_rn(m[4], b)), __fmul_rn(m[5], c));
    z = __fadd_rn(__fadd_rn(__fmul_rn(m[6], a), __fmul_rn(m[7], b)), __fmul_rn(m[8], c));
}

// {Plane,Cylindrical,Spherical}Projector::mapBackward (warpers_inl.hpp:222-236, 250-268, 283-300)
template <int KIND>
__device__ __forceinline__ void map_backward(const ProjParams &p, float u, float v, float &x, float &y)
{
    float z;
    if (KIND == SB_WARP_PLANE) {
        u = __fsub_rn(__fdiv_rn(u, p.scale), p.t[0]);
        v = __fsub_rn(__fdiv_rn(v, p.scale), p.t[1]);
        mul3(p.k_rinv, u, v, __fsub_rn(1.f, p.t[2]), x, y, z);
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
        return;
    }
    u = __fdiv_rn(u, p.scale);
    v = __fdiv_rn(v, p.scale);
    float x_, y_, z_;
    if (KIND == SB_WARP_SPHERICAL) {
        float sinv = sinf_exact(__fsub_rn(SB_PI_F, v));
        x_ = __fmul_rn(sinv, sinf_exact(u));
        y_ = cosf_exact(__fsub_rn(SB_PI_F, v));
        z_ = __fmul_rn(sinv, cosf_exact(u));
    } else {
        x_ = sinf_exact(u);
        y_ = v;
        z_ = cosf_exact(u);
    }
    mul3(p.k_rinv, x_, y_, z_, x, y, z);
    if (z > 0) {
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
    } else
        x = y = -1.f;
}

template <int KIND>
__device__ __forceinline__ void map_backward_tab(const float *__restrict__ k_rinv, float t2, float cs, float cc, float ra,
                                                 float rb, float &x, float &y)
{
    float z;
    if (KIND == SB_WARP_PLANE) {
        mul3(k_rinv, cs, ra, __fsub_rn(1.f, t2), x, y, z);
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
        return;
    }
    float x_, y_, z_;
    if (KIND == SB_WARP_SPHERICAL) {
        x_ = __fmul_rn(ra, cs); y_ = rb; z_ = __fmul_rn(ra, cc);
    } else {
        x_ = cs; y_ = ra; z_ = cc;
    }
    mul3(k_rinv, x_, y_, z_, x, y, z);
    if (z > 0) {
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
    } else
        x = y = -1.f;
}


}  // namespace sb
