// kernels_calib.cu — the once-per-(re)calibration steps next to the per-frame path:
//   * overlap statistics of GainCompensator::feed / BlocksGainCompensator::feed (exposure_compensate.cpp:93-126),
//   * the seam-mask refinement of the compose loop (stitcher.cpp:291-294): dilate 3x3, resize INTER_LINEAR (8UC1), AND.
// They stall the video pipeline for 1-2 s on the host whenever the rig is recalibrated (APP64:696-722); here they are
// a handful of streaming kernels.
#include "sb_device.cuh"
#include <algorithm>
#include <cmath>

#include "sb_kernels.h"

namespace sb {
using namespace sbd;

// One overlap of two images: the reference walks the overlap row-major and adds sqrt(double(r^2 + g^2 + b^2)) of both
// images where both masks carry their level value.  A double sum depends on its order, so a parallel reduction can not
// reproduce the reference's last bits; instead the sum is made EXACT and order independent: every term is a double in
// [1, 442) (or 0), i.e. an integer multiple of 2^-52 below 2^61, and is accumulated as a 128-bit integer (two 64-bit
// halves, carries resolved on the host).  The host rounds the exact sum once.  Result: deterministic run to run and
// within one rounding of the true sum (the reference's sequential sum carries up to N roundings).
__global__ void __launch_bounds__(256) k_overlap_stats(const OverlapPair *pairs, unsigned long long *out)
{
    const OverlapPair p = pairs[blockIdx.y];
    unsigned long long cnt = 0, lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
    for (int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); y < p.h; y += gridDim.x * (blockDim.x >> 5)) {
        const uint8_t *m1 = p.mask1 + (size_t)y * p.mstep1, *m2 = p.mask2 + (size_t)y * p.mstep2;
        const uint8_t *r1 = p.img1 + (size_t)y * p.istep1, *r2 = p.img2 + (size_t)y * p.istep2;
        for (int x = threadIdx.x & 31; x < p.w; x += 32) {
            if (m1[x] != p.val1 || m2[x] != p.val2) continue;
            ++cnt;
            const int a0 = r1[3 * x], a1 = r1[3 * x + 1], a2 = r1[3 * x + 2];
            const int b0 = r2[3 * x], b1 = r2[3 * x + 1], b2 = r2[3 * x + 2];
            const unsigned long long v1 = (unsigned long long)__double2ll_rz(__dmul_rn(__dsqrt_rn((double)(a0 * a0 + a1 * a1 + a2 * a2)), 4503599627370496.0));
            const unsigned long long v2 = (unsigned long long)__double2ll_rz(__dmul_rn(__dsqrt_rn((double)(b0 * b0 + b1 * b1 + b2 * b2)), 4503599627370496.0));
            lo1 += v1 & 0xffffffffull; hi1 += v1 >> 32;
            lo2 += v2 & 0xffffffffull; hi2 += v2 >> 32;
        }
    }
    unsigned long long v[5] = {cnt, lo1, hi1, lo2, hi2};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(out + (size_t)blockIdx.y * 5 + k, v[k]);   // integer adds: order independent
    }
}

int launch_overlap_stats(const OverlapPair *pairs_dev, int n_pairs, int max_h, unsigned long long *out_dev, cudaStream_t s)
{
    if (n_pairs <= 0) return SB_OK;
    SB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(unsigned long long) * 5 * (size_t)n_pairs, s));
    const int chunks = std::max(1, std::min(64, div_up(max_h, 32)));
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int np = std::min(65535, n_pairs - p0);
        k_overlap_stats<<<dim3(chunks, np), 256, 0, s>>>(pairs_dev + p0, out_dev + (size_t)p0 * 5);
        SB_LAUNCHED();
    }
    return SB_OK;
}

// cv::dilate(src, dst, Mat()): 3x3 rectangle, pixels outside the image do not take part (8UC1)
__global__ void __launch_bounds__(256) k_dilate3x3(const uint8_t *src, size_t sstep, uint8_t *dst, size_t dstep, int w, int h)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int m = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if ((unsigned)yy >= (unsigned)h) continue;
        const uint8_t *r = src + (size_t)yy * sstep;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
            if ((unsigned)(x + dx) < (unsigned)w) m = max(m, (int)r[x + dx]);
    }
    dst[(size_t)y * dstep + x] = (uint8_t)m;
}

int launch_dilate3x3(const DImage &src, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_8UC1 && dst.type == SB_8UC1 && src.rows == dst.rows && src.cols == dst.cols && src.data != dst.data);
    k_dilate3x3<<<dim3(div_up(src.cols, 32), div_up(src.rows, 8)), dim3(32, 8), 0, s>>>(src.ptr<uint8_t>(), src.step, dst.ptr<uint8_t>(), dst.step, src.cols, src.rows);
    SB_LAUNCHED();
    return SB_OK;
}

// cv::resize INTER_LINEAR on 8UC1 (OpenCV 2.4.11 imgwarp.cpp): 11-bit fixed-point coefficients (cvRound(f * 2048) as
// short), horizontal pass in int, vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
// AND_MASK: the result is ANDed with `andm` (stitcher.cpp:294), fused.
// mode 0: the fixed-point path; 1: exact 2x2 decimation rerouted to INTER_AREA; 2: equal sizes (copy).
template <bool AND_MASK>
__global__ void __launch_bounds__(256)
k_resize_linear_8u(const uint8_t *src, size_t sstep, int sw, int sh, uint8_t *dst, size_t dstep, int dw, int dh, double scale_x, double scale_y,
                   int mode, const uint8_t *andm, size_t astep)
{
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    int v;
    if (mode == 2) {
        v = src[(size_t)dy * sstep + dx];
    } else if (mode == 1) {
        const uint8_t *S0 = src + (size_t)(2 * dy) * sstep, *S1 = S0 + sstep;
        v = (S0[2 * dx] + S0[2 * dx + 1] + S1[2 * dx] + S1[2 * dx + 1] + 2) >> 2;
    } else {
        float fx = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)dx, 0.5), scale_x), 0.5);
        int sx = (int)floorf(fx);
        fx = __fsub_rn(fx, (float)sx);
        bool edge = false;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx + 1 >= sw) { edge = true; fx = 0.f; sx = sw - 1; }
        const int a0 = (short)cvround(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = (short)cvround(__fmul_rn(fx, 2048.f));
        float fy = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)dy, 0.5), scale_y), 0.5);
        const int sy = (int)floorf(fy);
        fy = __fsub_rn(fy, (float)sy);
        const int b0 = (short)cvround(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = (short)cvround(__fmul_rn(fy, 2048.f));
        const int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1);
        const uint8_t *S0 = src + (size_t)sy0 * sstep, *S1 = src + (size_t)sy1 * sstep;
        int r0, r1;
        if (!edge) { r0 = S0[sx] * a0 + S0[sx + 1] * a1; r1 = S1[sx] * a0 + S1[sx + 1] * a1; }
        else { r0 = S0[sx] * 2048; r1 = S1[sx] * 2048; }
        v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    }
    if (AND_MASK) v &= andm[(size_t)dy * astep + dx];
    dst[(size_t)dy * dstep + dx] = (uint8_t)v;
}

int launch_resize_linear_8u(const DImage &src, const DImage &dst, const DImage *and_mask, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_8UC1 && dst.type == SB_8UC1 && !src.empty() && !dst.empty());
    if (and_mask) SB_ASSERT(and_mask->type == SB_8UC1 && and_mask->rows == dst.rows && and_mask->cols == dst.cols);
    const double inv_scale_x = (double)dst.cols / src.cols, inv_scale_y = (double)dst.rows / src.rows;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    const int iscale_x = (int)nearbyint(scale_x), iscale_y = (int)nearbyint(scale_y);
    const bool area_fast = std::fabs(scale_x - iscale_x) < 2.220446049250313e-16 && std::fabs(scale_y - iscale_y) < 2.220446049250313e-16;
    const int mode = (src.cols == dst.cols && src.rows == dst.rows) ? 2 : (area_fast && iscale_x == 2 && iscale_y == 2) ? 1 : 0;
    const dim3 grid(div_up(dst.cols, 32), div_up(dst.rows, 8)), block(32, 8);
    if (and_mask)
        k_resize_linear_8u<true><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.step, src.cols, src.rows, dst.ptr<uint8_t>(), dst.step, dst.cols, dst.rows,
                                                        scale_x, scale_y, mode, and_mask->ptr<uint8_t>(), and_mask->step);
    else
        k_resize_linear_8u<false><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.step, src.cols, src.rows, dst.ptr<uint8_t>(), dst.step, dst.cols, dst.rows,
                                                         scale_x, scale_y, mode, nullptr, 0);
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
