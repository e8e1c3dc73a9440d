// kernels_fused.cu — panorama-centric fused kernels of the compositor.
//
// The reference accumulates camera by camera into the panorama (read-modify-write per feed,
// blenders.cpp:123-147, 300-356) and then normalises, collapses and crops in further passes.
// `dst += short(...)` wraps mod 2^16, so the sum over cameras is order independent and every
// zero-weight term is exactly 0: the same result is obtained by visiting each PANORAMA pixel once,
// summing the contributions of the cameras whose weight there is non-zero, and finishing the pixel
// (normalise -> [collapse add] -> mask -> convertTo(8U)) in registers.  No accumulator ever exists
// in HBM: no zero-fill, no read-modify-write, no separate normalise / crop passes.
//
//   k_feather_fused : remap (pre-quantised fixed-point map table, 8 B/px incl. the feather distance)
//                     + gain + convertTo(16S) + FeatherBlender::feed over all cameras + blend +
//                     convertTo(8U)  — ONE launch per frame.
//   k_band_fused    : one Laplacian band of MultiBandBlender for all cameras: Laplacian formed on
//                     the fly from the cameras' Gaussian levels, weighted sum, normalise, collapse
//                     add of the coarser restored band, and at band 0 crop + mask + convertTo(8U).
//
// HBM-bound integer/byte work: each thread owns 4 consecutive panorama pixels so that the weight
// sums load as float4 and the panorama stores as 3 x 32-bit (8U) or 3 x 64-bit (16S) words.
#include <climits>

#include "sb_device.cuh"
#include "sb_fused.h"
#include "sb_pyr.cuh"
#include "sb_warp.cuh"

namespace sb {
using namespace sbd;

#define SB_WEIGHT_EPS 1e-5f

__device__ __forceinline__ bool in_spans(const int span[4], int x0, int x1)   // [x0, x1) overlaps a span?
{
    return (x0 < span[1] && x1 > span[0]) || (x0 < span[3] && x1 > span[2]);
}

// ------------------------------------------------------------------------------------ feather
// setup: the bilinear product table shared by every remap-type kernel (sb_device.cuh)
__global__ void k_build_bilin_lut(uint2 *lut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1024) lut[i] = bilin_weights(i & 31, i >> 5);
}

int launch_build_bilin_lut(uint2 *lut, cudaStream_t s)
{
    k_build_bilin_lut<<<4, 256, 0, s>>>(lut);
    SB_LAUNCHED();
    return SB_OK;
}

// setup: one table entry per warped pixel = what cv::remap's map conversion would compute every
// frame (Appendix A1: sx, sy after >> 5 and the int16 clamp, 5+5 fractional bits) with the
// BORDER_REFLECT of the two tap coordinates already resolved, + the L1 distance behind the feather
// weight:   x = x0 | y0 << 13 | (x1 == x0) << 26 | (y1 == y0) << 27      y = fx | fy << 5 | dist << 16
// A non-zero distance implies the pixel lies inside the warped all-255 mask, i.e. its source
// coordinate rounds into the image (sx in [-1, sw-1], sy in [-1, sh-1]): there x1 is x0 or x0 + 1.
// Pixels whose taps are folded any other way are given distance 0 only if the mask is 0 there, so the
// kernel asserts nothing: entries with dist == 0 are skipped, exactly (short(p * 0) == 0, w += 0).
// KIND < 0: the projector's maps were built on the host (the libm-heavy projectors of warpers_inl.hpp:302-759) and are read
// from xm / ym instead of being recomputed
template <int KIND>
__global__ void __launch_bounds__(256)
k_build_feather_table(ProjParams p, int tl_x, int tl_y, const float *dist, size_t dstep, int w, int h, int sw, int sh, uint2 *table, size_t tstep,
                      const float *xm, const float *ym, size_t mstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float mx, my;
    if (KIND < 0) { mx = crow<float>(xm, mstep, y)[x]; my = crow<float>(ym, mstep, y)[x]; }
    else map_backward<(KIND < 0 ? 0 : KIND)>(p, (float)(tl_x + x), (float)(tl_y + y), mx, my);
    const int fsx = cvround(__fmul_rn(mx, 32.f)), fsy = cvround(__fmul_rn(my, 32.f));
    const int sx = sat_s16(fsx >> 5), sy = sat_s16(fsy >> 5);
    const float d = crow<float>(dist, dstep, y)[x];
    unsigned di = d >= 65535.f ? 65535u : (unsigned)d;       // exact integers below the 8192 saturation
    const int x0 = border_interp<BORDER_REFLECT>(sx, sw), x1 = border_interp<BORDER_REFLECT>(sx + 1, sw);
    const int y0 = border_interp<BORDER_REFLECT>(sy, sh), y1 = border_interp<BORDER_REFLECT>(sy + 1, sh);
    const bool xs = x1 == x0, ys = y1 == y0;
    // representable: taps are (x0, x0 or x0+1) x (y0, y0 or y0+1); anything else cannot carry weight
    const bool ok = (xs || x1 == x0 + 1) && (ys || y1 == y0 + 1);
    uint2 t;
    t.x = (unsigned)x0 | ((unsigned)y0 << 13) | ((unsigned)xs << 26) | ((unsigned)ys << 27) | ((unsigned)!ok << 28);
    t.y = (unsigned)(fsx & 31) | ((unsigned)(fsy & 31) << 5) | (di << 16);
    reinterpret_cast<uint2 *>(reinterpret_cast<char *>(table) + (size_t)y * tstep)[x] = t;
}

int launch_build_feather_table(const ProjParams &p, int tl_x, int tl_y, const DImage &dist, int sw, int sh, uint2 *table, size_t tstep, cudaStream_t s,
                               const DImage *xmap, const DImage *ymap)
{
    SB_ASSERT(dist.type == SB_32FC1);
    SB_ASSERT(sw <= 8192 && sh <= 8192);
    dim3 block(32, 8), grid(div_up(dist.cols, 32), div_up(dist.rows, 8));
    const float *xm = xmap ? xmap->ptr<float>() : nullptr, *ym = ymap ? ymap->ptr<float>() : nullptr;
    const size_t mstep = xmap ? xmap->step : 0;
    if (xmap) SB_ASSERT(ymap && xmap->type == SB_32FC1 && ymap->type == SB_32FC1 && xmap->rows == dist.rows && xmap->cols == dist.cols && ymap->step == xmap->step);
#define SB_FT(K) k_build_feather_table<K><<<grid, block, 0, s>>>(p, tl_x, tl_y, dist.ptr<float>(), dist.step, dist.cols, dist.rows, sw, sh, table, tstep, xm, ym, mstep)
    if (xmap) SB_FT(-1);
    else switch (p.kind) {
    case SB_WARP_PLANE: SB_FT(SB_WARP_PLANE); break;
    case SB_WARP_CYLINDRICAL: SB_FT(SB_WARP_CYLINDRICAL); break;
    case SB_WARP_SPHERICAL: SB_FT(SB_WARP_SPHERICAL); break;
    default: return fail(SB_ERR_BAD_ARG, "projector kind %d needs host-built maps", p.kind);
    }
#undef SB_FT
    SB_LAUNCHED();
    return SB_OK;
}

// setup: tile_cams[tile] |= 1 << cam when camera `cam` has a non-zero feather distance inside the
// SB_FT_W x SB_FT_H panorama tile; *bad counts weighted entries whose taps are not representable.
__global__ void __launch_bounds__(256) k_feather_tile_cams(FeatherCam c, int cam_index, int pw, int ph, int tiles_x, uint32_t *tile_cams, unsigned *bad)
{
    const int X = blockIdx.x * SB_FT_W + threadIdx.x, Y = blockIdx.y * SB_FT_H + threadIdx.y;
    const int x = X - c.dx, y = Y - c.dy;
    bool nz = false;
    if (X < pw && Y < ph && (unsigned)x < (unsigned)c.ww && (unsigned)y < (unsigned)c.wh) {
        const uint2 t = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep)[x];
        nz = (t.y >> 16) != 0u;
        if (nz && ((t.x >> 28) & 1u)) atomicAdd(bad, 1u);
    }
    if (__syncthreads_or(nz) && threadIdx.x == 0 && threadIdx.y == 0) tile_cams[blockIdx.y * tiles_x + blockIdx.x] |= 1u << cam_index;
}

int launch_feather_tile_cams(const FeatherCam &c, int cam_index, int pw, int ph, uint32_t *tile_cams, unsigned *bad, cudaStream_t s)
{
    dim3 block(SB_FT_W, SB_FT_H), grid(div_up(pw, SB_FT_W), div_up(ph, SB_FT_H));
    k_feather_tile_cams<<<grid, block, 0, s>>>(c, cam_index, pw, ph, grid.x, tile_cams, bad);
    SB_LAUNCHED();
    return SB_OK;
}

// One launch per frame, one panorama pixel per thread: a warp's 32 lanes sample 32 neighbouring
// pixels, so each tap request touches a few 32-byte sectors.  One load fetches the tile's camera
// bitmask; the table entries of every contributing camera are requested before any is consumed, so
// the DRAM latencies of the cameras overlap.  With 8-bit sources and weights in [0,1] every
// intermediate stays inside [0, 255*n]: no saturation or wrap can fire.
template <bool GAIN, bool OUT8>
__global__ void __launch_bounds__(256)
k_feather_fused_px1(const __grid_constant__ FeatherFusedArgs a)
{
    const int X = blockIdx.x * SB_FT_W + threadIdx.x;
    const int Y = blockIdx.y * SB_FT_H + threadIdx.y;
    if (X >= a.pw || Y >= a.ph) return;
    uint32_t cams = __ldg(a.tile_cams + blockIdx.y * a.tiles_x + blockIdx.x);      // block == tile
    constexpr int MAXC = 4;
    uint2 t[MAXC];
    int ci[MAXC];
    int nc = 0;
#pragma unroll
    for (int q = 0; q < MAXC; ++q) {
        t[q] = make_uint2(0u, 0u);
        ci[q] = 0;
        if (cams) {
            const int i = __ffs(cams) - 1;
            cams &= cams - 1;
            const FeatherCam &c = a.cam[i];
            const int y = Y - c.dy, x = X - c.dx;
            ci[q] = i;
            nc = q + 1;
            if ((unsigned)y < (unsigned)c.wh && (unsigned)x < (unsigned)c.ww)
                t[q] = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + x);
        }
    }
    int acc0 = 0, acc1 = 0, acc2 = 0;
    float wsum = 0.f;
    auto contribute = [&](const FeatherCam &c, const uint2 te) {
        const unsigned dist = te.y >> 16;
        if (dist == 0u) return;                             // weight 0: short(p * 0) == 0 and dst_w += 0
        const uint2 bw = __ldg(a.bilin_lut + (te.y & 1023u));
        const float w = fminf(__fmul_rn((float)dist, a.sharpness), 1.f);      // createWeightMap
        wsum = __fadd_rn(wsum, w);
        const unsigned x0 = te.x & 0x1fffu, y0 = (te.x >> 13) & 0x1fffu;
        const uint8_t *r0 = c.src + (size_t)y0 * c.sstep;
        const uint8_t *r1 = (te.x & (1u << 27)) ? r0 : r0 + c.sstep;
        const unsigned x1 = x0 + 1u - ((te.x >> 26) & 1u);                    // x1 == x0 at the image edge
        unsigned lo0, hi0, lo1, hi1;
        load_tap_row(r0, x0, x1, lo0, hi0);
        load_tap_row(r1, x0, x1, lo1, hi1);
        int v0, v1, v2;
        bilinear_rgb(lo0, hi0, lo1, hi1, bw, v0, v1, v2);
        if (GAIN) {                                                           // saturate_cast<uchar>(p * gain)
            const float g = c.gmap ? __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.gmap) + (size_t)(Y - c.dy) * c.gmstep) + (X - c.dx)) : c.gain;
            v0 = min(max(__float2int_rn(__fmul_rn((float)v0, g)), 0), 255);
            v1 = min(max(__float2int_rn(__fmul_rn((float)v1, g)), 0), 255);
            v2 = min(max(__float2int_rn(__fmul_rn((float)v2, g)), 0), 255);
        }
        acc0 += __float2int_rz(__fmul_rn((float)v0, w));                      // static_cast<short>(src * w)
        acc1 += __float2int_rz(__fmul_rn((float)v1, w));
        acc2 += __float2int_rz(__fmul_rn((float)v2, w));
    };
#pragma unroll
    for (int q = 0; q < MAXC; ++q)
        if (q < nc) contribute(a.cam[ci[q]], t[q]);
    for (; cams; cams &= cams - 1) {                        // more than MAXC overlapping cameras: rest in order
        const FeatherCam &c = a.cam[__ffs(cams) - 1];
        const int y = Y - c.dy, x = X - c.dx;
        if ((unsigned)y < (unsigned)c.wh && (unsigned)x < (unsigned)c.ww)
            contribute(c, __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + x));
    }
    // FeatherBlender::blend: normalizeUsingWeightMap, mask = weight > eps, zero unmasked, convertTo(8U)
    const int m = wsum > SB_WEIGHT_EPS ? 255 : 0;
    const SharedDiv div(__fadd_rn(wsum, SB_WEIGHT_EPS));    // in [1e-5, n + 1e-5]: fast-path range
    const int o0 = m ? __float2int_rz(div((float)acc0)) : 0, o1 = m ? __float2int_rz(div((float)acc1)) : 0,
              o2 = m ? __float2int_rz(div((float)acc2)) : 0;
    if (OUT8) {
        uint8_t *o = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X * 3;
        o[0] = (uint8_t)o0; o[1] = (uint8_t)o1; o[2] = (uint8_t)o2;
    } else {
        short *o = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X * 3;
        o[0] = (short)o0; o[1] = (short)o1; o[2] = (short)o2;
    }
    if (a.out_mask) a.out_mask[(size_t)Y * a.mask_step + X] = (uint8_t)m;
}

int launch_feather_fused(const FeatherFusedArgs &a, bool apply_gain, bool out8, cudaStream_t s)
{
    SB_ASSERT(a.sharpness > 0.f && a.bilin_lut && a.tile_cams);
    SB_ASSERT(div_up(a.pw, SB_FT_W) == a.tiles_x);
    dim3 block(SB_FT_W, SB_FT_H), grid(div_up(a.pw, SB_FT_W), div_up(a.ph, SB_FT_H));
    if (apply_gain) { if (out8) k_feather_fused_px1<true, true><<<grid, block, 0, s>>>(a); else k_feather_fused_px1<true, false><<<grid, block, 0, s>>>(a); }
    else            { if (out8) k_feather_fused_px1<false, true><<<grid, block, 0, s>>>(a); else k_feather_fused_px1<false, false><<<grid, block, 0, s>>>(a); }
    SB_LAUNCHED();
    return SB_OK;
}

// device self-test: SharedDiv vs __fdiv_rn over pseudo-random operands in the ranges the kernels use
__global__ void k_selftest_division(unsigned long long n, unsigned seed, unsigned long long *bad)
{
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        unsigned g = h * 747796405u + 2891336453u;
        g ^= g >> 16; g *= 2246822519u; g ^= g >> 13;
        // divisor: any float with exponent in the accepted range; numerator: integers and general floats
        float d = __uint_as_float(((87u + (h % 81u)) << 23) | (g & 0x7fffffu));
        float a_ = (i & 1) ? (float)((int)(g >> 8) % 65536 - 32768) : __uint_as_float((g & 0x80000000u) | ((87u + ((g >> 3) % 81u)) << 23) | (h & 0x7fffffu));
        if (!div_fast_ok(d, false) || !div_fast_ok(a_, true)) continue;
        SharedDiv sd(d);
        if (__float_as_uint(sd(a_)) != __float_as_uint(__fdiv_rn(a_, d))) ++local;
    }
    if (local) atomicAdd(bad, local);
}

int selftest_division(unsigned long long n, unsigned seed, unsigned long long *mismatches)
{
    unsigned long long *d_bad = nullptr;
    SB_CUDA(cudaMalloc(&d_bad, sizeof *d_bad));
    SB_CUDA(cudaMemset(d_bad, 0, sizeof *d_bad));
    k_selftest_division<<<148 * 8, 256>>>(n, seed, d_bad);
    SB_LAUNCHED();
    SB_CUDA(cudaMemcpy(mismatches, d_bad, sizeof *d_bad, cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    return SB_OK;
}

// ------------------------------------------------------------------------------------ multi-band
__device__ __forceinline__ int band_weighted(int lap, float w) { return (int)trunc_short(__fmul_rn((float)lap, w)); }
__device__ __forceinline__ int band_weighted(int lap, short w) { return (int)(short)((lap * (int)w) >> 8); }
__device__ __forceinline__ int band_normalized(short p, float w) { return trunc_short(__fdiv_rn((float)p, __fadd_rn(w, SB_WEIGHT_EPS))); }
__device__ __forceinline__ int band_normalized(short p, short w)
{
    const int wi = (int)w + 1;
    return wi == 0 ? 0 : (int)(short)((((int)p) << 8) / wi);
}
__device__ __forceinline__ bool band_masked(float w) { return w > SB_WEIGHT_EPS; }
__device__ __forceinline__ bool band_masked(short w) { return w > 0; }

// One band l for all cameras.  HAS_FINER_SRC: the cameras have a coarser Gaussian level (l < n);
// HAS_COARSE_R: a restored coarser band exists to be added (l < n); FINAL: l == 0.
template <typename WT, bool NOT_TOP, bool FINAL, bool OUT8>
__global__ void __launch_bounds__(128)
k_band_fused(const __grid_constant__ BandFusedArgs a)
{
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    const int lw = FINAL ? a.out_w : a.lw, lh = FINAL ? a.out_h : a.lh;    // band 0 is cropped to dst_roi_final_
    if (X0 >= lw || Y >= lh) return;
    int acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0;

    for (int i = 0; i < a.n; ++i) {
        const BandCam &c = a.cam[i];
        const int y = Y - c.ry;
        if ((unsigned)y >= (unsigned)c.rh) continue;
        if (!in_spans(c.span, X0, X0 + 4)) continue;
        const WT *wrow = crow<WT>(c.weight, c.wstep, y);
        const short *frow = crow<short>(c.fine, c.fstep, y);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = X0 + j - c.rx;
            if ((unsigned)x >= (unsigned)c.rw) continue;
            const WT w = wrow[x];
            if (w == (WT)0) continue;                     // short(lap * 0) == 0 and (lap * 0) >> 8 == 0
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int lap = frow[x * 3 + k];
                if (NOT_TOP) {
                    const int up = up_cast<short>(pyr_up_sum<short, 3>(c.coarse, c.cstep, c.rw >> 1, c.rh >> 1, y, x, k));
                    lap = sat_s16(lap - up);              // subtract(pyr[i], tmp, pyr[i]) saturates
                }
                acc[j][k] += band_weighted(lap, w);
            }
        }
    }
    const WT *ws = crow<WT>(a.wsum, a.wsum_step, Y) + X0;
    int o[4][3], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool in = X0 + j < lw;
        const WT w = in ? ws[j] : (WT)0;
        m[j] = band_masked(w) ? 255 : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int v = band_normalized((short)acc[j][k], w);
            if (NOT_TOP && in)   // restoreImageFromLaplacePyr: add(pyrUp(pyr[i+1]), pyr[i]) saturates
                v = sat_s16(up_cast<short>(pyr_up_sum<short, 3>(a.coarse_r, a.coarse_r_step, a.lw >> 1, a.lh >> 1, Y, X0 + j, k)) + v);
            if (FINAL) { v = m[j] ? v : 0; if (OUT8) v = sat_u8(v); }
            o[j][k] = v;
        }
    }
    if (FINAL && OUT8) {
        uint8_t *orow = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X0 * 3;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(orow) & 3) == 0)) {
            uint32_t *q = reinterpret_cast<uint32_t *>(orow);
            q[0] = o[0][0] | (o[0][1] << 8) | (o[0][2] << 16) | (o[1][0] << 24);
            q[1] = o[1][1] | (o[1][2] << 8) | (o[2][0] << 16) | (o[2][1] << 24);
            q[2] = o[2][2] | (o[3][0] << 8) | (o[3][1] << 16) | (o[3][2] << 24);
        } else {
            for (int j = 0; j < 4 && X0 + j < lw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (uint8_t)o[j][k];
        }
    } else {
        short *orow = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X0 * 3;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(orow) & 7) == 0)) {
            uint2 *q = reinterpret_cast<uint2 *>(orow);
            auto pk = [](int lo, int hi) { return (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16); };
            q[0] = make_uint2(pk(o[0][0], o[0][1]), pk(o[0][2], o[1][0]));
            q[1] = make_uint2(pk(o[1][1], o[1][2]), pk(o[2][0], o[2][1]));
            q[2] = make_uint2(pk(o[2][2], o[3][0]), pk(o[3][1], o[3][2]));
        } else {
            for (int j = 0; j < 4 && X0 + j < lw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (short)o[j][k];
        }
    }
    if (FINAL && a.out_mask) {
        uint8_t *mrow_ = a.out_mask + (size_t)Y * a.mask_step + X0;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(mrow_) & 3) == 0))
            *reinterpret_cast<uint32_t *>(mrow_) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
        else
            for (int j = 0; j < 4 && X0 + j < lw; ++j) mrow_[j] = (uint8_t)m[j];
    }
}

int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s)
{
    const int lw = final_band ? a.out_w : a.lw, lh = final_band ? a.out_h : a.lh;
    dim3 block(32, 4), grid(div_up(div_up(lw, 4), 32), div_up(lh, 4));
#define SB_BF(WT, NT, FIN, O8) k_band_fused<WT, NT, FIN, O8><<<grid, block, 0, s>>>(a)
#define SB_BF_W(WT)                                                                          \
    do {                                                                                     \
        if (final_band) { if (not_top) { if (out8) SB_BF(WT, true, true, true); else SB_BF(WT, true, true, false); } \
                          else { if (out8) SB_BF(WT, false, true, true); else SB_BF(WT, false, true, false); } }     \
        else { if (not_top) SB_BF(WT, true, false, false); else SB_BF(WT, false, false, false); }                    \
    } while (0)
    if (float_weights) SB_BF_W(float); else SB_BF_W(short);
#undef SB_BF_W
#undef SB_BF
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ spans
// flags[x] = 1 if column x of the weight image holds any non-zero weight (setup time only)
template <typename WT> __global__ void k_column_nonzero(const WT *w, size_t wstep, int cols, int rows, int *flags)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cols) return;
    int any = 0;
    for (int y = 0; y < rows && !any; ++y) any |= crow<WT>(w, wstep, y)[x] != (WT)0;
    flags[x] = any;
}

int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s)
{
    SB_ASSERT(w.type == SB_32FC1 || w.type == SB_16SC1);
    if (w.type == SB_32FC1) k_column_nonzero<float><<<div_up(w.cols, 128), 128, 0, s>>>(w.ptr<float>(), w.step, w.cols, w.rows, flags);
    else k_column_nonzero<short><<<div_up(w.cols, 128), 128, 0, s>>>(w.ptr<short>(), w.step, w.cols, w.rows, flags);
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
