// kernels_mb.cu — the compositor's multi-band fast path.
//
// Same arithmetic as MultiBandBlender::feed/blend (blenders.cpp:236-377) on the warped + gain-applied
// images, restructured for the GPU:
//   * After warp + gain every pixel is 8-bit and pyrDown keeps it 8-bit ((s+128)>>8 of a weighted
//     mean), so the cameras' Gaussian pyramids are stored as RGBX bytes (one aligned 32-bit word per
//     pixel) instead of CV_16SC3: a third less traffic, one load per tap, and the 5x5 / 3x3 filter
//     sums run on two colour channels at once in 16-bit lanes of a 32-bit register (the largest
//     intermediate, 255 * 256, still fits).  The values are identical to the CV_16S pipeline.
//   * cv::remap's per-frame map conversion (A1) and the BORDER_REFLECT of both remap and
//     copyMakeBorder are resolved once per calibration into an 8-byte tap table per padded pixel.
//   * The bands are produced panorama-centric, coarse to fine (see kernels_fused.cu): Laplacian on the
//     fly, weighted sum over the cameras that have non-zero weight in the tile, normalise, collapse.
// One thread per pixel with the warp's lanes on neighbouring pixels keeps every tap request inside
// a few 32-byte sectors; per-tile camera bitmasks keep the per-pixel camera loop to the 1-3 cameras
// that matter; all of a pixel's independent loads are issued before the first is consumed.
#include <climits>

#include "sb_device.cuh"
#include "sb_mb.h"
#include "sb_warp.cuh"

namespace sb {
using namespace sbd;

#define SB_WEIGHT_EPS 1e-5f

template <typename T> __device__ __forceinline__ const T *rowp(const void *base, size_t step, int y)
{
    return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) + (size_t)y * step);
}
template <typename T> __device__ __forceinline__ T *rowp(void *base, size_t step, int y)
{
    return reinterpret_cast<T *>(reinterpret_cast<char *>(base) + (size_t)y * step);
}

// ------------------------------------------------------------------------------------ setup: tap table
// entry.x = x0 | x1 << 12 | fx << 24      entry.y = y0 | y1 << 12 | fy << 24   (source size <= 4096)
template <int KIND>
__global__ void __launch_bounds__(256)
k_mb_tap_table(ProjParams p, int tl_x, int tl_y, int ww, int wh, int left, int top, int sw, int sh, uint2 *table, size_t tstep, int rw, int rh)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= rw || py >= rh) return;
    // copyMakeBorder(BORDER_REFLECT) of the warped image (blenders.cpp:272-274)
    const int wx = border_interp<BORDER_REFLECT>(px - left, ww), wy = border_interp<BORDER_REFLECT>(py - top, wh);
    float mx, my;
    map_backward<KIND>(p, (float)(tl_x + wx), (float)(tl_y + wy), mx, my);
    const int fsx = cvround(__fmul_rn(mx, 32.f)), fsy = cvround(__fmul_rn(my, 32.f));
    const int sx = sat_s16(fsx >> 5), sy = sat_s16(fsy >> 5);
    const unsigned x0 = border_interp<BORDER_REFLECT>(sx, sw), x1 = border_interp<BORDER_REFLECT>(sx + 1, sw);
    const unsigned y0 = border_interp<BORDER_REFLECT>(sy, sh), y1 = border_interp<BORDER_REFLECT>(sy + 1, sh);
    uint2 t;
    t.x = x0 | (x1 << 12) | ((unsigned)(fsx & 31) << 24);
    t.y = y0 | (y1 << 12) | ((unsigned)(fsy & 31) << 24);
    rowp<uint2>(table, tstep, py)[px] = t;
}

int launch_mb_tap_table(const ProjParams &p, int tl_x, int tl_y, int ww, int wh, int left, int top, int sw, int sh,
                        uint2 *table, size_t tstep, int rw, int rh, cudaStream_t s)
{
    SB_ASSERT(sw <= 4096 && sh <= 4096);
    dim3 block(32, 8), grid(div_up(rw, 32), div_up(rh, 8));
#define SB_TT(K) k_mb_tap_table<K><<<grid, block, 0, s>>>(p, tl_x, tl_y, ww, wh, left, top, sw, sh, table, tstep, rw, rh)
    switch (p.kind) {
    case SB_WARP_PLANE: SB_TT(SB_WARP_PLANE); break;
    case SB_WARP_CYLINDRICAL: SB_TT(SB_WARP_CYLINDRICAL); break;
    case SB_WARP_SPHERICAL: SB_TT(SB_WARP_SPHERICAL); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", p.kind);
    }
#undef SB_TT
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ setup: tile masks
// mask[tile] bit i set when camera i has a non-zero weight inside the 32x8 tile of band l
template <typename WT>
__global__ void __launch_bounds__(256) k_mb_tile_mask(MbBandGeom g, uint32_t *mask, int tiles_x)
{
    const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y;
    uint32_t m = 0;
    for (int i = 0; i < g.n; ++i) {
        const int x = X - g.cam[i].rx, y = Y - g.cam[i].ry;
        bool nz = false;
        if ((unsigned)x < (unsigned)g.cam[i].rw && (unsigned)y < (unsigned)g.cam[i].rh)
            nz = rowp<WT>(g.cam[i].weight, g.cam[i].wstep, y)[x] != (WT)0;
        if (__syncthreads_or(nz)) m |= 1u << i;
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) mask[blockIdx.y * tiles_x + blockIdx.x] = m;
}

int launch_mb_tile_mask(const MbBandGeom &g, bool float_weights, int lw, int lh, uint32_t *mask, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(lw, 32), div_up(lh, 8));
    if (float_weights) k_mb_tile_mask<float><<<grid, block, 0, s>>>(g, mask, grid.x);
    else k_mb_tile_mask<short><<<grid, block, 0, s>>>(g, mask, grid.x);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K1: warp -> G0 (RGBX)
// remap (A1) + GainCompensator::apply + convertTo(CV_16S) + copyMakeBorder, all cameras in one launch
// (blockIdx.z = camera).  (sum w*p + 2^14) >> 15 == (v + 512) >> 10 with the 5-bit weights; OpenCV's
// (0,0) table entry {32767,0,0,1} equals an exact copy for 8-bit data, as does {32768,0,0,0}.
template <bool GAIN>
__global__ void __launch_bounds__(256) k_mb_warp(const __grid_constant__ MbWarpArgs a)
{
    const MbWarpCam &c = a.cam[blockIdx.z];
    const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
    if (px >= c.rw || py >= c.rh) return;
    const uint2 t = __ldg(rowp<uint2>(c.table, c.tstep, py) + px);
    const int x0 = t.x & 0xfff, x1 = (t.x >> 12) & 0xfff, fx = t.x >> 24, ax = 32 - fx;
    const int y0 = t.y & 0xfff, y1 = (t.y >> 12) & 0xfff, fy = t.y >> 24, ay = 32 - fy;
    const uint8_t *r0 = c.src + (size_t)y0 * c.sstep, *r1 = c.src + (size_t)y1 * c.sstep;
    const uint8_t *p00 = r0 + x0 * 3, *p01 = r0 + x1 * 3, *p10 = r1 + x0 * 3, *p11 = r1 + x1 * 3;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int h0 = (int)__ldg(p00 + k) * ax + (int)__ldg(p01 + k) * fx;
        const int h1 = (int)__ldg(p10 + k) * ax + (int)__ldg(p11 + k) * fx;
        int v = (h0 * ay + h1 * fy + 512) >> 10;
        if (GAIN) v = min(max(__float2int_rn(__fmul_rn((float)v, c.gain)), 0), 255);      // saturate_cast<uchar>
        out |= (unsigned)v << (8 * k);
    }
    rowp<uint32_t>(c.g0, c.gstep, py)[px] = out;
}

int launch_mb_warp(const MbWarpArgs &a, bool apply_gain, int max_rw, int max_rh, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(max_rw, 32), div_up(max_rh, 8), a.n);
    if (apply_gain) k_mb_warp<true><<<grid, block, 0, s>>>(a); else k_mb_warp<false><<<grid, block, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K2: pyrDown on RGBX
// channels 0 and 2 ride in the 16-bit lanes of one register, channel 1 in another
__device__ __forceinline__ unsigned lanes02(unsigned v) { return v & 0x00ff00ffu; }
__device__ __forceinline__ unsigned lane1(unsigned v) { return (v >> 8) & 0xffu; }

__global__ void __launch_bounds__(256) k_mb_pyr_down(const __grid_constant__ MbPyrArgs a)
{
    const MbPyrCam &c = a.cam[blockIdx.z];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int dw = (c.sw + 1) >> 1, dh = (c.sh + 1) >> 1;
    if (x >= dw || y >= dh) return;
    int xs[5];
    const int cx = 2 * x;
    if (cx >= 2 && cx + 2 < c.sw) { xs[0] = cx - 2; xs[1] = cx - 1; xs[2] = cx; xs[3] = cx + 1; xs[4] = cx + 2; }
    else {
#pragma unroll
        for (int j = 0; j < 5; ++j) xs[j] = reflect101(cx + j - 2, c.sw);
    }
    unsigned v[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t *row = rowp<uint32_t>(c.src, c.sstep, reflect101(2 * y + i - 2, c.sh));
#pragma unroll
        for (int j = 0; j < 5; ++j) v[i][j] = __ldg(row + xs[j]);
    }
    unsigned h02[5], h1[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {       // row = s2*6 + (s1+s3)*4 + s0 + s4, <= 16*255 per lane
        h02[i] = lanes02(v[i][2]) * 6u + (lanes02(v[i][1]) + lanes02(v[i][3])) * 4u + lanes02(v[i][0]) + lanes02(v[i][4]);
        h1[i] = lane1(v[i][2]) * 6u + (lane1(v[i][1]) + lane1(v[i][3])) * 4u + lane1(v[i][0]) + lane1(v[i][4]);
    }
    // <= 256*255 = 65280 per lane: no carry between the lanes; (s + 128) >> 8 per lane
    const unsigned s02 = h02[2] * 6u + (h02[1] + h02[3]) * 4u + h02[0] + h02[4] + 0x00800080u;
    const unsigned s1 = h1[2] * 6u + (h1[1] + h1[3]) * 4u + h1[0] + h1[4] + 128u;
    // lane sums can reach 65280 + 128 = 65408 < 65536
    const unsigned out = ((s02 >> 8) & 0x00ff00ffu) | ((s1 >> 8) << 8);
    rowp<uint32_t>(c.dst, c.dstep, y)[x] = out;
}

int launch_mb_pyr_down(const MbPyrArgs &a, int max_dw, int max_dh, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(max_dw, 32), div_up(max_dh, 8), a.n);
    k_mb_pyr_down<<<grid, block, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K3: one band
// pyrUp(coarse)(y, x) for RGBX bytes, channels {0,2} packed + channel 1; value before the cast (<= 64*255)
struct Up3 { unsigned s02, s1; };
__device__ __forceinline__ Up3 pyr_up_rgbx(const uint32_t *coarse, size_t cstep, int cw, int ch, int y, int x)
{
    const int cx = x >> 1, cy = y >> 1;
    const int xl = cx == 0 ? (cw > 1 ? 1 : 0) : cx - 1, xr = min(cx + 1, cw - 1);
    const int yt = cy == 0 ? (ch > 1 ? 1 : 0) : cy - 1, yb = min(cy + 1, ch - 1);
    const uint32_t *r0 = rowp<uint32_t>(coarse, cstep, yt), *r1 = rowp<uint32_t>(coarse, cstep, cy), *r2 = rowp<uint32_t>(coarse, cstep, yb);
    const unsigned a0 = __ldg(r0 + xl), b0 = __ldg(r0 + cx), c0 = __ldg(r0 + xr);
    const unsigned a1 = __ldg(r1 + xl), b1 = __ldg(r1 + cx), c1 = __ldg(r1 + xr);
    const unsigned a2 = __ldg(r2 + xl), b2 = __ldg(r2 + cx), c2 = __ldg(r2 + xr);
    unsigned h0_02, h1_02, h2_02, h0_1, h1_1, h2_1;
    if (x & 1) {
        h0_02 = (lanes02(b0) + lanes02(c0)) * 4u; h1_02 = (lanes02(b1) + lanes02(c1)) * 4u; h2_02 = (lanes02(b2) + lanes02(c2)) * 4u;
        h0_1 = (lane1(b0) + lane1(c0)) * 4u; h1_1 = (lane1(b1) + lane1(c1)) * 4u; h2_1 = (lane1(b2) + lane1(c2)) * 4u;
    } else {
        h0_02 = lanes02(a0) + lanes02(b0) * 6u + lanes02(c0); h1_02 = lanes02(a1) + lanes02(b1) * 6u + lanes02(c1); h2_02 = lanes02(a2) + lanes02(b2) * 6u + lanes02(c2);
        h0_1 = lane1(a0) + lane1(b0) * 6u + lane1(c0); h1_1 = lane1(a1) + lane1(b1) * 6u + lane1(c1); h2_1 = lane1(a2) + lane1(b2) * 6u + lane1(c2);
    }
    Up3 u;
    if (y & 1) { u.s02 = (h1_02 + h2_02) * 4u; u.s1 = (h1_1 + h2_1) * 4u; }
    else { u.s02 = h1_02 * 6u + h0_02 + h2_02; u.s1 = h1_1 * 6u + h0_1 + h2_1; }
    return u;
}

__device__ __forceinline__ int mb_weighted(int lap, float w) { return __float2int_rz(__fmul_rn((float)lap, w)); }   // |lap*w| < 2^31: no x86 indefinite
__device__ __forceinline__ int mb_weighted(int lap, short w) { return (int)(short)((lap * (int)w) >> 8); }

// restored coarser band (CV_16SC4 storage: 3 channels + pad), pyrUp value before the cast, per channel
__device__ __forceinline__ void pyr_up_s16x4(const short4 *coarse, size_t cstep, int cw, int ch, int y, int x, int up[3])
{
    const int cx = x >> 1, cy = y >> 1;
    const int xl = cx == 0 ? (cw > 1 ? 1 : 0) : cx - 1, xr = min(cx + 1, cw - 1);
    const int yt = cy == 0 ? (ch > 1 ? 1 : 0) : cy - 1, yb = min(cy + 1, ch - 1);
    const short4 *r0 = rowp<short4>(coarse, cstep, yt), *r1 = rowp<short4>(coarse, cstep, cy), *r2 = rowp<short4>(coarse, cstep, yb);
    const short4 a0 = __ldg(r0 + xl), b0 = __ldg(r0 + cx), c0 = __ldg(r0 + xr);
    const short4 a1 = __ldg(r1 + xl), b1 = __ldg(r1 + cx), c1 = __ldg(r1 + xr);
    const short4 a2 = __ldg(r2 + xl), b2 = __ldg(r2 + cx), c2 = __ldg(r2 + xr);
#define SB_H(A, B, C, F) ((x & 1) ? ((int)B.F + (int)C.F) * 4 : (int)A.F + (int)B.F * 6 + (int)C.F)
#define SB_V(F) ((y & 1) ? (SB_H(a1, b1, c1, F) + SB_H(a2, b2, c2, F)) * 4 : SB_H(a1, b1, c1, F) * 6 + SB_H(a0, b0, c0, F) + SB_H(a2, b2, c2, F))
    up[0] = SB_V(x); up[1] = SB_V(y); up[2] = SB_V(z);
#undef SB_V
#undef SB_H
}

template <typename WT, bool NOT_TOP, bool FINAL, bool OUT8>
__global__ void __launch_bounds__(256) k_mb_band(const __grid_constant__ MbBandArgs a)
{
    const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y;
    const int lw = FINAL ? a.out_w : a.g.lw, lh = FINAL ? a.out_h : a.g.lh;       // band 0 is cropped to dst_roi_final_
    if (X >= lw || Y >= lh) return;
    uint32_t cams = __ldg(a.tile_mask + blockIdx.y * a.tiles_x + blockIdx.x);     // block-uniform
    const WT ws = __ldg(rowp<WT>(a.wsum, a.wsum_step, Y) + X);
    int up_r[3] = {0, 0, 0};
    if (NOT_TOP) pyr_up_s16x4(a.coarse_r, a.coarse_r_step, a.g.lw >> 1, a.g.lh >> 1, Y, X, up_r);
    int acc0 = 0, acc1 = 0, acc2 = 0;
    for (; cams; cams &= cams - 1) {
        const MbBandCam &c = a.g.cam[__ffs(cams) - 1];
        const int x = X - c.rx, y = Y - c.ry;
        if ((unsigned)x >= (unsigned)c.rw || (unsigned)y >= (unsigned)c.rh) continue;
        const WT w = __ldg(rowp<WT>(c.weight, c.wstep, y) + x);
        const unsigned g = __ldg(rowp<uint32_t>(c.fine, c.fstep, y) + x);
        int l0 = g & 0xff, l1 = (g >> 8) & 0xff, l2 = (g >> 16) & 0xff;
        if (NOT_TOP) {
            const Up3 u = pyr_up_rgbx(c.coarse, c.cstep, c.rw >> 1, c.rh >> 1, y, x);
            const unsigned u02 = ((u.s02 + 0x00200020u) >> 6) & 0x03ff03ffu;      // (v + 32) >> 6 per lane, <= 255
            l0 -= (int)(u02 & 0xffffu); l2 -= (int)(u02 >> 16); l1 -= (int)((u.s1 + 32u) >> 6);
        }
        if (w == (WT)0) continue;                          // short(lap * 0) == 0 and (lap * 0) >> 8 == 0
        acc0 += mb_weighted(l0, w); acc1 += mb_weighted(l1, w); acc2 += mb_weighted(l2, w);
    }
    // normalizeUsingWeightMap (blenders.cpp:383-424) on the wrapped 16-bit sums
    int v0, v1, v2;
    bool masked;
    if (sizeof(WT) == 4) {
        const float wf = (float)ws;
        masked = wf > SB_WEIGHT_EPS;
        const SharedDiv div(__fadd_rn(wf, SB_WEIGHT_EPS));          // [1e-5, n + 1e-5]: fast-path range
        v0 = trunc_short(div((float)(short)acc0)); v1 = trunc_short(div((float)(short)acc1)); v2 = trunc_short(div((float)(short)acc2));
    } else {
        const int wi = (int)ws + 1;
        masked = (int)ws > 0;
        v0 = wi ? (short)((((int)(short)acc0) << 8) / wi) : 0; v1 = wi ? (short)((((int)(short)acc1) << 8) / wi) : 0;
        v2 = wi ? (short)((((int)(short)acc2) << 8) / wi) : 0;
    }
    if (NOT_TOP) {   // restoreImageFromLaplacePyr: add(pyrUp(pyr[i+1]), pyr[i]) saturates
        v0 = sat_s16(sat_s16((up_r[0] + 32) >> 6) + v0); v1 = sat_s16(sat_s16((up_r[1] + 32) >> 6) + v1);
        v2 = sat_s16(sat_s16((up_r[2] + 32) >> 6) + v2);
    }
    if (FINAL) {
        if (!masked) v0 = v1 = v2 = 0;                     // Blender::blend: dst_.setTo(0, dst_mask_ == 0)
        if (OUT8) {
            uint8_t *o = rowp<uint8_t>(a.out, a.out_step, Y) + X * 3;
            o[0] = (uint8_t)sat_u8(v0); o[1] = (uint8_t)sat_u8(v1); o[2] = (uint8_t)sat_u8(v2);   // result.convertTo(CV_8U)
        } else {
            short *o = rowp<short>(a.out, a.out_step, Y) + X * 3;
            o[0] = (short)v0; o[1] = (short)v1; o[2] = (short)v2;
        }
        if (a.out_mask) a.out_mask[(size_t)Y * a.mask_step + X] = masked ? 255 : 0;
    } else {
        rowp<short4>(a.out, a.out_step, Y)[X] = make_short4((short)v0, (short)v1, (short)v2, 0);
    }
}

int launch_mb_band(const MbBandArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s)
{
    const int lw = final_band ? a.out_w : a.g.lw, lh = final_band ? a.out_h : a.g.lh;
    dim3 block(32, 8), grid(div_up(lw, 32), div_up(lh, 8));
    SB_ASSERT(div_up(a.g.lw, 32) == a.tiles_x);
#define SB_MB(WT, NT, FIN, O8) k_mb_band<WT, NT, FIN, O8><<<grid, block, 0, s>>>(a)
#define SB_MB_W(WT)                                                                          \
    do {                                                                                     \
        if (final_band) { if (not_top) { if (out8) SB_MB(WT, true, true, true); else SB_MB(WT, true, true, false); } \
                          else { if (out8) SB_MB(WT, false, true, true); else SB_MB(WT, false, true, false); } }     \
        else { if (not_top) SB_MB(WT, true, false, false); else SB_MB(WT, false, false, false); }                    \
    } while (0)
    if (float_weights) SB_MB_W(float); else SB_MB_W(short);
#undef SB_MB_W
#undef SB_MB
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
