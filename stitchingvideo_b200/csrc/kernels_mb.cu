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
#include <algorithm>
#include <cstdlib>
#include <climits>

#include <cooperative_groups.h>

#include "sb_device.cuh"
#include "sb_mb.h"
#include "sb_warp.cuh"
#include "sb_tma.cuh"

namespace sb {
using namespace sbd;
namespace cg = cooperative_groups;

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
template <int KIND>      // KIND < 0: the warped image's maps come from xm / ym (projectors evaluated on the host)
__global__ void __launch_bounds__(256)
k_mb_tap_table(ProjParams p, int tl_x, int tl_y, int ww, int wh, int left, int top, int sw, int sh, uint2 *table, size_t tstep, int rw, int rh,
               const float *xm, const float *ym, size_t mstep)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= rw || py >= rh) return;
    // copyMakeBorder(BORDER_REFLECT) of the warped image (blenders.cpp:272-274)
    const int wx = border_interp<BORDER_REFLECT>(px - left, ww), wy = border_interp<BORDER_REFLECT>(py - top, wh);
    float mx, my;
    if (KIND < 0) { mx = rowp<float>(xm, mstep, wy)[wx]; my = rowp<float>(ym, mstep, wy)[wx]; }
    else map_backward<(KIND < 0 ? 0 : KIND)>(p, (float)(tl_x + wx), (float)(tl_y + wy), mx, my);
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
                        uint2 *table, size_t tstep, int rw, int rh, cudaStream_t s, const DImage *xmap, const DImage *ymap)
{
    SB_ASSERT(sw <= 4096 && sh <= 4096);
    dim3 block(32, 8), grid(div_up(rw, 32), div_up(rh, 8));
    const float *xm = xmap ? xmap->ptr<float>() : nullptr, *ym = ymap ? ymap->ptr<float>() : nullptr;
    const size_t mstep = xmap ? xmap->step : 0;
    if (xmap) SB_ASSERT(ymap && xmap->rows == wh && xmap->cols == ww && ymap->step == xmap->step);
#define SB_TT(K) k_mb_tap_table<K><<<grid, block, 0, s>>>(p, tl_x, tl_y, ww, wh, left, top, sw, sh, table, tstep, rw, rh, xm, ym, mstep)
    if (xmap) SB_TT(-1);
    else switch (p.kind) {
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
// mask[tile] bit i set when camera i has a non-zero weight inside the 32x8 tile of band l.  Bit 31 (float weights):
// the tile is "plain" — exactly one camera carries weight, that weight is exactly 1.0f on every pixel of the tile and
// so is the weight sum; the band kernel then needs no float arithmetic there (see mb_band_thread).
template <typename WT>
__global__ void __launch_bounds__(256) k_mb_tile_mask(MbBandGeom g, const WT *wsum, size_t wsum_step, uint32_t *mask, int tiles_x)
{
    const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y;
    const bool inside = X < g.lw && Y < g.lh;
    uint32_t m = 0;
    int n_nz = 0;
    bool one = inside;
    for (int i = 0; i < g.n; ++i) {
        const int x = X - g.cam[i].rx, y = Y - g.cam[i].ry;
        bool nz = false;
        if (inside && (unsigned)x < (unsigned)g.cam[i].rw && (unsigned)y < (unsigned)g.cam[i].rh) {
            const WT w = rowp<WT>(g.cam[i].weight, g.cam[i].wstep, y)[x];
            nz = w != (WT)0;
            if (nz) { ++n_nz; one = one && sizeof(WT) == 4 && (float)w == 1.f; }
        }
        if (__syncthreads_or(nz)) m |= 1u << i;
    }
    one = one && n_nz == 1 && (float)rowp<WT>(wsum, wsum_step, inside ? Y : 0)[inside ? X : 0] == 1.f;
    const bool plain = __syncthreads_and(one) && __popc(m) == 1 && sizeof(WT) == 4;
    if (threadIdx.x == 0 && threadIdx.y == 0) mask[blockIdx.y * tiles_x + blockIdx.x] = m | (plain ? 0x80000000u : 0u);
}

int launch_mb_tile_mask(const MbBandGeom &g, bool float_weights, int lw, int lh, const void *wsum, size_t wsum_step, uint32_t *mask, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(lw, 32), div_up(lh, 8));
    if (float_weights) k_mb_tile_mask<float><<<grid, block, 0, s>>>(g, static_cast<const float *>(wsum), wsum_step, mask, grid.x);
    else k_mb_tile_mask<short><<<grid, block, 0, s>>>(g, static_cast<const short *>(wsum), wsum_step, mask, grid.x);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K1: warp -> G0 (RGBX)
// remap (A1) + GainCompensator::apply + convertTo(CV_16S) + copyMakeBorder, all cameras in one launch
// (blockIdx.z = camera), with the byte-dot-product bilinear core of sb_device.cuh.
// Each thread produces MB_WARP_ROWS pixels (same column, rows 8 apart): all their table entries are
// requested first, then all 12*ROWS byte taps, so several DRAM round trips overlap per thread.
constexpr int MB_WARP_ROWS = 4;

template <bool GAIN>
__global__ void __launch_bounds__(256) k_mb_warp(const __grid_constant__ MbWarpArgs a)
{
    const MbWarpCam &c = a.cam[blockIdx.z];
    const int px = blockIdx.x * 32 + threadIdx.x, py0 = blockIdx.y * (8 * MB_WARP_ROWS) + threadIdx.y;
    if (py0 >= c.rh || !((px >= c.cx[0] && px < c.cx[1]) || (px >= c.cx[2] && px < c.cx[3]))) return;
    uint2 t[MB_WARP_ROWS];
#pragma unroll
    for (int r = 0; r < MB_WARP_ROWS; ++r) {
        const int py = min(py0 + 8 * r, c.rh - 1);
        t[r] = __ldg(rowp<uint2>(c.table, c.tstep, py) + px);
    }
    unsigned tap[MB_WARP_ROWS][4];                          // lo/hi of tap row 0, lo/hi of tap row 1 (sb_device.cuh)
    uint2 bw[MB_WARP_ROWS];
#pragma unroll
    for (int r = 0; r < MB_WARP_ROWS; ++r) {
        const unsigned x0 = t[r].x & 0xfff, x1 = (t[r].x >> 12) & 0xfff, y0 = t[r].y & 0xfff, y1 = (t[r].y >> 12) & 0xfff;
        bw[r] = __ldg(a.bilin_lut + ((t[r].x >> 24) | ((t[r].y >> 24) << 5)));
        load_tap_row(c.src + (size_t)y0 * c.sstep, x0, x1, tap[r][0], tap[r][1]);
        load_tap_row(c.src + (size_t)y1 * c.sstep, x0, x1, tap[r][2], tap[r][3]);
    }
#pragma unroll
    for (int r = 0; r < MB_WARP_ROWS; ++r) {
        const int py = py0 + 8 * r;
        if (py >= c.rh) break;
        int v0, v1, v2;
        bilinear_rgb(tap[r][0], tap[r][1], tap[r][2], tap[r][3], bw[r], v0, v1, v2);
        if (GAIN) {                                         // saturate_cast<uchar>(p * gain): GainCompensator / BlocksGainCompensator
            const float g = c.gmap ? __ldg(rowp<float>(c.gmap, c.gmstep, py) + px) : c.gain;
            v0 = min(max(__float2int_rn(__fmul_rn((float)v0, g)), 0), 255);
            v1 = min(max(__float2int_rn(__fmul_rn((float)v1, g)), 0), 255);
            v2 = min(max(__float2int_rn(__fmul_rn((float)v2, g)), 0), 255);
        }
        rowp<uint32_t>(c.g0, c.gstep, py)[px] = (unsigned)v0 | ((unsigned)v1 << 8) | ((unsigned)v2 << 16);
    }
}

int launch_mb_warp(const MbWarpArgs &a, bool apply_gain, int max_rw, int max_rh, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(max_rw, 32), div_up(max_rh, 8 * MB_WARP_ROWS), a.n);
    if (apply_gain) k_mb_warp<true><<<grid, block, 0, s>>>(a); else k_mb_warp<false><<<grid, block, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K2: pyrDown on RGBX
// channels 0 and 2 ride in the 16-bit lanes of one register, channel 1 in another
__device__ __forceinline__ unsigned lanes02(unsigned v) { return v & 0x00ff00ffu; }
__device__ __forceinline__ unsigned lane1(unsigned v) { return __byte_perm(v, 0u, 0x4441); }   // (v >> 8) & 0xff in one PRMT

// Each thread produces a 2x2 block of outputs from a 7x7 window of inputs.  With x2 the thread's
// column, the window starts at input column 4*x2 - 2: per input row one 8-byte, one 16-byte and one
// 4-byte load, all naturally aligned and contiguous across the warp's lanes (5.25 loads per output
// instead of 25).  Windows that cross the image edge take the BORDER_REFLECT_101 scalar path.
__device__ __forceinline__ unsigned pd_h02(const unsigned *v) { return lanes02(v[2]) * 6u + (lanes02(v[1]) + lanes02(v[3])) * 4u + lanes02(v[0]) + lanes02(v[4]); }
__device__ __forceinline__ unsigned pd_h1(const unsigned *v) { return lane1(v[2]) * 6u + (lane1(v[1]) + lane1(v[3])) * 4u + lane1(v[0]) + lane1(v[4]); }

__device__ __forceinline__ void mb_pyr_down_thread(const MbPyrCam &c, int x2, int y2)
{
    const int dw = (c.sw + 1) >> 1, dh = (c.sh + 1) >> 1;
    if (2 * x2 >= dw || 2 * y2 >= dh) return;
    if (!((2 * x2 + 1 >= c.ox[0] && 2 * x2 < c.ox[1]) || (2 * x2 + 1 >= c.ox[2] && 2 * x2 < c.ox[3]))) return;
    const int ix = 4 * x2 - 2, iy = 4 * y2 - 2;            // top-left of the 7x7 input window
    const bool interior = ix >= 0 && ix + 6 < c.sw && iy >= 0 && iy + 6 < c.sh;
    unsigned ha02[7], ha1[7], hb02[7], hb1[7];              // horizontal sums of output columns 2*x2 and 2*x2+1, per input row
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        unsigned v[7];
        if (interior) {
            const uint32_t *row = rowp<uint32_t>(c.src, c.sstep, iy + i) + ix;
            const uint2 p = __ldg(reinterpret_cast<const uint2 *>(row));
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(row + 2));
            v[0] = p.x; v[1] = p.y; v[2] = q.x; v[3] = q.y; v[4] = q.z; v[5] = q.w; v[6] = __ldg(row + 6);
        } else {
            const uint32_t *row = rowp<uint32_t>(c.src, c.sstep, reflect101(iy + i, c.sh));
#pragma unroll
            for (int j = 0; j < 7; ++j) v[j] = __ldg(row + reflect101(ix + j, c.sw));
        }
        ha02[i] = pd_h02(v); ha1[i] = pd_h1(v);
        hb02[i] = pd_h02(v + 2); hb1[i] = pd_h1(v + 2);
    }
    // vertical: <= 256*255 = 65280 (+128) per 16-bit lane, no carry between lanes; (s + 128) >> 8
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
        const int y = 2 * y2 + oy;
        if (y >= dh) break;
        const unsigned *A02 = ha02 + 2 * oy, *A1 = ha1 + 2 * oy, *B02 = hb02 + 2 * oy, *B1 = hb1 + 2 * oy;
        const unsigned sa02 = A02[2] * 6u + (A02[1] + A02[3]) * 4u + A02[0] + A02[4] + 0x00800080u;
        const unsigned sa1 = A1[2] * 6u + (A1[1] + A1[3]) * 4u + A1[0] + A1[4] + 128u;
        const unsigned sb02 = B02[2] * 6u + (B02[1] + B02[3]) * 4u + B02[0] + B02[4] + 0x00800080u;
        const unsigned sb1 = B1[2] * 6u + (B1[1] + B1[3]) * 4u + B1[0] + B1[4] + 128u;
        const unsigned oa = ((sa02 >> 8) & 0x00ff00ffu) | ((sa1 >> 8) << 8), ob = ((sb02 >> 8) & 0x00ff00ffu) | ((sb1 >> 8) << 8);
        uint32_t *d = rowp<uint32_t>(c.dst, c.dstep, y) + 2 * x2;
        if (2 * x2 + 1 < dw) *reinterpret_cast<uint2 *>(d) = make_uint2(oa, ob);
        else *d = oa;
    }
}

__global__ void __launch_bounds__(256) k_mb_pyr_down(const __grid_constant__ MbPyrArgs a)
{
    mb_pyr_down_thread(a.cam[blockIdx.z], blockIdx.x * 32 + threadIdx.x, blockIdx.y * 8 + threadIdx.y);
}

// The same over a compacted tile list: only the 64-column tiles that intersect a camera's wanted output runs are launched.
// (One camera of a ring straddles the panorama seam: its rect is panorama-wide but only both ends are wanted, and the
// other cameras are a quarter as wide - a rectangular grid sized for the widest camera had 73 % of its warps exit at once.)
__global__ void __launch_bounds__(256) k_mb_pyr_down_list(const __grid_constant__ MbPyrListArgs a)
{
    int k = 0;
    while (k + 1 < a.n_seg && (int)blockIdx.x >= a.seg[k + 1].first) ++k;
    const MbPyrSeg sg = a.seg[k];
    const int local = (int)blockIdx.x - sg.first, bx = local % sg.ntx, by = local / sg.ntx;
    mb_pyr_down_thread(a.p.cam[sg.cam], (sg.tx0 + bx) * 32 + (int)(threadIdx.x & 31), by * 8 + (int)(threadIdx.x >> 5));
}

int launch_mb_pyr_down_list(MbPyrListArgs &a, cudaStream_t s)
{
    a.n_seg = 0;
    int total = 0;
    for (int i = 0; i < a.p.n; ++i) {
        const MbPyrCam &c = a.p.cam[i];
        const int dw = (c.sw + 1) >> 1, dh = (c.sh + 1) >> 1;
        const int rows = div_up(div_up(dh, 2), 8);
        int t0[2], t1[2], nr = 0;
        for (int r = 0; r < 2; ++r) {
            const int lo = std::max(0, c.ox[2 * r]), hi = std::min(dw, c.ox[2 * r + 1]);
            if (hi <= lo) continue;
            t0[nr] = lo / 64; t1[nr] = div_up(hi, 64); ++nr;
        }
        if (nr == 2 && t0[1] <= t1[0]) { t1[0] = std::max(t1[0], t1[1]); nr = 1; }      // (runs are ascending)
        for (int r = 0; r < nr; ++r) {
            a.seg[a.n_seg++] = MbPyrSeg{i, t0[r], t1[r] - t0[r], total};
            total += (t1[r] - t0[r]) * rows;
        }
    }
    if (total == 0) return SB_OK;
    k_mb_pyr_down_list<<<total, 256, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

int launch_mb_pyr_down(const MbPyrArgs &a, int max_dw, int max_dh, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(div_up(max_dw, 2), 32), div_up(div_up(max_dh, 2), 8), a.n);
    k_mb_pyr_down<<<grid, block, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ K3: one band
// Each thread owns a 2x2 block of band pixels (X0, Y0 even).  For l < n every feed rect is aligned to
// even coordinates at level l (blenders.cpp:252-253 snaps to 2^num_bands), so the block shares ONE
// 3x3 neighbourhood of the coarser level: pyrUp (Appendix A3) gives
//   even = s(-1) + 6 s(0) + s(+1)     odd = 4 (s(0) + s(+1))
// with reflect-101 on the left/top and replicate on the right/bottom, 9 taps for 4 outputs.
__device__ __forceinline__ int mb_weighted(int lap, float w) { return __float2int_rz(__fmul_rn((float)lap, w)); }   // |lap*w| < 2^31: no x86 indefinite
__device__ __forceinline__ int mb_weighted(int lap, short w) { return (int)(short)((lap * (int)w) >> 8); }

struct Nbr { int l, c, r; };                                // coarse neighbour indices along one axis
__device__ __forceinline__ Nbr nbr_of(int c, int n)
{
    Nbr k;
    k.c = c; k.l = c == 0 ? (n > 1 ? 1 : 0) : c - 1; k.r = min(c + 1, n - 1);
    return k;
}

template <typename WT, bool NOT_TOP, bool FINAL, bool OUT8>
__device__ __forceinline__ void mb_band_thread(const MbBandArgs &a, int X0, int Y0)
{
    const int lw = FINAL ? a.out_w : a.g.lw, lh = FINAL ? a.out_h : a.g.lh;       // band 0 is cropped to dst_roi_final_
    if (X0 >= lw || Y0 >= lh || X0 + 1 < a.x_begin || X0 >= a.x_end) return;
    uint32_t cams = __ldg(a.tile_mask + (Y0 >> 3) * a.tiles_x + (X0 >> 5));       // the 2x2 block lies inside one 32x8 mask tile
    // "plain" tile: one camera, weight and weight sum exactly 1.0f everywhere.  Then short(lap * 1.0f) = lap and
    // normalizeUsingWeightMap's short(lap / (1.0f + 1e-5f)) = lap - sign(lap) (the quotient lies strictly between
    // |lap| - 1 and |lap| for 1 <= |lap| < 65536): no weight loads, no float arithmetic.
    const bool plain = sizeof(WT) == 4 && (cams >> 31) != 0u;
    cams &= 0x7fffffffu;

    int acc[2][2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[j][i][0] = acc[j][i][1] = acc[j][i][2] = 0;

    for (; cams; cams &= cams - 1) {
        const MbBandCam &c = a.g.cam[__ffs(cams) - 1];
        const int x = X0 - c.rx, y = Y0 - c.ry;
        if (x + 1 < 0 || y + 1 < 0 || x >= c.rw || y >= c.rh) continue;   // (for l < n rects are even-aligned and even-sized)
        WT w[2][2];
        unsigned g[2][2];
        if (NOT_TOP) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint2 gg = __ldg(reinterpret_cast<const uint2 *>(rowp<uint32_t>(c.fine, c.fstep, y + j) + x));
                g[j][0] = gg.x; g[j][1] = gg.y;
                if (plain) { w[j][0] = w[j][1] = (WT)1; continue; }
                const WT *wr = rowp<WT>(c.weight, c.wstep, y + j) + x;
                w[j][0] = __ldg(wr); w[j][1] = __ldg(wr + 1);
            }
        } else {     // top level: rect corners may be odd, sizes may be odd
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool in = (unsigned)(x + i) < (unsigned)c.rw && (unsigned)(y + j) < (unsigned)c.rh;
                    w[j][i] = in ? __ldg(rowp<WT>(c.weight, c.wstep, y + j) + x + i) : (WT)0;
                    g[j][i] = in ? __ldg(rowp<uint32_t>(c.fine, c.fstep, y + j) + x + i) : 0u;
                }
        }
        if (w[0][0] == (WT)0 && w[0][1] == (WT)0 && w[1][0] == (WT)0 && w[1][1] == (WT)0) continue;
        unsigned u02[2][2] = {{0u, 0u}, {0u, 0u}}, u1[2][2] = {{0u, 0u}, {0u, 0u}};
        if (NOT_TOP) {
            const int cw = c.rw >> 1, ch = c.rh >> 1;
            const Nbr kx = nbr_of(x >> 1, cw), ky = nbr_of(y >> 1, ch);
            const int rows[3] = {ky.l, ky.c, ky.r};
            unsigned e02[3], o02[3], e1[3], o1[3];          // horizontal even / odd sums per coarse row
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const uint32_t *cr = rowp<uint32_t>(c.coarse, c.cstep, rows[r]);
                const unsigned ta = __ldg(cr + kx.l), tb = __ldg(cr + kx.c), tc = __ldg(cr + kx.r);
                const unsigned a02 = lanes02(ta), b02 = lanes02(tb), c02 = lanes02(tc), a1 = lane1(ta), b1 = lane1(tb), c1 = lane1(tc);
                e02[r] = a02 + b02 * 6u + c02; o02[r] = (b02 + c02) * 4u;
                e1[r] = a1 + b1 * 6u + c1;     o1[r] = (b1 + c1) * 4u;
            }
            // vertical: even row = r0 + 6 r1 + r2, odd row = 4 (r1 + r2); <= 64 * 255 per 16-bit lane; then (v + 32) >> 6
            u02[0][0] = (((e02[0] + e02[1] * 6u + e02[2]) + 0x00200020u) >> 6) & 0x03ff03ffu;
            u02[0][1] = (((o02[0] + o02[1] * 6u + o02[2]) + 0x00200020u) >> 6) & 0x03ff03ffu;
            u02[1][0] = ((((e02[1] + e02[2]) * 4u) + 0x00200020u) >> 6) & 0x03ff03ffu;
            u02[1][1] = ((((o02[1] + o02[2]) * 4u) + 0x00200020u) >> 6) & 0x03ff03ffu;
            u1[0][0] = ((e1[0] + e1[1] * 6u + e1[2]) + 32u) >> 6;
            u1[0][1] = ((o1[0] + o1[1] * 6u + o1[2]) + 32u) >> 6;
            u1[1][0] = (((e1[1] + e1[2]) * 4u) + 32u) >> 6;
            u1[1][1] = (((o1[1] + o1[2]) * 4u) + 32u) >> 6;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (w[j][i] == (WT)0) continue;              // short(lap * 0) == 0 and (lap * 0) >> 8 == 0
                const int l0 = (int)__byte_perm(g[j][i], 0u, 0x4440) - (int)(u02[j][i] & 0xffffu);
                const int l1 = (int)__byte_perm(g[j][i], 0u, 0x4441) - (int)u1[j][i];
                const int l2 = (int)__byte_perm(g[j][i], 0u, 0x4442) - (int)(u02[j][i] >> 16);
                if (plain) { acc[j][i][0] = l0; acc[j][i][1] = l1; acc[j][i][2] = l2; continue; }
                acc[j][i][0] += mb_weighted(l0, w[j][i]); acc[j][i][1] += mb_weighted(l1, w[j][i]); acc[j][i][2] += mb_weighted(l2, w[j][i]);
            }
    }

    // restored coarser band: pyrUp per channel on CV_16S values (short4 pixels), shared 3x3 neighbourhood
    int up[2][2][3];
    if (NOT_TOP) {
        const int cw = a.g.lw >> 1, ch = a.g.lh >> 1;
        const Nbr kx = nbr_of(X0 >> 1, cw), ky = nbr_of(Y0 >> 1, ch);
        const int rows[3] = {ky.l, ky.c, ky.r};
        int e[3][3], o[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const short4 *cr = rowp<short4>(a.coarse_r, a.coarse_r_step, rows[r]);
            const short4 ta = __ldg(cr + kx.l), tb = __ldg(cr + kx.c), tc = __ldg(cr + kx.r);
            e[r][0] = ta.x + tb.x * 6 + tc.x; o[r][0] = (tb.x + tc.x) * 4;
            e[r][1] = ta.y + tb.y * 6 + tc.y; o[r][1] = (tb.y + tc.y) * 4;
            e[r][2] = ta.z + tb.z * 6 + tc.z; o[r][2] = (tb.z + tc.z) * 4;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // (a weighted mean of 16-bit values with weights summing to 64: pyrUp's saturate_cast<short> can never clamp)
            up[0][0][k] = (e[0][k] + e[1][k] * 6 + e[2][k] + 32) >> 6;
            up[0][1][k] = (o[0][k] + o[1][k] * 6 + o[2][k] + 32) >> 6;
            up[1][0][k] = ((e[1][k] + e[2][k]) * 4 + 32) >> 6;
            up[1][1][k] = ((o[1][k] + o[2][k]) * 4 + 32) >> 6;
        }
    }

    // normalizeUsingWeightMap (blenders.cpp:383-424) on the wrapped 16-bit sums, collapse add, output
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int Y = Y0 + j;
        if (Y >= lh) break;
        int v[2][3];
        bool masked[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int X = X0 + i;
            if (plain) {
                masked[i] = X < lw;
#pragma unroll
                for (int k = 0; k < 3; ++k) { const int p = acc[j][i][k]; v[i][k] = p - (p > 0) + (p < 0); }
            } else {
            const WT ws = X < lw ? __ldg(rowp<WT>(a.wsum, a.wsum_step, Y) + X) : (WT)0;
            if (sizeof(WT) == 4) {
                const float wf = (float)ws;
                masked[i] = wf > SB_WEIGHT_EPS;
                const SharedDiv div(__fadd_rn(wf, SB_WEIGHT_EPS));      // [1e-5, n + 1e-5]: fast-path range
#pragma unroll
                for (int k = 0; k < 3; ++k) v[i][k] = trunc_short(div((float)(short)acc[j][i][k]));
            } else {
                const int wi = (int)ws + 1;
                masked[i] = (int)ws > 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) v[i][k] = wi ? (short)((((int)(short)acc[j][i][k]) << 8) / wi) : 0;
            }
            }
            if (NOT_TOP) {   // restoreImageFromLaplacePyr: add(pyrUp(pyr[i+1]), pyr[i]) saturates
#pragma unroll
                for (int k = 0; k < 3; ++k)      // (8-bit output clamps to [0, 255] right after: the 16-bit clamp is subsumed)
                    v[i][k] = (FINAL && OUT8) ? up[j][i][k] + v[i][k] : sat_s16(up[j][i][k] + v[i][k]);
            }
            if (FINAL && !masked[i]) v[i][0] = v[i][1] = v[i][2] = 0;   // Blender::blend: dst_.setTo(0, dst_mask_ == 0)
        }
        const bool two = X0 + 1 < lw;
        if (FINAL) {
            if (OUT8) {      // result.convertTo(CV_8U); 2 px = 6 bytes at an even byte offset: three 16-bit stores
                uint8_t *o = rowp<uint8_t>(a.out, a.out_step, Y) + X0 * 3;
                const int b0 = sat_u8(v[0][0]), b1 = sat_u8(v[0][1]), b2 = sat_u8(v[0][2]);
                if (two && ((reinterpret_cast<uintptr_t>(o) & 1) == 0)) {
                    const int b3 = sat_u8(v[1][0]), b4 = sat_u8(v[1][1]), b5 = sat_u8(v[1][2]);
                    unsigned short *q = reinterpret_cast<unsigned short *>(o);
                    q[0] = (unsigned short)(b0 | (b1 << 8)); q[1] = (unsigned short)(b2 | (b3 << 8)); q[2] = (unsigned short)(b4 | (b5 << 8));
                } else {
                    o[0] = (uint8_t)b0; o[1] = (uint8_t)b1; o[2] = (uint8_t)b2;
                    if (two) { o[3] = (uint8_t)sat_u8(v[1][0]); o[4] = (uint8_t)sat_u8(v[1][1]); o[5] = (uint8_t)sat_u8(v[1][2]); }
                }
            } else {
                short *o = rowp<short>(a.out, a.out_step, Y) + X0 * 3;
                o[0] = (short)v[0][0]; o[1] = (short)v[0][1]; o[2] = (short)v[0][2];
                if (two) { o[3] = (short)v[1][0]; o[4] = (short)v[1][1]; o[5] = (short)v[1][2]; }
            }
            if (a.out_mask) {
                uint8_t *m = a.out_mask + (size_t)Y * a.mask_step + X0;
                m[0] = masked[0] ? 255 : 0;
                if (two) m[1] = masked[1] ? 255 : 0;
            }
        } else {
            short4 *o = rowp<short4>(a.out, a.out_step, Y) + X0;
            if (two) {
                uint4 pk;
                pk.x = (unsigned)(v[0][0] & 0xffff) | ((unsigned)v[0][1] << 16); pk.y = (unsigned)(v[0][2] & 0xffff);
                pk.z = (unsigned)(v[1][0] & 0xffff) | ((unsigned)v[1][1] << 16); pk.w = (unsigned)(v[1][2] & 0xffff);
                *reinterpret_cast<uint4 *>(o) = pk;
            } else
                o[0] = make_short4((short)v[0][0], (short)v[0][1], (short)v[0][2], 0);
        }
    }
}

template <typename WT, bool NOT_TOP, bool FINAL, bool OUT8>
__global__ void __launch_bounds__(256) k_mb_band(const __grid_constant__ MbBandArgs a)
{
    // block = 32 x 8 threads = 64 x 16 band pixels = 2 x 2 mask tiles of 32 x 8
    sbt::pdl_wait();
    sbt::pdl_release();
    mb_band_thread<WT, NOT_TOP, FINAL, OUT8>(a, (blockIdx.x * 32 + threadIdx.x) * 2, (blockIdx.y * 8 + threadIdx.y) * 2);
}

int launch_mb_band(const MbBandArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s)
{
    const int lw = final_band ? a.out_w : a.g.lw, lh = final_band ? a.out_h : a.g.lh;
    dim3 block(32, 8), grid(div_up(lw, 64), div_up(lh, 16));
    SB_ASSERT(div_up(a.g.lw, 32) == a.tiles_x);
#define SB_MB(WT, NT, FIN, O8) SB_CUDA(launch_pdl(k_mb_band<WT, NT, FIN, O8>, grid, block, 0, s, a))
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

// ------------------------------------------------------------------------------------ multi-level launches
// The coarse pyramid levels are tiny (<= 1/16 of the level-0 pixels in total) and each of their launches costs
// 7-13 us of latency on its own.  These two cooperative kernels run several levels in ONE launch: a persistent
// grid walks the 32x8-thread work items of a level, then crosses a grid-wide barrier (the next level reads what
// this one wrote through L2), level after level.
__global__ void __launch_bounds__(256) k_mb_pyr_tail(const __grid_constant__ MbPyrTailArgs a)
{
    cg::grid_group grid = cg::this_grid();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int lv = 0; lv < a.n_levels; ++lv) {
        const MbPyrArgs &L = a.level[lv];
        for (int item = blockIdx.x; item < a.items[lv]; item += gridDim.x) {
            int i = 0;
            while (i + 1 < L.n && item >= a.first_item[lv][i + 1]) ++i;
            const int local = item - a.first_item[lv][i], bx = local % a.tiles_x[lv][i], by = local / a.tiles_x[lv][i];
            mb_pyr_down_thread(L.cam[i], bx * 32 + tx, by * 8 + ty);
        }
        if (lv + 1 < a.n_levels) grid.sync();
    }
}

int launch_mb_pyr_tail(const MbPyrTailArgs &a, int sm_count, cudaStream_t s)
{
    int most = 0;
    for (int lv = 0; lv < a.n_levels; ++lv) most = std::max(most, a.items[lv]);
    if (most == 0) return SB_OK;
    static int per_sm = 0;                                  // co-resident CTAs per SM: the cooperative grid may not exceed it
    if (!per_sm) SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mb_pyr_tail, 256, 0));
    dim3 grid(std::min(most, std::max(1, per_sm) * sm_count)), block(256);
    void *params[] = {const_cast<MbPyrTailArgs *>(&a)};
    SB_CUDA(cudaLaunchCooperativeKernel((const void *)k_mb_pyr_tail, grid, block, params, 0, s));
    SB_LAUNCHED();
    return SB_OK;
}

template <typename WT> __global__ void __launch_bounds__(256) k_mb_band_head(const __grid_constant__ MbBandHeadArgs a)
{
    cg::grid_group grid = cg::this_grid();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int lv = 0; lv < a.n_levels; ++lv) {                // lv 0 = the top band, then finer
        const MbBandArgs &L = a.level[lv];
        for (int item = blockIdx.x; item < a.items[lv]; item += gridDim.x) {
            const int bx = item % a.tiles_x[lv], by = item / a.tiles_x[lv];
            const int X0 = (bx * 32 + tx) * 2, Y0 = (by * 8 + ty) * 2;
            if (lv == 0 && a.top_is_top) mb_band_thread<WT, false, false, false>(L, X0, Y0);
            else mb_band_thread<WT, true, false, false>(L, X0, Y0);
        }
        if (lv + 1 < a.n_levels) grid.sync();
    }
}

int launch_mb_band_head(const MbBandHeadArgs &a, bool float_weights, int sm_count, cudaStream_t s)
{
    int most = 0;
    for (int lv = 0; lv < a.n_levels; ++lv) most = std::max(most, a.items[lv]);
    if (most == 0) return SB_OK;
    static int per_sm[2] = {0, 0};
    int &occ = per_sm[float_weights ? 1 : 0];
    if (!occ) {
        if (float_weights) SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mb_band_head<float>, 256, 0));
        else SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mb_band_head<short>, 256, 0));
    }
    dim3 grid(std::min(most, std::max(1, occ) * sm_count)), block(256);
    void *params[] = {const_cast<MbBandHeadArgs *>(&a)};
    if (float_weights) SB_CUDA(cudaLaunchCooperativeKernel((const void *)k_mb_band_head<float>, grid, block, params, 0, s));
    else SB_CUDA(cudaLaunchCooperativeKernel((const void *)k_mb_band_head<short>, grid, block, params, 0, s));
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ all the coarse levels, one launch
// Gaussian levels l0 -> ... -> top, then bands top ... l_lo, as ONE ordinary (non-cooperative) launch.  The stages depend on
// each other only through global memory; instead of a grid-wide barrier (which needs the whole grid co-resident and so
// serialises frames that are in flight on other streams) the CTAs draw work items from one counter IN STAGE ORDER and an
// item of stage s waits until every item of stage s - 1 has been counted done.  Whoever waits holds a later item than all
// the items it waits for, and those have been drawn by CTAs that are running: progress never depends on co-residency.
// The counters only ever grow (64-bit, one set per in-flight slot; the host keeps their running totals): a launch owns the
// item numbers [base, base + total) of its slot's work counter, nothing is reset between frames.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename WT> __global__ void __launch_bounds__(256) k_mb_coarse(const __grid_constant__ MbCoarseArgs a)
{
    __shared__ unsigned long long s_item;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n_down = a.down.n_levels, n_stage = n_down + a.band.n_levels;
    const unsigned long long base = a.base;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.sync, 1ull) - base;
        __syncthreads();
        const unsigned long long item_g = s_item;
        if (item_g >= (unsigned long long)a.total) break;
        int st = 0;
        while (st + 1 < n_stage && item_g >= (unsigned long long)a.first[st + 1]) ++st;
        const int item = (int)item_g - a.first[st];
        if (st > 0) {                                        // every item of the stage before must be done
            if (threadIdx.x == 0) {
                const unsigned long long want = a.want[st - 1];
                while (ld_acquire_u64(a.sync + 1 + (st - 1)) < want) __nanosleep(256);      // (several frames' kernels may be polling: keep them off the L2)
            }
            __syncthreads();
        }
        if (st < n_down) {
            const MbPyrArgs &L = a.down.level[st];
            int i = 0;
            while (i + 1 < L.n && item >= a.down.first_item[st][i + 1]) ++i;
            const int local = item - a.down.first_item[st][i], bx = local % a.down.tiles_x[st][i], by = local / a.down.tiles_x[st][i];
            mb_pyr_down_thread(L.cam[i], bx * 32 + tx, by * 8 + ty);
        } else {
            const int lv = st - n_down;
            const MbBandArgs &L = a.band.level[lv];
            const int bx = item % a.band.tiles_x[lv], by = item / a.band.tiles_x[lv];
            const int X0 = (bx * 32 + tx) * 2, Y0 = (by * 8 + ty) * 2;
            if (lv == 0 && a.band.top_is_top) mb_band_thread<WT, false, false, false>(L, X0, Y0);
            else mb_band_thread<WT, true, false, false>(L, X0, Y0);
        }
        __threadfence();                                     // this CTA's results before its count
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(a.sync + 1 + st, 1ull);
    }
}

// totals: the slot's running counter values (work counter, then one per stage), advanced here for the next launch
int launch_mb_coarse(MbCoarseArgs &a, unsigned long long *totals, bool float_weights, int sm_count, bool alone, cudaStream_t s)
{
    const int n_down = a.down.n_levels, n_stage = n_down + a.band.n_levels;
    if (n_stage == 0) return SB_OK;
    SB_ASSERT(n_stage <= 2 * SB_MB_MAX_FUSED_LEVELS && a.sync);
    int total = 0, most = 0;
    for (int st = 0; st < n_stage; ++st) {
        const int items = st < n_down ? a.down.items[st] : a.band.items[st - n_down];
        SB_ASSERT(items > 0);                                // (an empty stage would never be counted done)
        a.first[st] = total;
        total += items;
        most = std::max(most, items);
    }
    a.first[n_stage] = total;
    a.total = total;
    // a frame alone on the GPU wants the stages over as fast as possible (4 CTAs per SM: C3 217 us per frame against 276);
    // with frames in flight on other streams one CTA per SM leaves the SMs to their big kernels (136 us per frame against 179)
    static const int forced = getenv("SB_MB_COARSE_CTAS") ? atoi(getenv("SB_MB_COARSE_CTAS")) : 0;      // tuning hook
    const int per_sm = forced > 0 ? forced : alone ? 4 : 1;
    dim3 grid(std::min(most, per_sm * sm_count)), block(256);
    a.base = totals[0];
    totals[0] += (unsigned long long)total + grid.x;        // (every CTA draws once past the end)
    for (int st = 0; st < n_stage; ++st) {
        totals[1 + st] += (unsigned long long)(a.first[st + 1] - a.first[st]);
        a.want[st] = totals[1 + st];
    }
    if (float_weights) k_mb_coarse<float><<<grid, block, 0, s>>>(a);
    else k_mb_coarse<short><<<grid, block, 0, s>>>(a);
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
