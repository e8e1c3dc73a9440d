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
//   k_feather_fused : warp (maps on the fly from separable tables) + gain + convertTo(16S) +
//                     FeatherBlender::feed over all cameras + blend + convertTo(8U)  — ONE launch per frame.
//   k_band_fused    : one Laplacian band of MultiBandBlender for all cameras: Laplacian formed on
//                     the fly from the cameras' Gaussian levels, weighted sum, normalise, collapse
//                     add of the coarser restored band, and at band 0 crop + mask + convertTo(8U).
//
// HBM-bound integer/byte work: each thread owns 4 consecutive panorama pixels so that the weight
// sums load as float4 and the panorama stores as 3 x 32-bit (8U) or 3 x 64-bit (16S) words.
#include "sb_device.cuh"
#include "sb_fused.h"
#include "sb_pyr.cuh"

namespace sb {
using namespace sbd;

#define SB_WEIGHT_EPS 1e-5f

__device__ __forceinline__ bool in_spans(const int span[4], int x0, int x1)   // [x0, x1) overlaps a span?
{
    return (x0 < span[1] && x1 > span[0]) || (x0 < span[3] && x1 > span[2]);
}

// mapBackward with the trig factored into per-column / per-row tables (bit-identical to the per-pixel
// sinf/cosf form: the table entries ARE those sinf/cosf values) followed by the fixed-point bilinear
// sample of cv::remap with BORDER_REFLECT.  Weights: every table entry carries the factor 32, so
// (sum w*p + 2^14) >> 15 == (v + 512) >> 10 with v = sum of the 5-bit products; the (0,0) entry
// {32767,0,0,1} of OpenCV's table equals an exact copy for 8-bit data, as does {32768,0,0,0}.
template <int KIND>
__device__ __forceinline__ void warp_sample(const FusedCam &c, int wx, int wy, int out[3])
{
    const float cs = __ldg(c.col_sin + wx), cc = __ldg(c.col_cos + wx), ra = __ldg(c.row_a + wy);
    float x_, y_, z_;
    if (KIND == SB_WARP_SPHERICAL) {
        const float rb = __ldg(c.row_b + wy);
        x_ = __fmul_rn(ra, cs); y_ = rb; z_ = __fmul_rn(ra, cc);
    } else if (KIND == SB_WARP_CYLINDRICAL) {
        x_ = cs; y_ = ra; z_ = cc;
    } else {
        x_ = cs; y_ = ra; z_ = c.one_minus_t2;
    }
    const float *m = c.k_rinv;
    float x = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x_), __fmul_rn(m[1], y_)), __fmul_rn(m[2], z_));
    float y = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x_), __fmul_rn(m[4], y_)), __fmul_rn(m[5], z_));
    float z = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x_), __fmul_rn(m[7], y_)), __fmul_rn(m[8], z_));
    if (KIND == SB_WARP_PLANE || z > 0) {
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
    } else
        x = y = -1.f;
    const int fsx = cvround(__fmul_rn(x, 32.f)), fsy = cvround(__fmul_rn(y, 32.f));
    const int fx = fsx & 31, fy = fsy & 31;
    const int sx = sat_s16(fsx >> 5), sy = sat_s16(fsy >> 5);
    int x0 = sx, x1 = sx + 1, y0 = sy, y1 = sy + 1;
    if (!((unsigned)sx < (unsigned)(c.sw - 1) && (unsigned)sy < (unsigned)(c.sh - 1))) {
        x0 = border_interp<BORDER_REFLECT>(x0, c.sw); x1 = border_interp<BORDER_REFLECT>(x1, c.sw);
        y0 = border_interp<BORDER_REFLECT>(y0, c.sh); y1 = border_interp<BORDER_REFLECT>(y1, c.sh);
    }
    const uint8_t *r0 = c.src + (size_t)y0 * c.sstep, *r1 = c.src + (size_t)y1 * c.sstep;
    const uint8_t *p00 = r0 + x0 * 3, *p01 = r0 + x1 * 3, *p10 = r1 + x0 * 3, *p11 = r1 + x1 * 3;
    const int ax = 32 - fx, ay = 32 - fy;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int h0 = (int)__ldg(p00 + k) * ax + (int)__ldg(p01 + k) * fx;
        int h1 = (int)__ldg(p10 + k) * ax + (int)__ldg(p11 + k) * fx;
        int v = (h0 * ay + h1 * fy + 512) >> 10;          // <= 255: FixedPtCast never saturates here
        if (c.apply_gain) v = sat_u8_f(__fmul_rn((float)v, c.gain));
        out[k] = v;
    }
}

// ------------------------------------------------------------------------------------ feather
template <int KIND, bool OUT8>
__global__ void __launch_bounds__(128)
k_feather_fused(const __grid_constant__ FeatherFusedArgs a)
{
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X0 >= a.pw || Y >= a.ph) return;
    int acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0;

    for (int i = 0; i < a.n; ++i) {
        const FusedCam &c = a.cam[i];
        const int y = Y - c.dy;
        if ((unsigned)y >= (unsigned)c.wh) continue;
        if (!in_spans(c.span, X0, X0 + 4)) continue;
        const float *wrow = reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.weight) + (size_t)y * c.wstep);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = X0 + j - c.dx;
            if ((unsigned)x >= (unsigned)c.ww) continue;
            const float w = __ldg(wrow + x);
            if (w == 0.f) continue;                       // short(p * 0) == 0: exact skip
            int p[3];
            warp_sample<KIND>(c, x, y, p);
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[j][k] += (int)trunc_short(__fmul_rn((float)p[k], w));
        }
    }
    // FeatherBlender::blend: normalizeUsingWeightMap, mask = weight > eps, zero unmasked, convertTo(8U)
    const float *ws = reinterpret_cast<const float *>(reinterpret_cast<const char *>(a.wsum) + (size_t)Y * a.wsum_step) + X0;
    int o[4][3], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float w = (X0 + j < a.pw) ? ws[j] : 0.f;
        m[j] = w > SB_WEIGHT_EPS ? 255 : 0;
        const float d = __fadd_rn(w, SB_WEIGHT_EPS);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int v = trunc_short(__fdiv_rn((float)(short)acc[j][k], d));
            v = m[j] ? v : 0;
            o[j][k] = OUT8 ? sat_u8(v) : v;
        }
    }
    const bool full = X0 + 4 <= a.pw;
    if (OUT8) {
        uint8_t *orow = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X0 * 3;
        if (full && ((reinterpret_cast<uintptr_t>(orow) & 3) == 0)) {
            uint32_t *q = reinterpret_cast<uint32_t *>(orow);
            q[0] = o[0][0] | (o[0][1] << 8) | (o[0][2] << 16) | (o[1][0] << 24);
            q[1] = o[1][1] | (o[1][2] << 8) | (o[2][0] << 16) | (o[2][1] << 24);
            q[2] = o[2][2] | (o[3][0] << 8) | (o[3][1] << 16) | (o[3][2] << 24);
        } else {
            for (int j = 0; j < 4 && X0 + j < a.pw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (uint8_t)o[j][k];
        }
    } else {
        short *orow = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X0 * 3;
        for (int j = 0; j < 4 && X0 + j < a.pw; ++j)
            for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (short)o[j][k];
    }
    if (a.out_mask) {
        uint8_t *mrow_ = a.out_mask + (size_t)Y * a.mask_step + X0;
        if (full && ((reinterpret_cast<uintptr_t>(mrow_) & 3) == 0))
            *reinterpret_cast<uint32_t *>(mrow_) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
        else
            for (int j = 0; j < 4 && X0 + j < a.pw; ++j) mrow_[j] = (uint8_t)m[j];
    }
}

int launch_feather_fused(const FeatherFusedArgs &a, int kind, bool out8, cudaStream_t s)
{
    dim3 block(32, 4), grid(div_up(div_up(a.pw, 4), 32), div_up(a.ph, 4));
#define SB_FF(K) do { if (out8) k_feather_fused<K, true><<<grid, block, 0, s>>>(a); else k_feather_fused<K, false><<<grid, block, 0, s>>>(a); } while (0)
    switch (kind) {
    case SB_WARP_PLANE: SB_FF(SB_WARP_PLANE); break;
    case SB_WARP_CYLINDRICAL: SB_FF(SB_WARP_CYLINDRICAL); break;
    case SB_WARP_SPHERICAL: SB_FF(SB_WARP_SPHERICAL); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", kind);
    }
#undef SB_FF
    SB_LAUNCHED();
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
