// kernels_blend.cu — blending kernels.
//
// Reference: MultiBandBlender::feed/blend (blenders.cpp:236-377), normalizeUsingWeightMap
// (:383-424), FeatherBlender (:115-155), createWeightMap (:427-432), Blender::feed/blend (:81-112).
// Arithmetic that must match bit for bit: static_cast<short>(short * float) truncates toward
// zero; `+=` on short wraps mod 2^16; add/subtract on CV_16S saturate; the int16-weight variant
// uses arithmetic >> 8 and C integer division; the normalize is a true IEEE divide by (w + 1e-5f).
#include "sb_device.cuh"
#include "sb_pyr.cuh"
#include "sb_kernels.h"

namespace sb {
using namespace sbd;

#define SB_WEIGHT_EPS 1e-5f

// ------------------------------------------------------------------------------------ weights
template <typename WT> __global__ void __launch_bounds__(256)
k_mask_to_weight(const uint8_t *mask, size_t mstep, int mw, int mh, WT *w0, size_t wstep, int ww, int wh, int top, int left)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ww || y >= wh) return;
    int mx = x - left, my = y - top;
    int m = ((unsigned)mx < (unsigned)mw && (unsigned)my < (unsigned)mh) ? mask[(size_t)my * mstep + mx] : -1;
    WT v;
    if (sizeof(WT) == 4) v = (WT)(m < 0 ? 0.f : __fmul_rn((float)m, (float)(1. / 255.)));   // convertTo(CV_32F, 1./255.)
    else v = (WT)(m < 0 ? 0 : m + (m != 0));                                                // add(w, 1, w, mask != 0)
    mrow<WT>(w0, wstep, y)[x] = v;
}

int launch_mask_to_weight(const DImage &mask, const DImage &w0, int top, int left, cudaStream_t s)
{
    SB_ASSERT(mask.type == SB_8UC1 && (w0.type == SB_32FC1 || w0.type == SB_16SC1));
    dim3 block(256), grid(div_up(w0.cols, 256), w0.rows);
    if (w0.type == SB_32FC1)
        k_mask_to_weight<float><<<grid, block, 0, s>>>(mask.ptr<uint8_t>(), mask.step, mask.cols, mask.rows, w0.ptr<float>(), w0.step, w0.cols, w0.rows, top, left);
    else
        k_mask_to_weight<short><<<grid, block, 0, s>>>(mask.ptr<uint8_t>(), mask.step, mask.cols, mask.rows, w0.ptr<short>(), w0.step, w0.cols, w0.rows, top, left);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ accumulate
__device__ __forceinline__ short weighted(int lap, float w) { return trunc_short(__fmul_rn((float)lap, w)); }
__device__ __forceinline__ short weighted(int lap, short w) { return (short)((lap * (int)w) >> 8); }
__device__ __forceinline__ float wadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ short wadd(short a, short b) { return (short)(a + b); }

// dst(ox+x, oy+y) += weighted(lap(x,y), w(x,y)); dst_w += w.  lap = fine - pyrUp(coarse) or fine.
template <typename ST, typename WT, bool HAS_COARSE>
__global__ void __launch_bounds__(256)
k_lap_accumulate(const ST *__restrict__ fine, size_t fstep, int fw, int fh, const ST *__restrict__ coarse, size_t cstep,
                 const WT *__restrict__ w, size_t wstep, short *dst, size_t dstep, WT *dst_w, size_t dwstep, int ox, int oy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= fw || y >= fh) return;
    const WT wv = crow<WT>(w, wstep, y)[x];
    const ST *f = crow<ST>(fine, fstep, y) + x * 3;
    short *d = mrow<short>(dst, dstep, oy + y) + (ox + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int lap = (int)f[c];
        if (HAS_COARSE) {
            int up = up_cast<ST>(pyr_up_sum<ST, 3>(coarse, cstep, fw >> 1, fh >> 1, y, x, c));
            lap = sizeof(ST) == 2 ? sat_s16(lap - up) : lap - up;
        }
        d[c] = (short)(d[c] + weighted(lap, wv));
    }
    if (dst_w) {
        WT *dw = mrow<WT>(dst_w, dwstep, oy + y) + (ox + x);
        *dw = wadd(*dw, wv);
    }
}

template <typename WT> __global__ void __launch_bounds__(256)
k_weight_accumulate(const WT *__restrict__ w, size_t wstep, int ww, int wh, WT *dst_w, size_t dwstep, int ox, int oy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ww || y >= wh) return;
    WT *dw = mrow<WT>(dst_w, dwstep, oy + y) + (ox + x);
    *dw = wadd(*dw, crow<WT>(w, wstep, y)[x]);
}

int launch_weight_accumulate(const DImage &w, const DImage &dst_w, int ox, int oy, cudaStream_t s)
{
    SB_ASSERT((w.type == SB_32FC1 || w.type == SB_16SC1) && dst_w.type == w.type);
    SB_ASSERT(ox >= 0 && oy >= 0 && ox + w.cols <= dst_w.cols && oy + w.rows <= dst_w.rows);
    dim3 block(32, 8), grid(div_up(w.cols, 32), div_up(w.rows, 8));
    if (w.type == SB_32FC1) k_weight_accumulate<float><<<grid, block, 0, s>>>(w.ptr<float>(), w.step, w.cols, w.rows, dst_w.ptr<float>(), dst_w.step, ox, oy);
    else k_weight_accumulate<short><<<grid, block, 0, s>>>(w.ptr<short>(), w.step, w.cols, w.rows, dst_w.ptr<short>(), dst_w.step, ox, oy);
    SB_LAUNCHED();
    return SB_OK;
}

// FeatherBlender::createWeightMaps (blenders.cpp:176-183): tmp = weights_sum(roi) is a VIEW, so setTo(1, tmp < eps) writes
// back into the shared sum; then divide(weight, tmp, weight) with cv::divide's "x / 0 = 0"
__global__ void __launch_bounds__(256)
k_weight_normalize(float *w, size_t wstep, int ww, int wh, float *sum, size_t sstep, int ox, int oy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ww || y >= wh) return;
    float *t = mrow<float>(sum, sstep, oy + y) + (ox + x);
    float d = *t;
    if (d < 1.1920928955078125e-07f) { d = 1.f; *t = 1.f; }
    float *p = mrow<float>(w, wstep, y) + x;
    *p = d != 0.f ? __fdiv_rn(*p, d) : 0.f;
}

int launch_weight_normalize(const DImage &w, const DImage &sum, int ox, int oy, cudaStream_t s)
{
    SB_ASSERT(w.type == SB_32FC1 && sum.type == SB_32FC1);
    SB_ASSERT(ox >= 0 && oy >= 0 && ox + w.cols <= sum.cols && oy + w.rows <= sum.rows);
    dim3 block(32, 8), grid(div_up(w.cols, 32), div_up(w.rows, 8));
    k_weight_normalize<<<grid, block, 0, s>>>(w.ptr<float>(), w.step, w.cols, w.rows, sum.ptr<float>(), sum.step, ox, oy);
    SB_LAUNCHED();
    return SB_OK;
}

int launch_lap_accumulate(const DImage &fine, const DImage &coarse, const DImage &w, const DImage &dst,
                          const DImage &dst_w, int ox, int oy, cudaStream_t s)
{
    SB_ASSERT(fine.type == SB_16SC3 || fine.type == SB_8UC3);
    SB_ASSERT(dst.type == SB_16SC3 && (w.type == SB_32FC1 || w.type == SB_16SC1) && (dst_w.empty() || dst_w.type == w.type));
    SB_ASSERT(w.rows == fine.rows && w.cols == fine.cols);
    SB_ASSERT(ox >= 0 && oy >= 0 && ox + fine.cols <= dst.cols && oy + fine.rows <= dst.rows);
    const bool has_coarse = !coarse.empty();
    if (has_coarse) SB_ASSERT(coarse.type == fine.type && coarse.cols * 2 == fine.cols && coarse.rows * 2 == fine.rows);
    dim3 block(32, 8), grid(div_up(fine.cols, 32), div_up(fine.rows, 8));
#define SB_LA(ST, WT, HC) k_lap_accumulate<ST, WT, HC><<<grid, block, 0, s>>>(fine.ptr<ST>(), fine.step, fine.cols, fine.rows, coarse.ptr<ST>(), coarse.step, w.ptr<WT>(), w.step, dst.ptr<short>(), dst.step, dst_w.ptr<WT>(), dst_w.step, ox, oy)
    const bool s16 = fine.type == SB_16SC3, wf = w.type == SB_32FC1;
    if (s16 && wf) { if (has_coarse) SB_LA(short, float, true); else SB_LA(short, float, false); }
    else if (s16) { if (has_coarse) SB_LA(short, short, true); else SB_LA(short, short, false); }
    else if (wf) { if (has_coarse) SB_LA(uint8_t, float, true); else SB_LA(uint8_t, float, false); }
    else { if (has_coarse) SB_LA(uint8_t, short, true); else SB_LA(uint8_t, short, false); }
#undef SB_LA
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ normalize / collapse
__device__ __forceinline__ short normalized(short p, float w) { return trunc_short(__fdiv_rn((float)p, __fadd_rn(w, SB_WEIGHT_EPS))); }
__device__ __forceinline__ short normalized(short p, short w)
{
    int wi = (int)w + 1;
    return wi == 0 ? (short)0 : (short)((((int)p) << 8) / wi);   // x86 would trap on wi == 0; never reached with valid weights
}

template <typename WT, bool COLLAPSE>
__global__ void __launch_bounds__(256)
k_normalize(const short *__restrict__ coarse, size_t cstep, const WT *__restrict__ w, size_t wstep, short *fine, size_t fstep, int fw, int fh)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= fw || y >= fh) return;
    const WT wv = crow<WT>(w, wstep, y)[x];
    short *f = mrow<short>(fine, fstep, y) + x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int v = normalized(f[c], wv);
        if (COLLAPSE) v = sat_s16(up_cast<short>(pyr_up_sum<short, 3>(coarse, cstep, fw >> 1, fh >> 1, y, x, c)) + v);
        f[c] = (short)v;
    }
}

int launch_normalize(const DImage &weight, const DImage &src, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_16SC3 && (weight.type == SB_32FC1 || weight.type == SB_16SC1));
    SB_ASSERT(weight.rows == src.rows && weight.cols == src.cols);
    dim3 block(32, 8), grid(div_up(src.cols, 32), div_up(src.rows, 8));
    if (weight.type == SB_32FC1)
        k_normalize<float, false><<<grid, block, 0, s>>>(nullptr, 0, weight.ptr<float>(), weight.step, src.ptr<short>(), src.step, src.cols, src.rows);
    else
        k_normalize<short, false><<<grid, block, 0, s>>>(nullptr, 0, weight.ptr<short>(), weight.step, src.ptr<short>(), src.step, src.cols, src.rows);
    SB_LAUNCHED();
    return SB_OK;
}

int launch_normalize_collapse(const DImage &coarse, const DImage &w, const DImage &fine, cudaStream_t s)
{
    SB_ASSERT(fine.type == SB_16SC3 && coarse.type == SB_16SC3 && (w.type == SB_32FC1 || w.type == SB_16SC1));
    SB_ASSERT(w.rows == fine.rows && w.cols == fine.cols && coarse.cols * 2 == fine.cols && coarse.rows * 2 == fine.rows);
    dim3 block(32, 8), grid(div_up(fine.cols, 32), div_up(fine.rows, 8));
    if (w.type == SB_32FC1)
        k_normalize<float, true><<<grid, block, 0, s>>>(coarse.ptr<short>(), coarse.step, w.ptr<float>(), w.step, fine.ptr<short>(), fine.step, fine.cols, fine.rows);
    else
        k_normalize<short, true><<<grid, block, 0, s>>>(coarse.ptr<short>(), coarse.step, w.ptr<short>(), w.step, fine.ptr<short>(), fine.step, fine.cols, fine.rows);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ finalize
// MSRC 0: mask = float weight > eps; 1: mask = int16 weight > 0; 2: mask copied from src_mask
template <int MSRC, bool OUT8>
__global__ void __launch_bounds__(256)
k_finalize(const short *__restrict__ src, size_t sstep, const void *__restrict__ wsrc, size_t wstep, void *out, size_t ostep,
           uint8_t *out_mask, size_t omstep, int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int m;
    if (MSRC == 0) m = crow<float>(wsrc, wstep, y)[x] > SB_WEIGHT_EPS ? 255 : 0;
    else if (MSRC == 1) m = crow<short>(wsrc, wstep, y)[x] > 0 ? 255 : 0;
    else m = crow<uint8_t>(wsrc, wstep, y)[x];
    const short *p = crow<short>(src, sstep, y) + x * 3;
    int v0 = m ? p[0] : 0, v1 = m ? p[1] : 0, v2 = m ? p[2] : 0;     // Blender::blend: dst_.setTo(0, dst_mask_ == 0)
    if (OUT8) {
        uint8_t *o = mrow<uint8_t>(out, ostep, y) + x * 3;
        o[0] = (uint8_t)sat_u8(v0); o[1] = (uint8_t)sat_u8(v1); o[2] = (uint8_t)sat_u8(v2);   // result.convertTo(CV_8U)
    } else {
        short *o = mrow<short>(out, ostep, y) + x * 3;
        o[0] = (short)v0; o[1] = (short)v1; o[2] = (short)v2;
    }
    if (out_mask) out_mask[(size_t)y * omstep + x] = (uint8_t)m;
}

int launch_finalize(const DImage &src, const DImage &weight, const DImage *src_mask, const DImage &out,
                    const DImage &out_mask, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_16SC3 && (out.type == SB_16SC3 || out.type == SB_8UC3));
    SB_ASSERT(out.rows <= src.rows && out.cols <= src.cols);
    SB_ASSERT(out_mask.empty() || (out_mask.type == SB_8UC1 && out_mask.rows == out.rows && out_mask.cols == out.cols));
    dim3 block(32, 8), grid(div_up(out.cols, 32), div_up(out.rows, 8));
    uint8_t *om = out_mask.empty() ? nullptr : out_mask.ptr<uint8_t>();
    const bool out8 = out.type == SB_8UC3;
#define SB_FIN(M, WS) do { if (out8) k_finalize<M, true><<<grid, block, 0, s>>>(src.ptr<short>(), src.step, (WS).data, (WS).step, out.data, out.step, om, out_mask.step, out.cols, out.rows); \
                           else k_finalize<M, false><<<grid, block, 0, s>>>(src.ptr<short>(), src.step, (WS).data, (WS).step, out.data, out.step, om, out_mask.step, out.cols, out.rows); } while (0)
    if (src_mask) { SB_ASSERT(src_mask->type == SB_8UC1); SB_FIN(2, *src_mask); }
    else if (weight.type == SB_32FC1) SB_FIN(0, weight);
    else if (weight.type == SB_16SC1) SB_FIN(1, weight);
    else return fail(SB_ERR_ASSERT, "finalize: bad weight type %d", weight.type);
#undef SB_FIN
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ feather weights
// distanceTransform(mask, CV_DIST_L1, 3): exact city-block distance to the nearest zero pixel of
// the image, saturating at 8192 (the 16.16 fixed-point INIT_DIST0 of distanceTransform_3x3).
// L1 is separable: vertical distance per column, then a min-plus scan along each row.
#define SB_DIST_INF 8192

__global__ void __launch_bounds__(128) k_dist_columns(const uint8_t *mask, size_t mstep, int w, int h, int *dcol, size_t dstep_i)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int d = SB_DIST_INF;
    for (int y = 0; y < h; ++y) {
        d = mask[(size_t)y * mstep + x] ? min(d + 1, SB_DIST_INF) : 0;
        dcol[(size_t)y * dstep_i + x] = d;
    }
    d = SB_DIST_INF;
    for (int y = h - 1; y >= 0; --y) {
        d = mask[(size_t)y * mstep + x] ? min(d + 1, SB_DIST_INF) : 0;
        int *p = dcol + (size_t)y * dstep_i + x;
        *p = min(*p, d);
    }
}

// one warp per row: d(x) = min( min_{x'<=x}(c(x') - x') + x , min_{x'>=x}(c(x') + x') - x )
__global__ void __launch_bounds__(128) k_dist_rows(const int *dcol, size_t dstep_i, int w, int h, float *dist, size_t fstep)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= h) return;
    const int *c = dcol + (size_t)warp * dstep_i;
    float *out = mrow<float>(dist, fstep, warp);
    const int BIG = 1 << 28;
    int carry = BIG;
    for (int x0 = 0; x0 < w; x0 += 32) {           // forward: prefix min of c(x') - x'
        int x = x0 + lane;
        int v = x < w ? c[x] - x : BIG;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = min(v, t); }
        v = min(v, carry);
        carry = __shfl_sync(0xffffffffu, v, 31);
        if (x < w) out[x] = __int_as_float(v + x);
    }
    carry = BIG;
    for (int x0 = ((w - 1) / 32) * 32; x0 >= 0; x0 -= 32) {   // backward: suffix min of c(x') + x'
        int x = x0 + lane;
        int v = x < w ? c[x] + x : BIG;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = min(v, t); }
        v = min(v, carry);
        carry = __shfl_sync(0xffffffffu, v, 0);
        if (x < w) {
            int d = min(min(__float_as_int(out[x]), v - x), SB_DIST_INF);
            out[x] = (float)d;
        }
    }
}

int launch_distance_l1(const DImage &mask, const DImage &dist, DevBuf &scratch, cudaStream_t s)
{
    SB_ASSERT(mask.type == SB_8UC1 && dist.type == SB_32FC1 && mask.rows == dist.rows && mask.cols == dist.cols);
    const size_t stride = (size_t)mask.cols;
    SB_TRY(scratch.ensure(stride * mask.rows * sizeof(int)));
    k_dist_columns<<<div_up(mask.cols, 128), 128, 0, s>>>(mask.ptr<uint8_t>(), mask.step, mask.cols, mask.rows, static_cast<int *>(scratch.p), stride);
    SB_LAUNCHED();
    k_dist_rows<<<div_up(mask.rows * 32, 128), 128, 0, s>>>(static_cast<int *>(scratch.p), stride, mask.cols, mask.rows, dist.ptr<float>(), dist.step);
    SB_LAUNCHED();
    return SB_OK;
}

// threshold(weight * sharpness, weight, 1.f, 1.f, THRESH_TRUNC)  (blenders.cpp:431)
__global__ void __launch_bounds__(256) k_weight_from_dist(float *d, size_t step, int w, int h, float sharpness)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    float *p = mrow<float>(d, step, y) + x;
    float v = __fmul_rn(*p, sharpness);
    *p = v > 1.f ? 1.f : v;
}

int launch_weight_from_dist(const DImage &dist, float sharpness, cudaStream_t s)
{
    SB_ASSERT(dist.type == SB_32FC1);
    dim3 block(256), grid(div_up(dist.cols, 256), dist.rows);
    k_weight_from_dist<<<grid, block, 0, s>>>(dist.ptr<float>(), dist.step, dist.cols, dist.rows, sharpness);
    SB_LAUNCHED();
    return SB_OK;
}

// FeatherBlender::feed (blenders.cpp:123-147)
template <typename ST> __global__ void __launch_bounds__(256)
k_feather_accumulate(const ST *__restrict__ img, size_t istep, int w, int h, const float *__restrict__ wm, size_t wstep,
                     short *dst, size_t dstep, float *dst_w, size_t dwstep, int dx, int dy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float wv = crow<float>(wm, wstep, y)[x];
    const ST *p = crow<ST>(img, istep, y) + x * 3;
    short *d = mrow<short>(dst, dstep, dy + y) + (dx + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (short)(d[c] + weighted((int)p[c], wv));
    if (dst_w) {
        float *dw = mrow<float>(dst_w, dwstep, dy + y) + dx + x;
        *dw = __fadd_rn(*dw, wv);
    }
}

int launch_feather_accumulate(const DImage &img, const DImage &w, const DImage &dst, const DImage &dst_w, int dx,
                              int dy, cudaStream_t s)
{
    SB_ASSERT((img.type == SB_16SC3 || img.type == SB_8UC3) && w.type == SB_32FC1 && dst.type == SB_16SC3);
    SB_ASSERT(dst_w.empty() || dst_w.type == SB_32FC1);
    SB_ASSERT(w.rows == img.rows && w.cols == img.cols);
    SB_ASSERT(dx >= 0 && dy >= 0 && dx + img.cols <= dst.cols && dy + img.rows <= dst.rows);
    dim3 block(32, 8), grid(div_up(img.cols, 32), div_up(img.rows, 8));
    if (img.type == SB_16SC3)
        k_feather_accumulate<short><<<grid, block, 0, s>>>(img.ptr<short>(), img.step, img.cols, img.rows, w.ptr<float>(), w.step, dst.ptr<short>(), dst.step, dst_w.ptr<float>(), dst_w.step, dx, dy);
    else
        k_feather_accumulate<uint8_t><<<grid, block, 0, s>>>(img.ptr<uint8_t>(), img.step, img.cols, img.rows, w.ptr<float>(), w.step, dst.ptr<short>(), dst.step, dst_w.ptr<float>(), dst_w.step, dx, dy);
    SB_LAUNCHED();
    return SB_OK;
}

// Blender::feed (blenders.cpp:81-102)
template <typename ST> __global__ void __launch_bounds__(256)
k_masked_copy(const ST *__restrict__ img, size_t istep, int w, int h, const uint8_t *__restrict__ mask, size_t mstep,
              short *dst, size_t dstep, uint8_t *dst_mask, size_t dmstep, int dx, int dy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t m = mask[(size_t)y * mstep + x];
    if (m) {
        const ST *p = crow<ST>(img, istep, y) + x * 3;
        short *d = mrow<short>(dst, dstep, dy + y) + (dx + x) * 3;
        d[0] = p[0]; d[1] = p[1]; d[2] = p[2];
    }
    dst_mask[(size_t)(dy + y) * dmstep + dx + x] |= m;
}

int launch_masked_copy(const DImage &img, const DImage &mask, const DImage &dst, const DImage &dst_mask, int dx,
                       int dy, cudaStream_t s)
{
    SB_ASSERT((img.type == SB_16SC3 || img.type == SB_8UC3) && mask.type == SB_8UC1 && dst.type == SB_16SC3 && dst_mask.type == SB_8UC1);
    SB_ASSERT(mask.rows == img.rows && mask.cols == img.cols);
    SB_ASSERT(dx >= 0 && dy >= 0 && dx + img.cols <= dst.cols && dy + img.rows <= dst.rows);
    dim3 block(32, 8), grid(div_up(img.cols, 32), div_up(img.rows, 8));
    if (img.type == SB_16SC3)
        k_masked_copy<short><<<grid, block, 0, s>>>(img.ptr<short>(), img.step, img.cols, img.rows, mask.ptr<uint8_t>(), mask.step, dst.ptr<short>(), dst.step, dst_mask.ptr<uint8_t>(), dst_mask.step, dx, dy);
    else
        k_masked_copy<uint8_t><<<grid, block, 0, s>>>(img.ptr<uint8_t>(), img.step, img.cols, img.rows, mask.ptr<uint8_t>(), mask.step, dst.ptr<short>(), dst.step, dst_mask.ptr<uint8_t>(), dst_mask.step, dx, dy);
    SB_LAUNCHED();
    return SB_OK;
}

// mask_warped = seam_mask & mask_warped (stitcher.cpp:294)
__global__ void __launch_bounds__(256) k_and_8u(const uint8_t *a, size_t astep, uint8_t *b, size_t bstep, int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    b[(size_t)y * bstep + x] &= a[(size_t)y * astep + x];
}

int launch_and_8u(const DImage &a, const DImage &b_inout, cudaStream_t s)
{
    SB_ASSERT(a.type == SB_8UC1 && b_inout.type == SB_8UC1 && a.rows == b_inout.rows && a.cols == b_inout.cols);
    dim3 block(256), grid(div_up(a.cols, 256), a.rows);
    k_and_8u<<<grid, block, 0, s>>>(a.ptr<uint8_t>(), a.step, b_inout.ptr<uint8_t>(), b_inout.step, a.cols, a.rows);
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
