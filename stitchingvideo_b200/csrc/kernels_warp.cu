// kernels_warp.cu — SURVEY.md §8a rows a3 (map construction), a4 (fixed-point bilinear remap),
// and the fused per-frame warp (a3+a4+a5+a7+a10) used by the compositor.
//
// Reference semantics: warpers_inl.hpp:62-99,206-300 (projectors, buildMaps, warp) and
// OpenCV 2.4.11 cv::remap: coordinates cvRound(x*32), 5 fractional bits,
// int16 weights scaled 2^15, (sum + 2^14) >> 15.  HBM-bound gather: no tensor cores.
#include "sb_device.cuh"
#include "sb_kernels.h"
#include "sb_warp.cuh"

namespace sb {
using namespace sbd;

template <int KIND>
__global__ void __launch_bounds__(256) k_build_maps(ProjParams p, int tl_x, int tl_y, float *xmap, size_t xstep,
                                                    float *ymap, size_t ystep, int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float mx, my;
    map_backward<KIND>(p, (float)(tl_x + x), (float)(tl_y + y), mx, my);
    reinterpret_cast<float *>(reinterpret_cast<char *>(xmap) + y * xstep)[x] = mx;
    reinterpret_cast<float *>(reinterpret_cast<char *>(ymap) + y * ystep)[x] = my;
}

int launch_build_maps(const ProjParams &p, int tl_x, int tl_y, const DImage &xmap, const DImage &ymap, cudaStream_t s)
{
    SB_ASSERT(xmap.type == SB_32FC1 && ymap.type == SB_32FC1 && xmap.rows == ymap.rows && xmap.cols == ymap.cols);
    dim3 block(32, 8), grid(div_up(xmap.cols, 32), div_up(xmap.rows, 8));
    switch (p.kind) {
    case SB_WARP_PLANE:
        k_build_maps<SB_WARP_PLANE><<<grid, block, 0, s>>>(p, tl_x, tl_y, xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, xmap.cols, xmap.rows);
        break;
    case SB_WARP_CYLINDRICAL:
        k_build_maps<SB_WARP_CYLINDRICAL><<<grid, block, 0, s>>>(p, tl_x, tl_y, xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, xmap.cols, xmap.rows);
        break;
    case SB_WARP_SPHERICAL:
        k_build_maps<SB_WARP_SPHERICAL><<<grid, block, 0, s>>>(p, tl_x, tl_y, xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, xmap.cols, xmap.rows);
        break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", p.kind);
    }
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ remap
// initInterTab2D(INTER_LINEAR, fixpt): weights sum to 32768; entry (0,0) is {32767,0,0,1}
__device__ __forceinline__ void bilinear_weights(int fx, int fy, int &w00, int &w01, int &w10, int &w11)
{
    w00 = (32 - fx) * (32 - fy) * 32;
    w01 = fx * (32 - fy) * 32;
    w10 = (32 - fx) * fy * 32;
    w11 = fx * fy * 32;
    if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
}

// sx, sy: integer tap coordinates (already clamped to int16), fx, fy: the 5 fractional bits
template <int CN, int BORDER>
__device__ __forceinline__ void sample_linear_q(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, int sx, int sy,
                                                int fx, int fy, const uint8_t *cval, int out[CN])
{
    int w00, w01, w10, w11;
    bilinear_weights(fx, fy, w00, w01, w10, w11);
    if (BORDER == BORDER_CONSTANT && (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0)) {
#pragma unroll
        for (int k = 0; k < CN; ++k) out[k] = cval[k];
        return;
    }
    int x0, x1, y0, y1;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        x0 = sx; x1 = sx + 1; y0 = sy; y1 = sy + 1;
    } else {
        x0 = border_interp<BORDER>(sx, sw); x1 = border_interp<BORDER>(sx + 1, sw);
        y0 = border_interp<BORDER>(sy, sh); y1 = border_interp<BORDER>(sy + 1, sh);
    }
    const uint8_t *r0 = src + (size_t)max(y0, 0) * sstep, *r1 = src + (size_t)max(y1, 0) * sstep;
#pragma unroll
    for (int k = 0; k < CN; ++k) {
        int v0 = (BORDER != BORDER_CONSTANT || (x0 >= 0 && y0 >= 0)) ? __ldg(r0 + x0 * CN + k) : cval[k];
        int v1 = (BORDER != BORDER_CONSTANT || (x1 >= 0 && y0 >= 0)) ? __ldg(r0 + x1 * CN + k) : cval[k];
        int v2 = (BORDER != BORDER_CONSTANT || (x0 >= 0 && y1 >= 0)) ? __ldg(r1 + x0 * CN + k) : cval[k];
        int v3 = (BORDER != BORDER_CONSTANT || (x1 >= 0 && y1 >= 0)) ? __ldg(r1 + x1 * CN + k) : cval[k];
        out[k] = sat_u8((v0 * w00 + v1 * w01 + v2 * w10 + v3 * w11 + (1 << 14)) >> 15);
    }
}

template <int CN, int BORDER>
__device__ __forceinline__ void sample_linear(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, float mx,
                                              float my, const uint8_t *cval, int out[CN])
{
    const int fsx = cvround(__fmul_rn(mx, 32.f)), fsy = cvround(__fmul_rn(my, 32.f));
    sample_linear_q<CN, BORDER>(src, sw, sh, sstep, sat_s16(fsx >> 5), sat_s16(fsy >> 5), fsx & 31, fsy & 31, cval, out);
}

template <int CN, int BORDER>
__device__ __forceinline__ void sample_nearest_q(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, int sx, int sy,
                                                 const uint8_t *cval, int out[CN])
{
    if (!((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh)) {
        if (BORDER == BORDER_CONSTANT) {
#pragma unroll
            for (int k = 0; k < CN; ++k) out[k] = cval[k];
            return;
        }
        sx = border_interp<BORDER>(sx, sw);
        sy = border_interp<BORDER>(sy, sh);
    }
    const uint8_t *r = src + (size_t)sy * sstep + sx * CN;
#pragma unroll
    for (int k = 0; k < CN; ++k) out[k] = __ldg(r + k);
}

template <int CN, int BORDER>
__device__ __forceinline__ void sample_nearest(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, float mx,
                                               float my, const uint8_t *cval, int out[CN])
{
    sample_nearest_q<CN, BORDER>(src, sw, sh, sstep, sat_s16(cvround(mx)), sat_s16(cvround(my)), cval, out);
}

struct CVal { uint8_t v[4]; };

template <int CN, int BORDER, int INTERP>
__global__ void __launch_bounds__(256)
k_remap(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, uint8_t *__restrict__ dst, int dw, int dh,
        size_t dstep, const float *__restrict__ xmap, size_t xstep, const float *__restrict__ ymap, size_t ystep, CVal cv)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float mx = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(xmap) + y * xstep) + x);
    float my = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(ymap) + y * ystep) + x);
    int out[CN];
    if (INTERP == SB_INTER_LINEAR) sample_linear<CN, BORDER>(src, sw, sh, sstep, mx, my, cv.v, out);
    else sample_nearest<CN, BORDER>(src, sw, sh, sstep, mx, my, cv.v, out);
    uint8_t *d = dst + y * dstep + x * CN;
#pragma unroll
    for (int k = 0; k < CN; ++k) d[k] = (uint8_t)out[k];
}

// cv::remap with the fixed-point map pair of cv::convertMaps / initUndistortRectifyMap(CV_16SC2) — the format the
// app's video front end uses (APP64:201-238, 741): map1 CV_16SC2 = integer coordinates, map2 CV_16UC1 = fy * 32 + fx
// (may be absent for INTER_NEAREST).  INTER_NEAREST with a fractional map adds OpenCV's NNDeltaTab_i {fx < 16, fy < 16}.
template <int CN, int BORDER, int INTERP>
__global__ void __launch_bounds__(256)
k_remap_fixed(const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, uint8_t *__restrict__ dst, int dw, int dh,
              size_t dstep, const short2 *__restrict__ map1, size_t m1step, const unsigned short *__restrict__ map2, size_t m2step, CVal cv)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const short2 xy = __ldg(reinterpret_cast<const short2 *>(reinterpret_cast<const char *>(map1) + y * m1step) + x);
    const int a = map2 ? (__ldg(reinterpret_cast<const unsigned short *>(reinterpret_cast<const char *>(map2) + y * m2step) + x) & 1023) : 0;
    int out[CN];
    if (INTERP == SB_INTER_LINEAR) sample_linear_q<CN, BORDER>(src, sw, sh, sstep, xy.x, xy.y, a & 31, a >> 5, cv.v, out);
    else {
        int sx = xy.x, sy = xy.y;
        if (map2) { sx = (short)(sx + ((a & 31) < 16)); sy = (short)(sy + ((a >> 5) < 16)); }
        sample_nearest_q<CN, BORDER>(src, sw, sh, sstep, sx, sy, cv.v, out);
    }
    uint8_t *d = dst + y * dstep + x * CN;
#pragma unroll
    for (int k = 0; k < CN; ++k) d[k] = (uint8_t)out[k];
}

// cv::convertMaps CV_32FC1 x / y -> CV_16SC2 (+ CV_16UC1): the map conversion cv::remap performs on every call, done once
__global__ void __launch_bounds__(256)
k_convert_maps(const float *xmap, size_t xstep, const float *ymap, size_t ystep, short2 *map1, size_t m1step, unsigned short *map2,
               size_t m2step, int w, int h, int nn)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float mx = reinterpret_cast<const float *>(reinterpret_cast<const char *>(xmap) + y * xstep)[x];
    const float my = reinterpret_cast<const float *>(reinterpret_cast<const char *>(ymap) + y * ystep)[x];
    short2 o;
    if (nn) {
        o.x = (short)sat_s16(cvround(mx)); o.y = (short)sat_s16(cvround(my));
    } else {
        const int ix = cvround(__fmul_rn(mx, 32.f)), iy = cvround(__fmul_rn(my, 32.f));
        o.x = (short)sat_s16(ix >> 5); o.y = (short)sat_s16(iy >> 5);
        reinterpret_cast<unsigned short *>(reinterpret_cast<char *>(map2) + y * m2step)[x] = (unsigned short)((iy & 31) * 32 + (ix & 31));
    }
    reinterpret_cast<short2 *>(reinterpret_cast<char *>(map1) + y * m1step)[x] = o;
}

int launch_convert_maps(const DImage &xmap, const DImage &ymap, const DImage &map1, const DImage &map2, bool nn, cudaStream_t s)
{
    SB_ASSERT(xmap.type == SB_32FC1 && ymap.type == SB_32FC1 && map1.type == SB_16SC2);
    SB_ASSERT(map1.rows == xmap.rows && map1.cols == xmap.cols && ymap.rows == xmap.rows && ymap.cols == xmap.cols);
    SB_ASSERT(nn || (map2.type == SB_16UC1 && map2.rows == xmap.rows && map2.cols == xmap.cols && map2.data));
    dim3 block(32, 8), grid(div_up(xmap.cols, 32), div_up(xmap.rows, 8));
    k_convert_maps<<<grid, block, 0, s>>>(xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, map1.ptr<short2>(), map1.step,
                                          nn ? nullptr : map2.ptr<unsigned short>(), map2.step, xmap.cols, xmap.rows, nn ? 1 : 0);
    SB_LAUNCHED();
    return SB_OK;
}

template <int CN, int BORDER>
static void remap_dispatch_interp(int interp, dim3 grid, dim3 block, cudaStream_t s, const DImage &src, const DImage &dst,
                                  const DImage &xmap, const DImage &ymap, CVal cv)
{
    if (xmap.type == SB_16SC2) {
        const unsigned short *m2 = ymap.empty() ? nullptr : ymap.ptr<unsigned short>();
        if (interp == SB_INTER_LINEAR)
            k_remap_fixed<CN, BORDER, SB_INTER_LINEAR><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.cols, src.rows, src.step, dst.ptr<uint8_t>(), dst.cols, dst.rows, dst.step, xmap.ptr<short2>(), xmap.step, m2, ymap.step, cv);
        else
            k_remap_fixed<CN, BORDER, SB_INTER_NEAREST><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.cols, src.rows, src.step, dst.ptr<uint8_t>(), dst.cols, dst.rows, dst.step, xmap.ptr<short2>(), xmap.step, m2, ymap.step, cv);
        return;
    }
    if (interp == SB_INTER_LINEAR)
        k_remap<CN, BORDER, SB_INTER_LINEAR><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.cols, src.rows, src.step, dst.ptr<uint8_t>(), dst.cols, dst.rows, dst.step, xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, cv);
    else
        k_remap<CN, BORDER, SB_INTER_NEAREST><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.cols, src.rows, src.step, dst.ptr<uint8_t>(), dst.cols, dst.rows, dst.step, xmap.ptr<float>(), xmap.step, ymap.ptr<float>(), ymap.step, cv);
}
template <int CN>
static int remap_dispatch_border(int border, int interp, dim3 grid, dim3 block, cudaStream_t s, const DImage &src,
                                 const DImage &dst, const DImage &xmap, const DImage &ymap, CVal cv)
{
    switch (border) {
    case SB_BORDER_CONSTANT: remap_dispatch_interp<CN, BORDER_CONSTANT>(interp, grid, block, s, src, dst, xmap, ymap, cv); break;
    case SB_BORDER_REPLICATE: remap_dispatch_interp<CN, BORDER_REPLICATE>(interp, grid, block, s, src, dst, xmap, ymap, cv); break;
    case SB_BORDER_REFLECT: remap_dispatch_interp<CN, BORDER_REFLECT>(interp, grid, block, s, src, dst, xmap, ymap, cv); break;
    case SB_BORDER_WRAP: remap_dispatch_interp<CN, BORDER_WRAP>(interp, grid, block, s, src, dst, xmap, ymap, cv); break;
    case SB_BORDER_REFLECT_101: remap_dispatch_interp<CN, BORDER_REFLECT_101>(interp, grid, block, s, src, dst, xmap, ymap, cv); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported border mode %d", border);
    }
    return SB_OK;
}

int launch_remap(const DImage &src, const DImage &dst, const DImage &xmap, const DImage &ymap, int interp, int border,
                 const uint8_t bv[4], cudaStream_t s)
{
    SB_ASSERT(src.type == SB_8UC1 || src.type == SB_8UC3);
    SB_ASSERT(dst.type == src.type);
    SB_ASSERT(interp == SB_INTER_LINEAR || interp == SB_INTER_NEAREST);
    if (xmap.type == SB_16SC2) {      // fixed-point pair; the fractional map is optional for INTER_NEAREST only (as in OpenCV)
        SB_ASSERT(ymap.empty() ? interp == SB_INTER_NEAREST : (ymap.type == SB_16UC1 && ymap.rows == xmap.rows && ymap.cols == xmap.cols));
    } else {
        SB_ASSERT(xmap.type == SB_32FC1 && ymap.type == SB_32FC1 && ymap.rows == xmap.rows && ymap.cols == xmap.cols);
    }
    SB_ASSERT(dst.rows == xmap.rows && dst.cols == xmap.cols);
    SB_ASSERT(src.rows > 0 && src.cols > 0);
    CVal cv;
    for (int i = 0; i < 4; ++i) cv.v[i] = bv ? bv[i] : 0;
    dim3 block(32, 8), grid(div_up(dst.cols, 32), div_up(dst.rows, 8));
    if (src.type == SB_8UC1) SB_TRY(remap_dispatch_border<1>(border, interp, grid, block, s, src, dst, xmap, ymap, cv));
    else SB_TRY(remap_dispatch_border<3>(border, interp, grid, block, s, src, dst, xmap, ymap, cv));
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ fused warp
// The trig in mapBackward is separable: sinf/cosf(u/scale) depend only on the column,
// sinf/cosf(pi - v/scale) only on the row.  Two tiny tables replace the 8 B/px float maps.
template <int KIND>
__global__ void k_warp_tables(ProjParams p, int tl_x, int tl_y, int w, int h, float *col_sin, float *col_cos,
                              float *row_a, float *row_b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w) {
        float u = (float)(tl_x + i);
        if (KIND == SB_WARP_PLANE) {
            col_sin[i] = __fsub_rn(__fdiv_rn(u, p.scale), p.t[0]);
            col_cos[i] = 0.f;
        } else {
            u = __fdiv_rn(u, p.scale);
            col_sin[i] = sinf_exact(u);
            col_cos[i] = cosf_exact(u);
        }
    } else if (i < w + h) {
        int j = i - w;
        float v = (float)(tl_y + j);
        if (KIND == SB_WARP_PLANE) {
            row_a[j] = __fsub_rn(__fdiv_rn(v, p.scale), p.t[1]);
            row_b[j] = 0.f;
        } else if (KIND == SB_WARP_CYLINDRICAL) {
            row_a[j] = __fdiv_rn(v, p.scale);
            row_b[j] = 0.f;
        } else {
            v = __fdiv_rn(v, p.scale);
            row_a[j] = sinf_exact(__fsub_rn(SB_PI_F, v));
            row_b[j] = cosf_exact(__fsub_rn(SB_PI_F, v));
        }
    }
}

int launch_build_warp_tables(const ProjParams &p, int tl_x, int tl_y, int w, int h, float *col_sin, float *col_cos,
                             float *row_a, float *row_b, cudaStream_t s)
{
    int n = w + h, block = 128, grid = div_up(n, block);
    switch (p.kind) {
    case SB_WARP_PLANE: k_warp_tables<SB_WARP_PLANE><<<grid, block, 0, s>>>(p, tl_x, tl_y, w, h, col_sin, col_cos, row_a, row_b); break;
    case SB_WARP_CYLINDRICAL: k_warp_tables<SB_WARP_CYLINDRICAL><<<grid, block, 0, s>>>(p, tl_x, tl_y, w, h, col_sin, col_cos, row_a, row_b); break;
    case SB_WARP_SPHERICAL: k_warp_tables<SB_WARP_SPHERICAL><<<grid, block, 0, s>>>(p, tl_x, tl_y, w, h, col_sin, col_cos, row_a, row_b); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", p.kind);
    }
    SB_LAUNCHED();
    return SB_OK;
}

// One thread per destination pixel of the padded rect.  OUT16: write CV_16SC3, else CV_8UC3.
template <int KIND, bool OUT16, bool GAIN>
__global__ void __launch_bounds__(256)
k_warp_fused(ProjParams p, WarpTables t, const uint8_t *__restrict__ src, int sw, int sh, size_t sstep, int ww, int wh,
             int left, int top, float gain, void *__restrict__ dst, int dw, int dh, size_t dstep)
{
    int px = blockIdx.x * blockDim.x + threadIdx.x;
    int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= dw || py >= dh) return;
    // copyMakeBorder(BORDER_REFLECT) of the warped image (blenders.cpp:272-274), folded into the load
    int wx = border_interp<BORDER_REFLECT>(px - left, ww);
    int wy = border_interp<BORDER_REFLECT>(py - top, wh);
    float mx, my;
    map_backward_tab<KIND>(p.k_rinv, p.t[2], __ldg(t.col_sin + wx), __ldg(t.col_cos + wx), __ldg(t.row_a + wy),
                           __ldg(t.row_b + wy), mx, my);
    int out[3];
    sample_linear<3, BORDER_REFLECT>(src, sw, sh, sstep, mx, my, nullptr, out);
    if (GAIN) {
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = sat_u8_f(__fmul_rn((float)out[k], gain));
    }
    if (OUT16) {
        short *d = reinterpret_cast<short *>(reinterpret_cast<char *>(dst) + py * dstep) + px * 3;
        d[0] = (short)out[0]; d[1] = (short)out[1]; d[2] = (short)out[2];
    } else {
        uint8_t *d = reinterpret_cast<uint8_t *>(dst) + py * dstep + px * 3;
        d[0] = (uint8_t)out[0]; d[1] = (uint8_t)out[1]; d[2] = (uint8_t)out[2];
    }
}

template <int KIND>
static void warp_fused_dispatch(bool out16, bool gain_on, dim3 grid, dim3 block, cudaStream_t s, const ProjParams &p,
                                const WarpTables &t, const DImage &src, int ww, int wh, int left, int top, float gain,
                                const DImage &dst)
{
#define SB_WF(O, G) k_warp_fused<KIND, O, G><<<grid, block, 0, s>>>(p, t, src.ptr<uint8_t>(), src.cols, src.rows, src.step, ww, wh, left, top, gain, dst.data, dst.cols, dst.rows, dst.step)
    if (out16) { if (gain_on) SB_WF(true, true); else SB_WF(true, false); }
    else       { if (gain_on) SB_WF(false, true); else SB_WF(false, false); }
#undef SB_WF
}

int launch_warp_fused(const ProjParams &p, const WarpTables &t, const DImage &src, int warped_w, int warped_h, int left,
                      int top, float gain, bool apply_gain, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_8UC3);
    SB_ASSERT(dst.type == SB_16SC3 || dst.type == SB_8UC3);
    dim3 block(32, 8), grid(div_up(dst.cols, 32), div_up(dst.rows, 8));
    bool out16 = dst.type == SB_16SC3;
    switch (p.kind) {
    case SB_WARP_PLANE: warp_fused_dispatch<SB_WARP_PLANE>(out16, apply_gain, grid, block, s, p, t, src, warped_w, warped_h, left, top, gain, dst); break;
    case SB_WARP_CYLINDRICAL: warp_fused_dispatch<SB_WARP_CYLINDRICAL>(out16, apply_gain, grid, block, s, p, t, src, warped_w, warped_h, left, top, gain, dst); break;
    case SB_WARP_SPHERICAL: warp_fused_dispatch<SB_WARP_SPHERICAL>(out16, apply_gain, grid, block, s, p, t, src, warped_w, warped_h, left, top, gain, dst); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", p.kind);
    }
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
