// kernels_pointwise.cu — streaming kernels: exposure gain apply,
// convertTo (a7), copyMakeBorder (a10), cv::resize INTER_LINEAR float (a6), zero fill.
// Reference: exposure_compensate.cpp:150-153,225-246; stitcher.cpp:285,313; blenders.cpp:272-274.
#include "sb_device.cuh"
#include "sb_kernels.h"

namespace sb {
using namespace sbd;

template <typename T> __device__ __forceinline__ T *row_ptr(void *base, size_t step, int y)
{
    return reinterpret_cast<T *>(reinterpret_cast<char *>(base) + (size_t)y * step);
}
template <typename T> __device__ __forceinline__ const T *row_ptr(const void *base, size_t step, int y)
{
    return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) + (size_t)y * step);
}

// image *= gain  ->  saturate_cast<uchar>(p * (float)gain)   (cvtScale_<uchar,uchar,float>)
__global__ void __launch_bounds__(256) k_scale_8u(uint8_t *img, size_t step, int n_elems, int rows, float gain)
{
    int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int y = blockIdx.y;
    if (x >= n_elems || y >= rows) return;
    uint8_t *r = img + (size_t)y * step;
    if (x + 4 <= n_elems && ((reinterpret_cast<uintptr_t>(r) & 3) == 0)) {
        uchar4 v = *reinterpret_cast<uchar4 *>(r + x);
        v.x = (uint8_t)sat_u8_f(__fmul_rn((float)v.x, gain));
        v.y = (uint8_t)sat_u8_f(__fmul_rn((float)v.y, gain));
        v.z = (uint8_t)sat_u8_f(__fmul_rn((float)v.z, gain));
        v.w = (uint8_t)sat_u8_f(__fmul_rn((float)v.w, gain));
        *reinterpret_cast<uchar4 *>(r + x) = v;
    } else {
        for (int i = x; i < min(x + 4, n_elems); ++i) r[i] = (uint8_t)sat_u8_f(__fmul_rn((float)r[i], gain));
    }
}

int launch_scale_8u(const DImage &img, float gain, cudaStream_t s)
{
    SB_ASSERT(type_depth(img.type) == SB_8U);
    int n = img.cols * type_cn(img.type);
    dim3 block(256), grid(div_up(div_up(n, 4), 256), img.rows);
    k_scale_8u<<<grid, block, 0, s>>>(img.ptr<uint8_t>(), img.step, n, img.rows, gain);
    SB_LAUNCHED();
    return SB_OK;
}

__global__ void __launch_bounds__(256)
k_mul_map_8u(uint8_t *img, size_t step, int cols, int rows, const float *gmap, size_t gstep)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    float g = row_ptr<float>(gmap, gstep, y)[x];
    uint8_t *p = img + (size_t)y * step + x * 3;
    p[0] = (uint8_t)sat_u8_f(__fmul_rn((float)p[0], g));
    p[1] = (uint8_t)sat_u8_f(__fmul_rn((float)p[1], g));
    p[2] = (uint8_t)sat_u8_f(__fmul_rn((float)p[2], g));
}

int launch_mul_map_8u(const DImage &img, const DImage &gain_full, cudaStream_t s)
{
    SB_ASSERT(img.type == SB_8UC3 && gain_full.type == SB_32FC1);
    SB_ASSERT(img.rows == gain_full.rows && img.cols == gain_full.cols);
    dim3 block(256), grid(div_up(img.cols, 256), img.rows);
    k_mul_map_8u<<<grid, block, 0, s>>>(img.ptr<uint8_t>(), img.step, img.cols, img.rows, gain_full.ptr<float>(), gain_full.step);
    SB_LAUNCHED();
    return SB_OK;
}

// cv::resize(INTER_LINEAR) on CV_32FC1 (imgwarp.cpp HResizeLinear/VResizeLinear, float coefficients)
__global__ void __launch_bounds__(256)
k_resize_linear_32f(const float *src, size_t sstep, int sw, int sh, float *dst, size_t dstep, int dw, int dh,
                    double scale_x, double scale_y)
{
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw || dy >= dh) return;
    float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale_x), 0.5);
    int sx = __float2int_rd(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale_y), 0.5);
    int sy = __float2int_rd(fy);
    fy = __fsub_rn(fy, (float)sy);
    int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1), sx1 = min(sx + 1, sw - 1);
    const float *S0 = row_ptr<float>(src, sstep, sy0), *S1 = row_ptr<float>(src, sstep, sy1);
    float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
    float r0 = __fadd_rn(__fmul_rn(S0[sx], a0), __fmul_rn(S0[sx1], a1));
    float r1 = __fadd_rn(__fmul_rn(S1[sx], a0), __fmul_rn(S1[sx1], a1));
    row_ptr<float>(dst, dstep, dy)[dx] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
}

int launch_resize_linear_32f(const DImage &src, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.type == SB_32FC1 && dst.type == SB_32FC1);
    double inv_x = (double)dst.cols / src.cols, inv_y = (double)dst.rows / src.rows;
    dim3 block(256), grid(div_up(dst.cols, 256), dst.rows);
    k_resize_linear_32f<<<grid, block, 0, s>>>(src.ptr<float>(), src.step, src.cols, src.rows, dst.ptr<float>(), dst.step, dst.cols, dst.rows, 1. / inv_x, 1. / inv_y);
    SB_LAUNCHED();
    return SB_OK;
}

// convertTo between the depths the path uses: 8U -> 16S (widen), 16S -> 8U (saturate), same -> copy
template <typename S, typename D> __global__ void __launch_bounds__(256)
k_convert(const S *src, size_t sstep, D *dst, size_t dstep, int n_elems, int rows)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n_elems || y >= rows) return;
    int v = row_ptr<S>(src, sstep, y)[x];
    if (sizeof(D) == 1 && sizeof(S) == 2) v = sat_u8(v);
    row_ptr<D>(dst, dstep, y)[x] = (D)v;
}

int launch_convert(const DImage &src, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.rows == dst.rows && src.cols == dst.cols && type_cn(src.type) == type_cn(dst.type));
    int n = src.cols * type_cn(src.type);
    int sd = type_depth(src.type), dd = type_depth(dst.type);
    dim3 block(256), grid(div_up(n, 256), src.rows);
    if (sd == dd) {
        SB_CUDA(cudaMemcpy2DAsync(dst.data, dst.step, src.data, src.step, (size_t)src.cols * elem_size(src.type), src.rows, cudaMemcpyDeviceToDevice, s));
        return SB_OK;
    } else if (sd == SB_8U && dd == SB_16S)
        k_convert<uint8_t, short><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.step, dst.ptr<short>(), dst.step, n, src.rows);
    else if (sd == SB_16S && dd == SB_8U)
        k_convert<short, uint8_t><<<grid, block, 0, s>>>(src.ptr<short>(), src.step, dst.ptr<uint8_t>(), dst.step, n, src.rows);
    else if (sd == SB_8U && dd == SB_32F)
        k_convert<uint8_t, float><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.step, dst.ptr<float>(), dst.step, n, src.rows);
    else
        return fail(SB_ERR_NOT_IMPL, "convert %d -> %d not on the compositing path", src.type, dst.type);
    SB_LAUNCHED();
    return SB_OK;
}

// copyMakeBorder for pixel sizes 1,2,3,4,6 bytes; BORDER_CONSTANT pads with zeros
template <int ESZ, int BORDER> __global__ void __launch_bounds__(256)
k_copy_make_border(const uint8_t *src, size_t sstep, int sw, int sh, uint8_t *dst, size_t dstep, int dw, int dh, int top, int left)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    int sx = border_interp<BORDER>(x - left, sw), sy = border_interp<BORDER>(y - top, sh);
    uint8_t *d = dst + (size_t)y * dstep + (size_t)x * ESZ;
    if (sx < 0 || sy < 0) {
#pragma unroll
        for (int k = 0; k < ESZ; ++k) d[k] = 0;
    } else {
        const uint8_t *p = src + (size_t)sy * sstep + (size_t)sx * ESZ;
#pragma unroll
        for (int k = 0; k < ESZ; ++k) d[k] = p[k];
    }
}

template <int ESZ>
static int cmb_dispatch(int border, dim3 grid, dim3 block, cudaStream_t s, const DImage &src, const DImage &dst, int top, int left)
{
#define SB_CMB(B) k_copy_make_border<ESZ, B><<<grid, block, 0, s>>>(src.ptr<uint8_t>(), src.step, src.cols, src.rows, dst.ptr<uint8_t>(), dst.step, dst.cols, dst.rows, top, left)
    switch (border) {
    case SB_BORDER_CONSTANT: SB_CMB(BORDER_CONSTANT); break;
    case SB_BORDER_REPLICATE: SB_CMB(BORDER_REPLICATE); break;
    case SB_BORDER_REFLECT: SB_CMB(BORDER_REFLECT); break;
    case SB_BORDER_WRAP: SB_CMB(BORDER_WRAP); break;
    case SB_BORDER_REFLECT_101: SB_CMB(BORDER_REFLECT_101); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported border mode %d", border);
    }
#undef SB_CMB
    return SB_OK;
}

int launch_copy_make_border(const DImage &src, const DImage &dst, int top, int left, int border, cudaStream_t s)
{
    SB_ASSERT(src.type == dst.type);
    SB_ASSERT(top >= 0 && left >= 0 && dst.rows >= src.rows + top && dst.cols >= src.cols + left);
    dim3 block(256), grid(div_up(dst.cols, 256), dst.rows);
    switch (elem_size(src.type)) {
    case 1: SB_TRY(cmb_dispatch<1>(border, grid, block, s, src, dst, top, left)); break;
    case 2: SB_TRY(cmb_dispatch<2>(border, grid, block, s, src, dst, top, left)); break;
    case 3: SB_TRY(cmb_dispatch<3>(border, grid, block, s, src, dst, top, left)); break;
    case 4: SB_TRY(cmb_dispatch<4>(border, grid, block, s, src, dst, top, left)); break;
    case 6: SB_TRY(cmb_dispatch<6>(border, grid, block, s, src, dst, top, left)); break;
    default: return fail(SB_ERR_NOT_IMPL, "copyMakeBorder: unsupported type %d", src.type);
    }
    SB_LAUNCHED();
    return SB_OK;
}

int launch_set_zero(const DImage &img, cudaStream_t s)
{
    if (img.empty()) return SB_OK;
    SB_CUDA(cudaMemset2DAsync(img.data, img.step, 0, (size_t)img.cols * elem_size(img.type), img.rows, s));
    return SB_OK;
}

}  // namespace sb
