// kernels_pyr.cu — Gaussian / Laplacian pyramid kernels.
//
// Reference: createLaplacePyr / restoreImageFromLaplacePyr (blenders.cpp:435-489, 520-530) on top of
// OpenCV 2.4.11 pyrDown/pyrUp:
//   pyrDown: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, integer (s + 128) >> 8, float s * (1/256)
//   pyrUp  : exact 2x, even = s(-1) + 6 s(0) + s(1), odd = 4 (s(0) + s(1)), left/top reflect-101,
//            right/bottom replicate, (s + 32) >> 6; 16S add/subtract saturate.
// Memory-bound stencils: shared-memory staging of the horizontal pass, no tensor cores.
#include "sb_device.cuh"
#include "sb_pyr.cuh"
#include "sb_kernels.h"

namespace sb {
using namespace sbd;

// ------------------------------------------------------------------------------------ pyrDown
template <typename T> struct Acc { typedef int type; };
template <> struct Acc<float> { typedef float type; };

// horizontal 5-tap, scalar order of pyrDown_: s2*6 + (s1+s3)*4 + s0 + s4
__device__ __forceinline__ int hsum(int s0, int s1, int s2, int s3, int s4) { return s2 * 6 + (s1 + s3) * 4 + s0 + s4; }
__device__ __forceinline__ float hsum(float s0, float s1, float s2, float s3, float s4)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s2, 6.f), __fmul_rn(__fadd_rn(s1, s3), 4.f)), s0), s4);
}
// FixedPtCast<int,T,8> / FltCast<float,8>
__device__ __forceinline__ void store_down(uint8_t *d, int v) { *d = (uint8_t)sat_u8((v + 128) >> 8); }
__device__ __forceinline__ void store_down(short *d, int v) { *d = (short)sat_s16((v + 128) >> 8); }
__device__ __forceinline__ void store_down(float *d, float v) { *d = __fmul_rn(v, 1.f / 256.f); }

constexpr int PD_TW = 64;   // output pixels per tile row
constexpr int PD_TH = 8;    // output rows per tile

template <typename T, int CN>
__global__ void __launch_bounds__(256)
k_pyr_down(const T *__restrict__ src, size_t sstep, int sw, int sh, T *__restrict__ dst, size_t dstep, int dw, int dh)
{
    typedef typename Acc<T>::type A;
    constexpr int TWE = PD_TW * CN;           // output elements per tile row
    constexpr int SROWS = 2 * PD_TH + 3;
    __shared__ A hs[SROWS][TWE];
    const int x0 = blockIdx.x * PD_TW, y0 = blockIdx.y * PD_TH;
    // phase 1: horizontal pass for the 2*TH+3 source rows this tile needs
    for (int i = threadIdx.x; i < SROWS * TWE; i += blockDim.x) {
        int r = i / TWE, e = i - r * TWE;
        int x = x0 + e / CN, c = e % CN;
        A v = 0;
        if (x < dw) {
            int sy = reflect101(2 * y0 + r - 2, sh);
            const T *row = crow<T>(src, sstep, sy);
            int xs0, xs1, xs2, xs3, xs4;
            int cx = 2 * x;
            if (cx >= 2 && cx + 2 < sw) { xs0 = cx - 2; xs1 = cx - 1; xs2 = cx; xs3 = cx + 1; xs4 = cx + 2; }
            else {
                xs0 = reflect101(cx - 2, sw); xs1 = reflect101(cx - 1, sw); xs2 = reflect101(cx, sw);
                xs3 = reflect101(cx + 1, sw); xs4 = reflect101(cx + 2, sw);
            }
            v = hsum((A)row[xs0 * CN + c], (A)row[xs1 * CN + c], (A)row[xs2 * CN + c], (A)row[xs3 * CN + c], (A)row[xs4 * CN + c]);
        }
        hs[r][e] = v;
    }
    __syncthreads();
    // phase 2: vertical pass
    for (int i = threadIdx.x; i < PD_TH * TWE; i += blockDim.x) {
        int r = i / TWE, e = i - r * TWE;
        int x = x0 + e / CN, y = y0 + r;
        if (x >= dw || y >= dh) continue;
        A v = hsum(hs[2 * r][e], hs[2 * r + 1][e], hs[2 * r + 2][e], hs[2 * r + 3][e], hs[2 * r + 4][e]);
        store_down(mrow<T>(dst, dstep, y) + (x0 * CN + e), v);
    }
}

int launch_pyr_down(const DImage &src, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.type == dst.type);
    SB_ASSERT(dst.cols == (src.cols + 1) / 2 && dst.rows == (src.rows + 1) / 2);
    SB_ASSERT(src.rows > 0 && src.cols > 0);
    dim3 block(256), grid(div_up(dst.cols, PD_TW), div_up(dst.rows, PD_TH));
#define SB_PD(T, CN) k_pyr_down<T, CN><<<grid, block, 0, s>>>(src.ptr<T>(), src.step, src.cols, src.rows, dst.ptr<T>(), dst.step, dst.cols, dst.rows)
    switch (src.type) {
    case SB_8UC1: SB_PD(uint8_t, 1); break;
    case SB_8UC3: SB_PD(uint8_t, 3); break;
    case SB_16SC1: SB_PD(short, 1); break;
    case SB_16SC3: SB_PD(short, 3); break;
    case SB_32FC1: SB_PD(float, 1); break;
    default: return fail(SB_ERR_NOT_IMPL, "pyrDown: unsupported type %d", src.type);
    }
#undef SB_PD
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ pyrUp
// MODE 0: dst = pyrUp(coarse)                      (T -> T)
// MODE 1: dst = saturate(fine - pyrUp(coarse))     (T fine -> 16S dst)   Laplacian level
// MODE 2: fine = saturate(pyrUp(coarse) + fine)    (16S in place)        collapse step
template <typename T, int CN, int MODE>
__global__ void __launch_bounds__(256)
k_pyr_up(const T *__restrict__ coarse, size_t cstep, int cw, int ch, const T *fine, size_t fstep, void *dst, size_t dstep)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int dw = cw * 2, dh = ch * 2;
    if (e >= dw * CN || y >= dh) return;
    const int x = e / CN, c = e - x * CN;
    int up = up_cast<T>(pyr_up_sum<T, CN>(coarse, cstep, cw, ch, y, x, c));
    if (MODE == 0) mrow<T>(dst, dstep, y)[e] = (T)up;
    else if (MODE == 1) mrow<short>(dst, dstep, y)[e] = (short)sat_s16((int)crow<T>(fine, fstep, y)[e] - up);
    else mrow<short>(dst, dstep, y)[e] = (short)sat_s16(up + (int)crow<T>(fine, fstep, y)[e]);
}

template <int MODE>
static int pyr_up_launch(const DImage &coarse, const DImage *fine, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(dst.cols == coarse.cols * 2 && dst.rows == coarse.rows * 2);
    const int cn = type_cn(coarse.type);
    dim3 block(64, 4), grid(div_up(dst.cols * cn, 64), div_up(dst.rows, 4));
    const void *f = fine ? fine->data : nullptr;
    size_t fs = fine ? fine->step : 0;
#define SB_PU(T, CN) k_pyr_up<T, CN, MODE><<<grid, block, 0, s>>>(coarse.ptr<T>(), coarse.step, coarse.cols, coarse.rows, static_cast<const T *>(f), fs, dst.data, dst.step)
    switch (coarse.type) {
    case SB_8UC1: SB_PU(uint8_t, 1); break;
    case SB_8UC3: SB_PU(uint8_t, 3); break;
    case SB_16SC1: SB_PU(short, 1); break;
    case SB_16SC3: SB_PU(short, 3); break;
    default: return fail(SB_ERR_NOT_IMPL, "pyrUp: unsupported type %d", coarse.type);
    }
#undef SB_PU
    SB_LAUNCHED();
    return SB_OK;
}

int launch_pyr_up(const DImage &src, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(src.type == dst.type);
    return pyr_up_launch<0>(src, nullptr, dst, s);
}

int launch_laplace_level(const DImage &fine, const DImage &coarse, const DImage &dst, cudaStream_t s)
{
    SB_ASSERT(fine.type == coarse.type && type_depth(dst.type) == SB_16S && type_cn(dst.type) == type_cn(fine.type));
    SB_ASSERT(fine.rows == dst.rows && fine.cols == dst.cols);
    return pyr_up_launch<1>(coarse, &fine, dst, s);
}

int launch_collapse_level(const DImage &coarse, const DImage &fine_inout, cudaStream_t s)
{
    SB_ASSERT(coarse.type == fine_inout.type && type_depth(coarse.type) == SB_16S);
    return pyr_up_launch<2>(coarse, &fine_inout, fine_inout, s);
}

}  // namespace sb
