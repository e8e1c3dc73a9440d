// tma_f32.cu — does a FLOAT32 2-D tensor map with a 32x32 box and negative / out-of-range coordinates load? (debugging experiment)
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Args { CUtensorMap m[64]; int which; int x, y; float *out; };
__global__ void k(const __grid_constant__ Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sm)), "l"(reinterpret_cast<uint64_t>(a.m + a.which)), "r"(a.x), "r"(a.y), "r"(s32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)) : "memory");
    const float *f = reinterpret_cast<const float *>(sm);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) a.out[i] = f[i];
}
int main()
{
    const int W = 180, H = 100;
    size_t pitch = 768;
    float *src, *out; cudaMalloc(&src, pitch * H); cudaMalloc(&out, 4096);
    float h[192 * 100];
    for (int y = 0; y < H; ++y) for (int x = 0; x < 192; ++x) h[y * 192 + x] = y * 1000 + x;
    cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    Args a{};
    cuuint64_t gd[2] = {W, H}, gs[1] = {pitch};
    cuuint32_t bx[2] = {32, 32}, es[2] = {1, 1};
    for (int k = 0; k < 64; ++k) {
        CUresult r = enc(&a.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) printf("encode %d -> %d\n", k, (int)r);
    }
    a.out = out;
    int xs[] = {32, -31, 3, -31, -16, -1, -31, 1, 33, 147, 148, 149, 150, 179, 5, 5}, ys[] = {3, 3, -31, -31, -16, -1, -1, 1, 65, 5, 5, 5, 5, 5, 68, 69};
    for (int t = 0; t < 16; ++t) {
        a.which = t; a.x = xs[t]; a.y = ys[t];
        k<<<1, 128, 4096>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        float r[1024] = {0}; cudaMemcpy(r, out, sizeof r, cudaMemcpyDeviceToHost);
        printf("x %4d y %4d: %s  first row: %.0f ... %.0f; last row: %.0f\n", a.x, a.y, cudaGetErrorString(e), r[0], r[31], r[31 * 32]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
