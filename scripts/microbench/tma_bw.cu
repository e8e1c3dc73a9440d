// tma_bw.cu — per-SM throughput of small 2-D tensor TMA loads vs 1-D bulk loads on B200 (tuning experiment).
// Each CTA (one per SM) keeps DEPTH copies in flight from NW issuing warps; the source (31 MB) is L2 resident after the
// first pass, so the figure is the TMA / L2->SM path, not DRAM.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned ph)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(ph), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap *m, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
#ifndef NWARPS
#define NWARPS 4
#endif
constexpr int NW = NWARPS;
// mode 0: 2-D boxes (bw x rows) at pseudo-random positions; mode 1: 1-D bulk copies of `bytes`; mode 2: one 1-D bulk copy per box row
__global__ void __launch_bounds__(NW * 32, 1) k_tma(const __grid_constant__ CUtensorMap map, const unsigned char *src, int pitch, int W3, int H, int bw, int rows,
                                                    int mode, int iters, int STAGES, unsigned long long *cycles)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar[128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned box_bytes = (unsigned)bw * rows;
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long t0 = clock64();
    unsigned rng = blockIdx.x * 7919u + warp * 104729u + 12345u;
    // warp w owns stages w, w + NW, ...: issue, then wait for the stage issued STAGES / NW iterations ago
    for (int it = 0; it < iters; ++it) {
        const int st = (it * NW + warp) % STAGES;
        const unsigned ph = (unsigned)((it * NW + warp) / STAGES) & 1u;
        if (it * NW + warp >= STAGES) mbar_wait(&bar[st], ph ^ 1u);      // the previous use of this stage has landed
        rng = rng * 1664525u + 1013904223u;
        const int x = (int)((rng >> 8) % (unsigned)(W3 - bw)) & ~15, y = (int)((rng >> 4) % (unsigned)(H - rows));
        const uint32_t dst = s32(sm) + (uint32_t)st * ((box_bytes + 127u) & ~127u);
        if (lane == 0) {
            mbar_expect(&bar[st], box_bytes);
            if (mode == 0) tma2d(dst, &map, x, y, &bar[st]);
            else if (mode == 1) bulk1d(dst, src + ((size_t)y * pitch + x), box_bytes, &bar[st]);
        }
        if (mode == 2) {
            __syncwarp();
            for (int r = lane; r < rows; r += 32) bulk1d(dst + r * bw, src + ((size_t)(y + r) * pitch + x), bw, &bar[st]);
        }
    }
    // drain
    for (int k = 0; k < STAGES; ++k) {
        const int idx = iters * NW - STAGES + k;
        if (idx >= 0 && idx % NW == warp) mbar_wait(&bar[idx % STAGES], (unsigned)(idx / STAGES) & 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main()
{
    const int W3 = 5760, H = 5400, pitch = 5760;
    unsigned char *src; cudaMalloc(&src, (size_t)pitch * H); cudaMemset(src, 1, (size_t)pitch * H);
    unsigned long long *cyc; cudaMalloc(&cyc, 148 * 8);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    struct Cfg { int mode, bw, rows, stages; };
    std::vector<Cfg> cfgs;
    for (int st : {32}) {
        cfgs.push_back({0, 160, 16, st}); cfgs.push_back({0, 160, 48, st > 16 ? 16 : st}); cfgs.push_back({1, 4096, 1, st});
    }
    for (auto c : cfgs) {
        CUtensorMap m;
        cuuint64_t gd[2] = {(cuuint64_t)W3, (cuuint64_t)H}, gs[1] = {(cuuint64_t)pitch};
        cuuint32_t bx[2] = {(cuuint32_t)(c.mode == 0 ? c.bw : 16), (cuuint32_t)(c.mode == 0 ? c.rows : 1)}, es[2] = {1, 1};
        enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const unsigned box_bytes = (unsigned)c.bw * c.rows;
        const int iters = (int)(48u * 1024 * 1024 / box_bytes / 148 / NW) + 8;      // ~48 MB per launch over the GPU
        const int STAGES = c.stages;
        const int smem = STAGES * ((box_bytes + 127) & ~127);
        if (smem > 200 * 1024) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_tma<<<148, NW * 32, smem>>>(m, src, pitch, W3, H, c.bw, c.rows, c.mode, iters, STAGES, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)box_bytes * iters * NW * 148;
        unsigned long long hc[148]; cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (auto v : hc) mx = v > mx ? v : mx;
        const double us = mx / 1965.0;      // SM clock 1965 MHz
        printf("warps %2d mode %d  box %4d x %2d (%5u B) in flight %2d ops (%6.1f KB): kernel %6.1f us (events %6.1f), %7.1f GB/s total, %5.1f B/clk/SM, %6.0f clk per op (%s)\n", NW, c.mode, c.bw, c.rows,
               box_bytes, STAGES, STAGES * box_bytes / 1024.0, us, ms * 1e3, bytes / us / 1e3, bytes / 148 / (double)mx, (double)mx / (iters * NW), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
