// launch_gap.cu — how long does a back-to-back launch of a (nearly) empty persistent-shaped kernel take on B200,
// as a function of dynamic shared memory, block size and argument size?  (tuning experiment, not product code)
#include <cstdio>
#include <cuda_runtime.h>
struct Big { char b[6600]; };
struct Small { char b[64]; };
template <class A> __global__ void __launch_bounds__(1024, 1) k_empty(const __grid_constant__ A a, int *out)
{
    extern __shared__ int sm[];
    if (threadIdx.x == 0 && a.b[0] == 77) { sm[0] = 1; out[blockIdx.x] = sm[0]; }
}
template <class A> float run(int grid, int threads, int smem, int n)
{
    A a{}; int *out; cudaMalloc(&out, 4096);
    cudaFuncSetAttribute(k_empty<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; ++i) k_empty<A><<<grid, threads, smem>>>(a, out);
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i) k_empty<A><<<grid, threads, smem>>>(a, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(out);
    return ms * 1e3f / n;
}
int main()
{
    const int n = 2000;
    for (int smem : {0, 48 * 1024, 100 * 1024, 205 * 1024, 227 * 1024})
        for (int threads : {256, 800})
            for (int grid : {148, 296})
                printf("smem %3d KB threads %3d grid %3d : small args %.2f us/launch, 6.6 KB args %.2f us/launch\n", smem / 1024, threads, grid,
                       run<Small>(grid, threads, smem, n), run<Big>(grid, threads, smem, n));
    return 0;
}
