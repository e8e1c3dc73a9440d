// sb_tma.cuh — mbarrier + bulk async copy (TMA, cp.async.bulk) primitives for sm_100a, inline PTX.
// Used by the persistent frame kernels to stream the sequence-constant tables into shared memory
// ahead of the threads that consume them (the table stream is 50-80 % of a frame's HBM bytes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbt {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy before the first bulk copy targets them
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const uint32_t a = smem_u32(bar);
    unsigned done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(100);                                   // a waiting warp must not eat the issue slots of the warps it waits for
    }
}
// global -> shared bulk copy (16-byte aligned addresses, size a multiple of 16), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void bulk_g2s_addr(uint32_t dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 16-byte asynchronous copy global -> shared without register staging (LDGSTS)
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// the barrier receives one (pre-counted) arrival once every cp.async this thread issued so far has landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 2-D tensor copy global -> shared (UTMALDG), completion on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// mbarrier wait without polling instructions: try_wait suspends the warp in hardware until the phase completes or a
// time limit passes (only then does the loop go round)
__device__ __forceinline__ void mbar_wait_hw(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}


// Programmatic dependent launch, device side (host side: launch_pdl in sb_internal.h).  pdl_wait(): everything the
// predecessor grid wrote is visible after it (a no-op for a plain launch); it must come before the first access to memory
// the predecessor touches.  pdl_release(): the successor grid may start being scheduled from here on (it will itself wait).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace sbt
