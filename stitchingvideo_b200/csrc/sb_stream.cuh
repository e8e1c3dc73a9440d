// sb_stream.cuh — the shared machinery of the persistent, warp-specialised streaming frame kernels
// (k_feather_stream in kernels_feather_tma.cu, k_mb_warp_stream in kernels_mb_stream.cu): shared-memory layout,
// the producer warps (table blocks by TMA bulk copy, source boxes by cp.async, per-calibration ring plan), and the
// consumer-side tap fetch from a staged source box.
#pragma once
#include "sb_device.cuh"
#include "sb_fused.h"
#include "sb_tma.cuh"

namespace sb {
using namespace sbd;
using namespace sbt;

// Shared memory: a byte ring (128-byte units) that holds, per (tile, camera) in flight, the table block followed by
// the source box; a ring of SB_FTT_STAGES tile entries (descriptor + full/empty barriers); the bilinear weight table.
constexpr int FTS_RING_UNITS = SB_FTS_RING_BYTES / 128;
struct FtsSmem {
    unsigned char ring[SB_FTS_RING_BYTES];
    uint2 lut[1024];                                        // bilin_lut (sb_device.cuh): 8 KB, read once per camera pixel
    uint4 desc[SB_FTT_STAGES][1 + SB_FTT_MAXC];             // [0] = {n_cams | short << 2, tile origin, output byte offset, mask byte offset}; [1 + k].x = shared address of slot k
    uint64_t full[SB_FTT_STAGES], empty[SB_FTT_STAGES];
};

__device__ __forceinline__ uint2 lds_u2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// The two tap rows of one boxed table entry: lo = bytes 0..3, hi = bytes 4..7 of the 6 tap bytes of each row.
// tex = byte shift (bits 0-4) | word offset in the box (bits 5-20) | (y1 == y0) << 27
__device__ __forceinline__ void fts_taps(uint32_t box, unsigned pitch, unsigned tex, unsigned &lo0, unsigned &hi0, unsigned &lo1, unsigned &hi1)
{
    const uint32_t r0 = box + ((tex >> 3) & 0x3fffcu);
    const uint32_t r1 = (tex & (1u << 27)) ? r0 : r0 + pitch;
    unsigned a0, a1, a2, b0, b1, b2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(r0));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(a1) : "r"(r0));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(a2) : "r"(r0));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(r1));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(b1) : "r"(r1));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(b2) : "r"(r1));
    lo0 = __funnelshift_r(a0, a1, tex); hi0 = __funnelshift_r(a1, a2, tex);      // (shift amount = tex mod 32)
    lo1 = __funnelshift_r(b0, b1, tex); hi1 = __funnelshift_r(b1, b2, tex);
}
// the taps of a row-major entry (box too large for shared memory), gathered from global memory
template <class Cam>
__device__ __forceinline__ void fts_taps_direct(const Cam &c, unsigned tex, unsigned &lo0, unsigned &hi0, unsigned &lo1, unsigned &hi1)
{
    const unsigned x0 = tex & 0x1fffu, y0 = (tex >> 13) & 0x1fffu;
    const uint8_t *g0 = c.src + (size_t)(y0 * c.sstep);
    const uint8_t *g1 = (tex & (1u << 27)) ? g0 : g0 + c.sstep;
    const unsigned x1 = x0 + 1u - ((tex >> 26) & 1u);
    load_tap_row(g0, x0, x1, lo0, hi0);
    load_tap_row(g1, x0, x1, lo1, hi1);
}
// bilinear_rgb (sb_device.cuh); MINUS1: max(v - 1, 0) instead of v, folded into the rounding bias:
// v = (s + 512) >> 10 with s >= 0, so max(v - 1, 0) = max(s - 512, 0) >> 10.
template <bool MINUS1>
__device__ __forceinline__ void fts_bilinear(unsigned lo0, unsigned hi0, unsigned lo1, unsigned hi1, uint2 w, int &v0, int &v1, int &v2)
{
    if (!MINUS1) { bilinear_rgb(lo0, hi0, lo1, hi1, w, v0, v1, v2); return; }
    const unsigned m0 = __byte_perm(lo0, hi0, 0x5241), m1 = __byte_perm(lo1, hi1, 0x5241);
    const unsigned p0 = __byte_perm(lo0, lo1, 0x7430), p1 = __byte_perm(m0, m1, 0x5410), p2 = __byte_perm(m0, m1, 0x7632);
    v0 = max((int)bilin_sum(p0, w, 0xfffffe00u), 0) >> 10;
    v1 = max((int)bilin_sum(p1, w, 0xfffffe00u), 0) >> 10;
    v2 = max((int)bilin_sum(p2, w, 0xfffffe00u), 0) >> 10;
}
// The producer side of a streaming frame kernel; called by every producer warp (pw = 0 .. SB_FTS_PRODUCER_WARPS-1) of a
// persistent CTA of a grid of G.  Args: .desc, .n_tiles, .cam[i].{src, sstep, tiles}.  (ostep, obpp, mstep): the stage
// descriptor's [0].z / .w receive the byte offsets Y0 * ostep + X0 * obpp and Y0 * mstep + X0 of the tile origin.
template <class Args>
__device__ __forceinline__ void stream_producer(const Args &a, FtsSmem &sm, int pw, int lane, int G, unsigned ostep, unsigned obpp, unsigned mstep)
{
    // Producer warp w fetches tiles w, w + P, ... of this CTA's sequence (seq = 0, 1, 2 ...; tile = blockIdx + seq * G),
    // so the per-tile issue latency overlaps across warps.  Where a tile lands in the ring and how many of the CTA's
    // tiles must have been consumed first come from the ring plan in the descriptor (k_fts_ring_plan).
    const uint4 *dbase = a.desc;
    const uint32_t ring0 = smem_u32(&sm.ring[0]);
    int tile = blockIdx.x + pw * G;
    uint4 d_next = make_uint4(0u, 0u, 0u, 0u);
    if (tile < a.n_tiles && lane <= SB_FTT_MAXC) d_next = __ldg(dbase + (size_t)tile * (1 + SB_FTT_MAXC) + lane);
    for (int seq = pw; tile < a.n_tiles; tile += SB_FTS_PRODUCER_WARPS * G, seq += SB_FTS_PRODUCER_WARPS) {
        const uint4 d = d_next;
        if (tile + SB_FTS_PRODUCER_WARPS * G < a.n_tiles && lane <= SB_FTT_MAXC)       // one of this warp's tiles ahead
            d_next = __ldg(dbase + (size_t)(tile + SB_FTS_PRODUCER_WARPS * G) * (1 + SB_FTT_MAXC) + lane);
        const int stage = seq % SB_FTT_STAGES;
        const int nc = (int)(__shfl_sync(0xffffffffu, d.x, 0) & 3u);
        const uint32_t tile_base = ring0 + __shfl_sync(0xffffffffu, d.z, 0) * 128u;
        const int need = max((int)__shfl_sync(0xffffffffu, d.w, 0), seq - SB_FTT_STAGES + 1);
        if (need > 0) {                                 // tiles retire in order: waiting for tile need - 1 covers all before it
            const int t = need - 1;                     // (t >= seq - STAGES, so the barrier is at most one phase ahead)
            // A parity wait is only meaningful when the barrier is in the awaited phase or the one after it: this warp
            // must already know tile t - STAGES (the barrier's previous phase) consumed.  What it knows from its previous
            // tile (seq - P) is that tile seq - P - STAGES was consumed; if the plan lets t run ahead of seq - P, step
            // through that tile first (its barrier is then at most one phase behind, and t - STAGES <= seq - P).
            const int w = seq - SB_FTS_PRODUCER_WARPS;
            if (w >= 0 && t > w) mbar_wait(&sm.empty[w % SB_FTT_STAGES], (unsigned)(w / SB_FTT_STAGES) & 1u);
            mbar_wait(&sm.empty[t % SB_FTT_STAGES], (unsigned)(t / SB_FTT_STAGES) & 1u);
        }
        {
            uint4 ds = d;
            if (lane == 0) {                            // tile origin -> byte offsets of its first pixel in the panorama and the mask
                const unsigned X0 = d.y & 0xffffu, Y0 = d.y >> 16;
                ds.z = Y0 * ostep + X0 * obpp;
                ds.w = Y0 * mstep + X0;
            } else {
                ds.x = tile_base + ((d.z >> 16) << 7);
            }
            if (lane <= SB_FTT_MAXC) sm.desc[stage][lane] = ds;
            __syncwarp();
            // table blocks: one bulk copy (TMA) per camera slot, completion by expect_tx
            if (lane == 0) {
                if (nc == 0) mbar_arrive(&sm.full[stage]);
                else mbar_arrive_expect_tx(&sm.full[stage], (unsigned)nc * (unsigned)SB_FTT_TAB_BYTES);
            }
            __syncwarp();
            for (int k = 0; k < nc; ++k) {
                const unsigned dx_ = __shfl_sync(0xffffffffu, d.x, k + 1), dy_ = __shfl_sync(0xffffffffu, d.y, k + 1);
                const unsigned dw_ = __shfl_sync(0xffffffffu, d.w, k + 1), dz_ = __shfl_sync(0xffffffffu, d.z, k + 1);
                const auto &c = a.cam[dw_ & 15u];
                const uint32_t slot = tile_base + ((dz_ >> 16) << 7);
                const unsigned pitch = dz_ & 0xffffu;
                if (lane == 0)
                    bulk_g2s_addr(slot, c.tiles + (size_t)((dw_ >> 4) & 0x7ffffffu) * (SB_FTT_W * SB_FTT_H), SB_FTT_TAB_BYTES, &sm.full[stage]);
                unsigned n_rows = dy_ >> 24;
                if (n_rows == (unsigned)SB_FTS_DIRECT) n_rows = 0u;
                // source box: 16-byte cp.async chunks, all lanes (row length clamped to the pitch of the source image)
                const unsigned xlo = dx_ & 0xffffu, ylo = dx_ >> 16, need_end = dy_ & 0xffffffu;
                const unsigned cpr = (min((need_end + 15u) & ~15u, c.sstep) - xlo) >> 4;       // chunks per row
                const uint32_t box = slot + (uint32_t)SB_FTT_TAB_BYTES;
                const uint8_t *g = c.src + (size_t)ylo * c.sstep + xlo;
                if (cpr <= 32u) {
                    // lanes = (32 / cw rows) x (cw chunk columns), cw = the power of two >= cpr
                    const unsigned sh = cpr > 1u ? 32u - (unsigned)__clz((int)(cpr - 1u)) : 0u;
                    const unsigned col = (unsigned)lane & ((1u << sh) - 1u), rstep = 32u >> sh;
                    unsigned r = (unsigned)lane >> sh;
                    uint32_t dst = box + r * pitch + col * 16u;
                    const uint8_t *src = g + (size_t)r * c.sstep + col * 16u;
                    if (col < cpr)
                        for (; r < n_rows; r += rstep, dst += rstep * pitch, src += (size_t)rstep * c.sstep) cp_async_16(dst, src);
                } else {
                    const float inv = __frcp_rn((float)cpr);
                    const unsigned n_chunks = n_rows * cpr;
                    for (unsigned ch = lane; ch < n_chunks; ch += 32u) {
                        unsigned r = (unsigned)__float2int_rz(__fmul_rn((float)ch + 0.5f, inv));    // ch / cpr for ch < 1024 ...
                        if (r * cpr > ch) --r;                                                      // ... made exact
                        else if ((r + 1u) * cpr <= ch) ++r;
                        const unsigned col = ch - r * cpr;
                        cp_async_16(box + r * pitch + col * 16u, g + (size_t)r * c.sstep + col * 16u);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(&sm.full[stage]);    // one arrival per lane when its chunks have landed
        }
    }
}

}  // namespace sb
