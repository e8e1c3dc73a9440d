// kernels_mb_stream.cu — the multi-band path's first stage (remap + gain + convertTo(16S) + copyMakeBorder for every
// camera, Gaussian level 0 as RGBX bytes) on the streaming machinery of sb_stream.cuh.
//
// k_mb_warp (kernels_mb.cu) gathers the four taps of every padded-rect pixel straight from global memory: the profile
// is long-scoreboard bound (two dependent round trips, ~120 instructions per pixel).  Here the padded rect of each
// camera is cut into 32 x SB_FTT_H tiles; per tile the resolved-tap table is stored tile-major as one contiguous block
// and the bounding box of the source pixels the tile samples is known per calibration.  Producer warps stream both
// into the shared-memory ring, the consumer warps compute from shared memory and write whole 128-byte rows of RGBX.
// Same arithmetic, bit for bit (cv::remap's fixed-point bilinear core of sb_device.cuh).
#include <algorithm>
#include <climits>

#include "sb_mb.h"
#include "sb_stream.cuh"

namespace sb {

// resolved taps of one padded-rect pixel (k_mb_tap_table): x0 | x1 << 12 | fx << 24, y0 | y1 << 12 | fy << 24.
// BORDER_REFLECT of the source may fold a pair (x1 == x0) or reverse it (x1 == x0 - 1); both are expressed as a pair
// starting at `b` with adjusted fractions: folded or weightless second column -> fraction 0; folded or weightless
// second row -> the row is read twice (yfold); reversed -> start at x1 / y1 with fraction 32 - f.
struct MbsTap { unsigned bx, by, fx, fy, yfold; };
__device__ __forceinline__ MbsTap mbs_resolve(uint2 t)
{
    const unsigned x0 = t.x & 0xfffu, x1 = (t.x >> 12) & 0xfffu, fx = t.x >> 24;
    const unsigned y0 = t.y & 0xfffu, y1 = (t.y >> 12) & 0xfffu, fy = t.y >> 24;
    MbsTap r;
    if (x1 == x0 || fx == 0u) { r.bx = x0; r.fx = 0u; }              // (the second column is read with weight 0: inside the row pitch)
    else if (x1 == x0 + 1u) { r.bx = x0; r.fx = fx; }
    else { r.bx = x1; r.fx = 32u - fx; }
    if (y1 == y0 || fy == 0u) { r.by = y0; r.fy = fy; r.yfold = 1u; }   // second row = first row: the two row weights add up
    else if (y1 == y0 + 1u) { r.by = y0; r.fy = fy; r.yfold = 0u; }
    else { r.by = y1; r.fy = 32u - fy; r.yfold = 0u; }
    return r;
}

// pass 1: per tile of the camera's padded rect, the bounding box of the source pixels it samples -> box record
__global__ void __launch_bounds__(256) k_mbs_bbox(const uint2 *table, size_t tstep, int rw, int rh, int ntx, uint4 *rec)
{
    __shared__ int red[4][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
    for (int e = tid; e < SB_FTT_W * SB_FTT_H; e += blockDim.x) {
        const int x = (int)blockIdx.x * SB_FTT_W + e % SB_FTT_W, y = (int)blockIdx.y * SB_FTT_H + e / SB_FTT_W;
        if (x >= rw || y >= rh) continue;
        const MbsTap t = mbs_resolve(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x]);
        const int x1 = (int)t.bx + (t.fx ? 1 : 0), y1 = (int)t.by + (t.yfold ? 0 : 1);
        mnx = min(mnx, (int)t.bx); mxx = max(mxx, x1); mny = min(mny, (int)t.by); mxy = max(mxy, y1);
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            mnx = min(mnx, red[0][w]); mny = min(mny, red[1][w]); mxx = max(mxx, red[2][w]); mxy = max(mxy, red[3][w]);
        }
        const unsigned xlo = ((unsigned)mnx * 3u) & ~15u;                 // box start, 16-byte aligned
        const unsigned need_end = (unsigned)(mxx + 1) * 3u + 3u;          // last needed byte + the aligned-word slack of the 6-byte tap read
        const unsigned pitch = ((need_end + 15u) & ~15u) - xlo;
        unsigned n_rows = (unsigned)(mxy - mny + 1);
        if (n_rows > (unsigned)SB_FTS_MAX_ROWS || n_rows * pitch > (unsigned)SB_FTS_BOX_BYTES) n_rows = SB_FTS_DIRECT;
        rec[blockIdx.y * ntx + blockIdx.x] = make_uint4(xlo | ((unsigned)mny << 16), need_end | (n_rows << 24), pitch, 0u);
    }
}

// pass 2: tile-major entries in the boxed format of sb_stream.cuh (fts_taps); tiles whose box is too large keep the
// row-major tap format and are gathered directly
__global__ void __launch_bounds__(256) k_mbs_entries(const uint2 *table, size_t tstep, int rw, int rh, int ntx, const uint4 *rec, uint2 *tiles)
{
    const uint4 r = rec[blockIdx.y * ntx + blockIdx.x];
    const unsigned xlo = r.x & 0xffffu, ylo = r.x >> 16, pitch = r.z;
    const bool direct = (r.y >> 24) == (unsigned)SB_FTS_DIRECT;
    uint2 *dst = tiles + ((size_t)blockIdx.y * ntx + blockIdx.x) * (SB_FTT_W * SB_FTT_H);
    for (int e = threadIdx.x; e < SB_FTT_W * SB_FTT_H; e += blockDim.x) {
        const int x = (int)blockIdx.x * SB_FTT_W + e % SB_FTT_W, y = (int)blockIdx.y * SB_FTT_H + e / SB_FTT_W;
        uint2 t = make_uint2(0u, 0u);
        if (x < rw && y < rh) {
            const uint2 s = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x];
            if (direct) {
                t = s;
            } else {
                const MbsTap m = mbs_resolve(s);
                const unsigned off = (m.by - ylo) * pitch + m.bx * 3u - xlo;
                t.x = ((off & 3u) << 3) | ((off >> 2) << 5) | (m.yfold << 27);
                t.y = (m.fx | (m.fy << 5)) << 3;
            }
        }
        dst[e] = t;
    }
}

int launch_mbs_camera_tiles(const uint2 *table, size_t tstep, int rw, int rh, int ntx, int nty, uint4 *rec, uint2 *tiles, cudaStream_t s)
{
    k_mbs_bbox<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, rw, rh, ntx, rec);
    SB_LAUNCHED();
    k_mbs_entries<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, rw, rh, ntx, rec, tiles);
    SB_LAUNCHED();
    return SB_OK;
}

// The padded rect's resolved taps in the row-major format the k_fs2 setup reads (kernels_fused.cu: k_build_feather_table):
//   x = x0 | y0 << 13 | (no second column) << 26 | (no second row) << 27,   y = fx | fy << 5 | 255 << 16 inside the wanted
// column runs cx (a zero "distance" elsewhere: the pixel is not produced).  The folded and reversed pairs of BORDER_REFLECT
// become plain pairs exactly as in mbs_resolve: the bilinear products are 32*a*b, so a row or column read twice carries the
// sum of its two weights, which is the weight of fraction 0.
__global__ void __launch_bounds__(256) k_mbs_feather_format(const uint2 *table, size_t tstep, int rw, int rh, int c0, int c1, int c2, int c3,
                                                            uint2 *out, size_t ostep)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= rw || y >= rh) return;
    const MbsTap m = mbs_resolve(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x]);
    const bool wanted = (x >= c0 && x < c1) || (x >= c2 && x < c3);
    uint2 t;
    t.x = m.bx | (m.by << 13) | (m.fx == 0u ? 1u << 26 : 0u) | (m.yfold << 27);
    t.y = m.fx | ((m.yfold ? 0u : m.fy) << 5) | (wanted ? 255u << 16 : 0u);
    reinterpret_cast<uint2 *>(reinterpret_cast<char *>(out) + (size_t)y * ostep)[x] = t;
}

int launch_mbs_feather_format(const uint2 *table, size_t tstep, int rw, int rh, const int cx[4], uint2 *out, size_t ostep, cudaStream_t s)
{
    k_mbs_feather_format<<<dim3(div_up(rw, 32), div_up(rh, 8)), 256, 0, s>>>(table, tstep, rw, rh, cx[0], cx[1], cx[2], cx[3], out, ostep);
    SB_LAUNCHED();
    return SB_OK;
}

// pass 3: one 64-byte descriptor per listed tile (list[t] = camera | tile block index << 4), the layout of
// k_fts_descriptors with exactly one camera slot; tile origin in rect-local pixels
__global__ void k_mbs_descriptors(MbsSetup a, const unsigned *list, int n_tiles, uint4 *desc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const unsigned cam = list[t] & 15u, block = list[t] >> 4;
    uint4 r = a.rec[cam][block];
    const unsigned n_rows = r.y >> 24;
    const unsigned box_bytes = n_rows == (unsigned)SB_FTS_DIRECT ? 0u : n_rows * r.z;
    const unsigned units = ((unsigned)SB_FTT_TAB_BYTES + box_bytes + 127u) >> 7;
    r.w = cam | (block << 4);
    const unsigned tx = block % (unsigned)a.ntx[cam], ty = block / (unsigned)a.ntx[cam];
    uint4 *d = desc + (size_t)t * (1 + SB_FTT_MAXC);
    d[0] = make_uint4(1u | (units << 8), (tx * SB_FTT_W) | ((ty * SB_FTT_H) << 16), 0u, 0u);
    d[1] = r;
    for (int j = 2; j <= SB_FTT_MAXC; ++j) d[j] = make_uint4(0u, 0u, 0u, 0u);
}

int launch_mbs_descriptors(const MbsSetup &a, const unsigned *list_dev, int n_tiles, uint4 *desc, int grid, cudaStream_t s)
{
    k_mbs_descriptors<<<div_up(n_tiles, 128), 128, 0, s>>>(a, list_dev, n_tiles, desc);
    SB_LAUNCHED();
    return fts_schedule(desc, n_tiles, grid, s);
}

// ------------------------------------------------------------------------------------ frame kernel
template <bool GAIN>
__global__ void __launch_bounds__(SB_FTS_THREADS, SB_FTS_CTAS_PER_SM)
k_mb_warp_stream(const __grid_constant__ MbStreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FtsSmem &sm = *reinterpret_cast<FtsSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < SB_FTT_STAGES; ++s) {
            mbar_init(&sm.full[s], 32 + 1);
            mbar_init(&sm.empty[s], SB_FTS_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp >= SB_FTS_CONSUMER_WARPS) {
        stream_producer(a, sm, warp - SB_FTS_CONSUMER_WARPS, lane, G, 0u, 0u, 0u);
        return;
    }
    constexpr int RPP = SB_FTS_CONSUMER_WARPS, PX = SB_FTT_H / RPP;
    constexpr uint32_t ROW_STRIDE = (uint32_t)RPP * SB_FTT_W * 8u;
    const int lx = lane, ly = warp;
    for (int i = tid; i < 1024; i += SB_FTS_CONSUMER_WARPS * 32) sm.lut[i] = __ldg(a.bilin_lut + i);
    asm volatile("bar.sync 1, %0;" ::"n"(SB_FTS_CONSUMER_WARPS * 32) : "memory");
    const uint32_t lut0 = smem_u32(&sm.lut[0]);
    const uint32_t tab_off = (uint32_t)(ly * SB_FTT_W + lx) * 8u;
    int stage = 0;
    unsigned parity = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += G) {
        mbar_wait(&sm.full[stage], parity);
        const uint4 d0 = sm.desc[stage][0];
        const uint4 rec = sm.desc[stage][1];
        const MbStreamCam &c = a.cam[rec.w & 15u];
        const int X = (int)(d0.y & 0xffffu) + lx, Y0 = (int)(d0.y >> 16) + ly;
        const unsigned pitch = rec.z & 0xffffu;
        const bool direct = (rec.y >> 24) == (unsigned)SB_FTS_DIRECT;        // block-uniform
        const uint32_t tab = rec.x + tab_off, box = rec.x + (uint32_t)SB_FTT_TAB_BYTES;
        unsigned out[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const uint2 te = lds_u2(tab + (uint32_t)p * ROW_STRIDE);
            unsigned lo0, hi0, lo1, hi1;
            uint2 bw;
            if (!direct) {
                bw = lds_u2(lut0 + (te.y & 0x1ff8u));
                fts_taps(box, pitch, te.x, lo0, hi0, lo1, hi1);
            } else {                                        // box too large for the ring: the row-major tap format, gathered from global memory
                const unsigned x0 = te.x & 0xfffu, x1 = (te.x >> 12) & 0xfffu, y0 = te.y & 0xfffu, y1 = (te.y >> 12) & 0xfffu;
                bw = lds_u2(lut0 + (((te.x >> 24) | ((te.y >> 24) << 5)) << 3));
                load_tap_row(c.src + (size_t)y0 * c.sstep, x0, x1, lo0, hi0);
                load_tap_row(c.src + (size_t)y1 * c.sstep, x0, x1, lo1, hi1);
            }
            int v0, v1, v2;
            bilinear_rgb(lo0, hi0, lo1, hi1, bw, v0, v1, v2);
            if (GAIN) {                                     // saturate_cast<uchar>(p * gain): GainCompensator / BlocksGainCompensator
                const int Y = min(Y0 + RPP * p, c.rh - 1), Xc = min(X, c.rw - 1);
                const float g = c.gmap ? __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.gmap) + (size_t)Y * c.gmstep) + Xc) : c.gain;
                v0 = min(max(__float2int_rn(__fmul_rn((float)v0, g)), 0), 255);
                v1 = min(max(__float2int_rn(__fmul_rn((float)v1, g)), 0), 255);
                v2 = min(max(__float2int_rn(__fmul_rn((float)v2, g)), 0), 255);
            }
            out[p] = (unsigned)v0 | ((unsigned)v1 << 8) | ((unsigned)v2 << 16);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        if (X < c.rw) {
            uint32_t *o = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(c.g0) + (size_t)Y0 * c.gstep) + X;
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (Y0 + RPP * p >= c.rh) break;
                *o = out[p];
                o = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(o) + (size_t)RPP * c.gstep);
            }
        }
        if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
    }
}

int launch_mb_warp_stream(const MbStreamArgs &a, bool apply_gain, int sm_count, cudaStream_t s)
{
    SB_ASSERT(a.bilin_lut && a.desc && a.n_tiles > 0 && a.n <= 16);
    const size_t smem = sizeof(FtsSmem);
    static bool configured_dev[64][2] = {};
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    const void *fn[2] = {(const void *)k_mb_warp_stream<false>, (const void *)k_mb_warp_stream<true>};
    const int v = apply_gain ? 1 : 0;
    if (!configured_dev[dev & 63][v]) {
        SB_CUDA(cudaFuncSetAttribute(fn[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev[dev & 63][v] = true;
    }
    void *params[] = {const_cast<MbStreamArgs *>(&a)};
    SB_CUDA(cudaLaunchKernel(fn[v], dim3(fts_grid(a.n_tiles, sm_count)), dim3(SB_FTS_THREADS), params, smem, s));
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
