// kernels_feather_tma.cu — the feather frame kernel as a persistent, warp-specialised streaming kernel.
//
// Same arithmetic as k_feather_fused_px1 (kernels_fused.cu): remap + gain + convertTo(16S) +
// FeatherBlender::feed over all cameras + blend + convertTo(8U), one launch per frame.  What changes
// is how the bytes move.  In the one-pixel-per-thread kernel every warp waits for its tile mask, then
// for its table entries, then for its taps: three dependent DRAM round trips, and the profile is
// latency-bound (long-scoreboard stalls, < 20 % of HBM bandwidth).  Here
//   * the panorama is cut into 32 x SB_FTT_H tiles; per (tile, camera) the sequence-constant table is
//     stored TILE-MAJOR as one contiguous block, and the bounding box of the source pixels the tile
//     samples is known per calibration (a 64-byte descriptor per tile);
//   * each CTA is persistent and walks tiles round-robin; PRODUCER warps stream, many tiles ahead, the
//     table blocks (one cp.async.bulk / TMA each, mbarrier complete_tx) and the source boxes (16-byte
//     cp.async chunks, lanes in parallel, mbarrier arrive.noinc) into a BYTE-GRANULAR shared-memory
//     ring (a tile takes what its boxes need, ~7 KB per camera instead of a fixed 16 KB slot);
//   * the CONSUMER warps compute pixels purely from shared memory (table entry -> box-relative word
//     offset + byte shift -> aligned LDS + funnel shift -> byte dot products against a shared-memory
//     copy of the bilinear weight table) and hand the stage back through an "empty" mbarrier.
//     No consumer thread ever waits on a global load.
#include <algorithm>
#include <climits>
#include <vector>

#include "sb_stream.cuh"

namespace sb {

#define SB_WEIGHT_EPS 1e-5f

// ------------------------------------------------------------------------------------ setup
// pass 1: per (camera, tile) the source bounding box of the weighted entries -> descriptor record
__global__ void __launch_bounds__(256)
k_fts_bbox(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, float sharpness, uint4 *rec)
{
    __shared__ int red[4][8];
    bool all_one = true;                                    // every pixel of the tile carries the full weight 1.0f
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
    for (int e = tid; e < SB_FTT_W * SB_FTT_H; e += blockDim.x) {
        const int lx = e % SB_FTT_W, ly = e / SB_FTT_W;
        const int x = (tx0 + (int)blockIdx.x) * SB_FTT_W + lx - dx, y = (ty0 + (int)blockIdx.y) * SB_FTT_H + ly - dy;
        if ((unsigned)x >= (unsigned)ww || (unsigned)y >= (unsigned)wh) { all_one = false; continue; }
        const uint2 t = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x];
        all_one = all_one && fminf(__fmul_rn((float)(t.y >> 16), sharpness), 1.f) == 1.f;
        if ((t.y >> 16) == 0u) continue;
        const int x0 = t.x & 0x1fff, y0 = (t.x >> 13) & 0x1fff;
        const int x1 = x0 + 1 - (int)((t.x >> 26) & 1u), y1 = y0 + 1 - (int)((t.x >> 27) & 1u);
        mnx = min(mnx, x0); mxx = max(mxx, x1); mny = min(mny, y0); mxy = max(mxy, y1);
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
    const int every = __syncthreads_and(all_one);
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            mnx = min(mnx, red[0][w]); mny = min(mny, red[1][w]); mxx = max(mxx, red[2][w]); mxy = max(mxy, red[3][w]);
        }
        uint4 r = make_uint4(0u, 0u, 0u, 0u);                // n_rows == 0: the camera carries no weight in this tile
        if (mxx >= mnx) {
            const unsigned xlo = ((unsigned)mnx * 3u) & ~15u;             // box start, 16-byte aligned
            const unsigned need_end = (unsigned)(mxx + 1) * 3u + 3u;      // last needed byte + the aligned-word slack of the 6-byte tap read
            const unsigned pitch = ((need_end + 15u) & ~15u) - xlo;
            unsigned n_rows = (unsigned)(mxy - mny + 1);
            if (n_rows > (unsigned)SB_FTS_MAX_ROWS || n_rows * pitch > (unsigned)SB_FTS_BOX_BYTES) n_rows = SB_FTS_DIRECT;
            r = make_uint4(xlo | ((unsigned)mny << 16), need_end | (n_rows << 24), pitch, every ? 0x80000000u : 0u);
        }
        rec[blockIdx.y * ntx + blockIdx.x] = r;
    }
}

// pass 2: tile-major entries, tap position relative to the tile's source box:
//   x = (box byte offset of tap (y0, x0) & 3) << 3 | (that offset >> 2) << 5 | (y1 == y0) << 27
//       (low 5 bits = the funnel-shift amount of the aligned-word tap read; box rows are 16-byte pitched, so the
//        second tap row shares it)
//   y = (fx | fy << 5) << 3 | dist << 16       (byte offset into the 8-byte bilinear weight table)
// An x-folded pair (x1 == x0 at the source edge) is stored with fx = 0: the weights of the second column are then
// zero, sum w*p is the same integer, and the consumer needs no special case (the bytes it multiplies by zero are
// inside the box pitch).
// Boxes too large for shared memory: x = x0 | y0 << 13 | flags, the row-major format, taps gathered directly.
__global__ void __launch_bounds__(256)
k_fts_entries(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, const uint4 *rec, uint2 *tiles)
{
    const uint4 r = rec[blockIdx.y * ntx + blockIdx.x];
    const unsigned xlo = r.x & 0xffffu, ylo = r.x >> 16, pitch = r.z;
    uint2 *dst = tiles + ((size_t)blockIdx.y * ntx + blockIdx.x) * (SB_FTT_W * SB_FTT_H);
    for (int e = threadIdx.x; e < SB_FTT_W * SB_FTT_H; e += blockDim.x) {
        const int lx = e % SB_FTT_W, ly = e / SB_FTT_W;
        const int x = (tx0 + (int)blockIdx.x) * SB_FTT_W + lx - dx, y = (ty0 + (int)blockIdx.y) * SB_FTT_H + ly - dy;
        uint2 t = make_uint2(0u, 0u);
        if ((unsigned)x < (unsigned)ww && (unsigned)y < (unsigned)wh) {
            const uint2 s = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x];
            if ((s.y >> 16) != 0u) {
                const unsigned x0 = s.x & 0x1fffu, y0 = (s.x >> 13) & 0x1fffu;
                unsigned fxy = s.y & 1023u;
                if ((r.y >> 24) == (unsigned)SB_FTS_DIRECT) {
                    t.x = s.x;
                } else {
                    const unsigned off = (y0 - ylo) * pitch + x0 * 3u - xlo;
                    t.x = ((off & 3u) << 3) | ((off >> 2) << 5) | (s.x & (1u << 27));
                    if (s.x & (1u << 26)) fxy &= ~31u;
                }
                t.y = (fxy << 3) | (s.y & 0xffff0000u);
            }
        }
        dst[e] = t;
    }
}

int launch_fts_camera_tiles(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                            float sharpness, uint4 *rec, uint2 *tiles, cudaStream_t s)
{
    k_fts_bbox<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, ww, wh, dx, dy, tx0, ty0, ntx, sharpness, rec);
    SB_LAUNCHED();
    k_fts_entries<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, ww, wh, dx, dy, tx0, ty0, ntx, rec, tiles);
    SB_LAUNCHED();
    return SB_OK;
}

// pass 3: one 64-byte descriptor per panorama tile:
//   [0]     = {n_cams | short path << 2 | ring units << 8, tile origin X0 | Y0 << 16, ring start unit, tiles to retire first}
//             (ring unit = 128 bytes; .z and .w are filled by k_fts_ring_plan)
//   [1 + k] = per camera slot (ascending camera index = feed order)
//             {xlo | ylo << 16, need_end | n_rows << 24, pitch | ring unit offset inside the tile << 16,
//              cam | table block index << 4 | full weight << 31}
// *status: bit 0 = some tile has more than SB_FTT_MAXC cameras, bit 1 = block index overflow.
__global__ void k_fts_descriptors(FtsSetup a, uint4 *desc, int *status)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.n_tiles) return;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    uint4 out[1 + SB_FTT_MAXC];
    for (int k = 0; k <= SB_FTT_MAXC; ++k) out[k] = make_uint4(0u, 0u, 0u, 0u);
    int k = 0, bad = 0;
    unsigned units = 0;
    for (int i = 0; i < a.n; ++i) {
        const int rx = tx - a.cam[i].tx0, ry = ty - a.cam[i].ty0;
        if ((unsigned)rx >= (unsigned)a.cam[i].ntx || (unsigned)ry >= (unsigned)a.cam[i].nty) continue;
        const unsigned block = (unsigned)(ry * a.cam[i].ntx + rx);
        uint4 r = a.cam[i].rec[block];
        const unsigned n_rows = r.y >> 24;
        if (n_rows == 0u) continue;
        if (block >= (1u << 27)) bad |= 2;
        if (k == SB_FTT_MAXC) { bad |= 1; break; }
        const unsigned box_bytes = n_rows == (unsigned)SB_FTS_DIRECT ? 0u : n_rows * r.z;
        r.z |= units << 16;
        units += ((unsigned)SB_FTT_TAB_BYTES + box_bytes + 127u) >> 7;
        r.w = (unsigned)i | (block << 4) | (r.w & 0x80000000u);
        out[1 + k++] = r;
    }
    // one camera at full weight over the whole tile, box staged in shared memory: the consumers take the short path
    const unsigned short_path = (k == 1 && (out[1].w >> 31) && (out[1].y >> 24) != (unsigned)SB_FTS_DIRECT) ? 1u : 0u;
    out[0].x = (unsigned)k | (short_path << 2) | (units << 8);
    out[0].y = (unsigned)(tx * SB_FTT_W) | ((unsigned)(ty * SB_FTT_H) << 16);
    for (int j = 0; j <= SB_FTT_MAXC; ++j) desc[(size_t)tile * (1 + SB_FTT_MAXC) + j] = out[j];
    if (bad) atomicOr(status, bad);
}

// pass 4 (after the host has put the tiles in schedule order): the shared-memory ring plan.  CTA b of a grid of G
// walks tiles b, b + G, ...; a tile takes `units` contiguous 128-byte ring units at the head, wrapping to 0 when the
// end of the ring is too short, and tiles retire in order.  Both are a pure function of the tile sizes, so the start
// unit of every tile and the number of the CTA's tiles that must have been consumed before its copies may land are
// computed once per calibration: [0].z = start unit, [0].w = tiles to retire first (also covers the reuse of the tile's
// stage entry).  The producer warps then need no shared bookkeeping at all.
__global__ void k_fts_ring_plan(uint4 *desc, int n_tiles, int G)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= G) return;
    int hist[SB_FTT_STAGES];
    int head = 0, tail = 0, oldest = 0, seq = 0;
    for (int tile = b; tile < n_tiles; tile += G, ++seq) {
        uint4 d0 = desc[(size_t)tile * (1 + SB_FTT_MAXC)];
        const int units = (int)(d0.x >> 8);
        int start;
        for (;;) {
            if (seq - oldest < SB_FTT_STAGES) {
                if (seq == oldest) head = tail = 0;     // nothing in flight
                if (head >= tail) {                     // in use: [tail, head)
                    if (head + units <= SB_FTS_RING_BYTES / 128) { start = head; break; }
                    if (units < tail) { start = 0; break; }
                } else if (head + units < tail) {       // in use: [tail, end) and [0, head)
                    start = head; break;
                }
            }
            ++oldest;
            tail = oldest < seq ? hist[oldest % SB_FTT_STAGES] : head;
        }
        head = start + units;
        hist[seq % SB_FTT_STAGES] = start;
        d0.z = (unsigned)start;
        d0.w = (unsigned)oldest;
        desc[(size_t)tile * (1 + SB_FTT_MAXC)] = d0;
    }
}

// Schedule order: CTA b takes positions b, b + G, ... of the descriptor array, so listing the tiles by descending cost
// (blended tiles with 3, 2, 1 cameras, then the single-camera short-path tiles, then empty ones; row-major inside a
// class) deals every CTA the same number of tiles of each class, the expensive ones first.
int fts_schedule(uint4 *desc, int n_tiles, int grid, cudaStream_t s)
{
    const size_t per = 1 + SB_FTT_MAXC;
    std::vector<uint4> h((size_t)n_tiles * per), o((size_t)n_tiles * per);
    SB_CUDA(cudaMemcpyAsync(h.data(), desc, h.size() * sizeof(uint4), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    std::vector<int> order(n_tiles);
    for (int t = 0; t < n_tiles; ++t) order[t] = t;
    auto cost = [&](int t) {
        const unsigned x = h[(size_t)t * per].x;
        const int nc = (int)(x & 3u);
        return nc == 0 ? 0 : (x & 4u) ? 1 : 1 + nc;
    };
    std::stable_sort(order.begin(), order.end(), [&](int l, int r) { return cost(l) > cost(r); });
    // ... except that every CTA's FIRST tile is a cheap one: the first wave of copies (grid x producer warps tiles at once)
    // is what the consumers wait for at kernel start, and a one-camera tile is 2-3x fewer bytes
    if (n_tiles >= 2 * grid) std::rotate(order.begin(), order.end() - grid, order.end());
    for (int t = 0; t < n_tiles; ++t)
        for (size_t j = 0; j < per; ++j) o[(size_t)t * per + j] = h[(size_t)order[t] * per + j];
    SB_CUDA(cudaMemcpyAsync(desc, o.data(), o.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
    k_fts_ring_plan<<<div_up(grid, 64), 64, 0, s>>>(desc, n_tiles, grid);
    SB_LAUNCHED();
    SB_CUDA(cudaStreamSynchronize(s));                      // (o goes out of scope)
    return SB_OK;
}

int launch_fts_descriptors(const FtsSetup &a, uint4 *desc, int *status, int grid, cudaStream_t s)
{
    k_fts_descriptors<<<div_up(a.n_tiles, 128), 128, 0, s>>>(a, desc, status);
    SB_LAUNCHED();
    return fts_schedule(desc, a.n_tiles, grid, s);
}
int fts_grid(int n_tiles, int sm_count) { return std::min(n_tiles, SB_FTS_CTAS_PER_SM * sm_count); }

// ------------------------------------------------------------------------------------ frame kernel
// (shared-memory layout, producer warps and tap fetch: sb_stream.cuh)
// exposure gain of camera `c` at panorama pixel (X, Y): the scalar of GainCompensator or the resized block map
__device__ __forceinline__ float fts_gain(const FeatherTmaCam &c, int X, int Y)
{
    return c.gmap ? __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.gmap) + (size_t)(Y - c.dy) * c.gmstep) + (X - c.dx)) : c.gain;
}


template <bool OUT8>
__device__ __forceinline__ void fts_store(unsigned char *o, int v0, int v1, int v2)
{
    if (OUT8) {
        o[0] = (uint8_t)v0; o[1] = (uint8_t)v1; o[2] = (uint8_t)v2;
    } else {
        short *q = reinterpret_cast<short *>(o);
        q[0] = (short)v0; q[1] = (short)v1; q[2] = (short)v2;
    }
}

// NOBLEND: Blender::feed / blend without blending (blenders.cpp:81-112): the pixel of the LAST camera (feed order)
// whose mask is non-zero, dst_mask = OR of the mask bytes, 0 where no camera has a mask.
template <bool GAIN, bool OUT8, bool NOBLEND>
__global__ void __launch_bounds__(SB_FTS_THREADS, SB_FTS_CTAS_PER_SM)
k_feather_stream(const __grid_constant__ FeatherTmaArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FtsSmem &sm = *reinterpret_cast<FtsSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < SB_FTT_STAGES; ++s) {
            mbar_init(&sm.full[s], 32 + 1);                 // the owning producer warp's lanes (cp.async) + the table copies' expect_tx
            mbar_init(&sm.empty[s], SB_FTS_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= SB_FTS_CONSUMER_WARPS) {
        // producer warps: sb_stream.cuh
        stream_producer(a, sm, warp - SB_FTS_CONSUMER_WARPS, lane, G, (unsigned)a.out_step, OUT8 ? 3u : 6u, (unsigned)a.mask_step);
        return;
    }

    // ---------------------------------------------------- consumer warps
    // warp w covers the 32 columns of tile rows w + RPP * p, p = 0 .. PX-1
    constexpr int RPP = SB_FTS_CONSUMER_WARPS, PX = SB_FTT_H / RPP;
    static_assert(SB_FTT_W == 32 && SB_FTT_H % RPP == 0, "one warp per tile row");
    constexpr uint32_t ROW_STRIDE = (uint32_t)RPP * SB_FTT_W * 8u;        // table bytes between a thread's pixels
    const int lx = lane, ly = warp;
    // the weight table is loaded by the consumer warps while the producers already fetch the first tiles
    for (int i = tid; i < 1024; i += SB_FTS_CONSUMER_WARPS * 32) sm.lut[i] = __ldg(a.bilin_lut + i);
    asm volatile("bar.sync 1, %0;" ::"n"(SB_FTS_CONSUMER_WARPS * 32) : "memory");
    const uint32_t lut0 = smem_u32(&sm.lut[0]);
    const uint32_t tab_off = (uint32_t)(ly * SB_FTT_W + lx) * 8u;
    const unsigned ostep = (unsigned)a.out_step, mstep = (unsigned)a.mask_step;      // (the launcher checks the panorama is < 4 GB)
    unsigned char *const out_t = reinterpret_cast<unsigned char *>(a.out) + ((unsigned)ly * ostep + (unsigned)lx * (OUT8 ? 3u : 6u));
    uint8_t *const mask_t = a.out_mask ? a.out_mask + ((unsigned)ly * mstep + (unsigned)lx) : nullptr;
    int stage = 0;
    unsigned parity = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += G) {
        mbar_wait(&sm.full[stage], parity);
        const uint4 d0 = sm.desc[stage][0];
        const int nc = (int)(d0.x & 3u);
        const int X = (int)(d0.y & 0xffffu) + lx, Y0 = (int)(d0.y >> 16) + ly;       // (only the gain map and the edge tiles use them)
        unsigned char *orow = out_t + d0.z;
        uint8_t *mrow_ = mask_t ? mask_t + d0.w : nullptr;
        if (d0.x & 4u) {
            // Short path (block-uniform; ~80 % of a ring panorama): ONE camera with weight exactly 1.0f on every pixel
            // of the tile.  Then dst = short(p * 1.0f) = p, dst_w = 1.0f, and normalizeUsingWeightMap gives
            // short(p / (1.0f + 1e-5f)) = p - 1 for p in 1..255 and 0 for p = 0 (the quotient lies strictly between
            // p - 1 and p); the mask is 255.  No float op is needed at all.
            const uint4 rec = sm.desc[stage][1];
            const unsigned pitch = rec.z & 0xffffu;
            const uint32_t tab = rec.x + tab_off, box = rec.x + (uint32_t)SB_FTT_TAB_BYTES;
            int v[PX][3];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const uint2 te = lds_u2(tab + (uint32_t)p * ROW_STRIDE);
                const uint2 bw = lds_u2(lut0 + (te.y & 0x1ff8u));
                unsigned lo0, hi0, lo1, hi1;
                fts_taps(box, pitch, te.x, lo0, hi0, lo1, hi1);
                fts_bilinear<!GAIN && !NOBLEND>(lo0, hi0, lo1, hi1, bw, v[p][0], v[p][1], v[p][2]);
                if (GAIN) {                                 // saturate_cast<uchar>(p * gain)
                    const float g = fts_gain(a.cam[rec.w & 15u], X, Y0 + RPP * p);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        v[p][k] = min(max(__float2int_rn(__fmul_rn((float)v[p][k], g)), 0), 255);
                        if (!NOBLEND) v[p][k] -= min(v[p][k], 1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                fts_store<OUT8>(orow, v[p][0], v[p][1], v[p][2]);
                if (mrow_) { *mrow_ = 255; mrow_ += RPP * mstep; }
                orow += RPP * ostep;
            }
            if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
            continue;
        }
        int acc[PX][3];
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0;
        if (NOBLEND) {
            // which camera slot supplies each pixel (the last one in feed order with a non-zero mask), OR of the masks
            int sel[PX];
            unsigned mor[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) { sel[p] = -1; mor[p] = 0u; }
            for (int k = 0; k < nc; ++k) {
                const uint32_t tab = sm.desc[stage][1 + k].x + tab_off;
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const unsigned m = lds_u2(tab + (uint32_t)p * ROW_STRIDE).y >> 16;
                    mor[p] |= m;
                    if (m) sel[p] = k;
                }
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (sel[p] < 0) continue;
                const uint4 rec = sm.desc[stage][1 + sel[p]];
                const uint2 te = lds_u2(rec.x + tab_off + (uint32_t)p * ROW_STRIDE);
                const uint2 bw = lds_u2(lut0 + (te.y & 0x1ff8u));
                unsigned lo0, hi0, lo1, hi1;
                if ((rec.y >> 24) != (unsigned)SB_FTS_DIRECT) fts_taps(rec.x + (uint32_t)SB_FTT_TAB_BYTES, rec.z & 0xffffu, te.x, lo0, hi0, lo1, hi1);
                else fts_taps_direct(a.cam[rec.w & 15u], te.x, lo0, hi0, lo1, hi1);
                bilinear_rgb(lo0, hi0, lo1, hi1, bw, acc[p][0], acc[p][1], acc[p][2]);
                if (GAIN) {
                    const float g = fts_gain(a.cam[rec.w & 15u], X, Y0 + RPP * p);
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[p][k] = min(max(__float2int_rn(__fmul_rn((float)acc[p][k], g)), 0), 255);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            if (X < a.pw) {
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    if (Y0 + RPP * p >= a.ph) break;
                    fts_store<OUT8>(orow, acc[p][0], acc[p][1], acc[p][2]);
                    if (mrow_) { *mrow_ = (uint8_t)mor[p]; mrow_ += RPP * mstep; }
                    orow += RPP * ostep;
                }
            }
            if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
            continue;
        }
        float wsum[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) wsum[p] = 0.f;
        for (int k = 0; k < nc; ++k) {                      // ascending camera index = feed order (float weight sums)
            const uint4 rec = sm.desc[stage][1 + k];
            const unsigned pitch = rec.z & 0xffffu;
            const bool direct = (rec.y >> 24) == (unsigned)SB_FTS_DIRECT;    // block-uniform
            const uint32_t tab = rec.x + tab_off, box = rec.x + (uint32_t)SB_FTT_TAB_BYTES;
            uint2 te[PX], bw[PX];
            unsigned lo0[PX], hi0[PX], lo1[PX], hi1[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                te[p] = lds_u2(tab + (uint32_t)p * ROW_STRIDE);
                bw[p] = lds_u2(lut0 + (te[p].y & 0x1ff8u));
                if (!direct) {
                    fts_taps(box, pitch, te[p].x, lo0[p], hi0[p], lo1[p], hi1[p]);
                } else {                                    // box too large for the ring: gather from global memory
                    lo0[p] = hi0[p] = lo1[p] = hi1[p] = 0u;
                    if ((te[p].y >> 16) != 0u) fts_taps_direct(a.cam[rec.w & 15u], te[p].x, lo0[p], hi0[p], lo1[p], hi1[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const unsigned dist = te[p].y >> 16;        // 0: short(p * 0) == 0 and dst_w += 0 -> contributes nothing
                const float w = fminf(__fmul_rn((float)dist, a.sharpness), 1.f);   // createWeightMap
                wsum[p] = __fadd_rn(wsum[p], w);
                int v0, v1, v2;
                bilinear_rgb(lo0[p], hi0[p], lo1[p], hi1[p], bw[p], v0, v1, v2);
                if (GAIN && dist != 0u) {                   // saturate_cast<uchar>(p * gain)
                    const float g = fts_gain(a.cam[rec.w & 15u], X, Y0 + RPP * p);
                    v0 = min(max(__float2int_rn(__fmul_rn((float)v0, g)), 0), 255);
                    v1 = min(max(__float2int_rn(__fmul_rn((float)v1, g)), 0), 255);
                    v2 = min(max(__float2int_rn(__fmul_rn((float)v2, g)), 0), 255);
                }
                acc[p][0] += __float2int_rz(__fmul_rn((float)v0, w));              // static_cast<short>(src * w)
                acc[p][1] += __float2int_rz(__fmul_rn((float)v1, w));
                acc[p][2] += __float2int_rz(__fmul_rn((float)v2, w));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);       // this warp no longer reads the stage
        // FeatherBlender::blend: normalizeUsingWeightMap, mask = weight > eps, zero unmasked, convertTo(8U)
        if (X < a.pw) {
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (Y0 + RPP * p >= a.ph) break;
                const int m = wsum[p] > SB_WEIGHT_EPS ? 255 : 0;
                const SharedDiv div(__fadd_rn(wsum[p], SB_WEIGHT_EPS));            // in [1e-5, n + 1e-5]: fast-path range
                const int o0 = m ? __float2int_rz(div((float)acc[p][0])) : 0, o1 = m ? __float2int_rz(div((float)acc[p][1])) : 0,
                          o2 = m ? __float2int_rz(div((float)acc[p][2])) : 0;
                fts_store<OUT8>(orow, o0, o1, o2);
                if (mrow_) { *mrow_ = (uint8_t)m; mrow_ += RPP * mstep; }
                orow += RPP * ostep;
            }
        }
        if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
    }
}

int launch_feather_stream(const FeatherTmaArgs &a, bool apply_gain, bool out8, int sm_count, cudaStream_t s)
{
    SB_ASSERT(a.sharpness > 0.f && a.bilin_lut && a.desc && a.n_tiles > 0 && a.n <= 16);
    SB_ASSERT(a.pw < 65536 && a.ph < 65536 && (unsigned long long)a.ph * a.out_step < (1ull << 32) && (unsigned long long)a.ph * a.mask_step < (1ull << 32));
    const size_t smem = sizeof(FtsSmem);
    static bool configured_dev[64][8] = {};                  // the attribute is per device (context)
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    bool *configured = configured_dev[dev & 63];
    const void *fn[8] = {(const void *)k_feather_stream<false, false, false>, (const void *)k_feather_stream<false, true, false>,
                         (const void *)k_feather_stream<true, false, false>, (const void *)k_feather_stream<true, true, false>,
                         (const void *)k_feather_stream<false, false, true>, (const void *)k_feather_stream<false, true, true>,
                         (const void *)k_feather_stream<true, false, true>, (const void *)k_feather_stream<true, true, true>};
    const int v = (a.no_blend ? 4 : 0) + (apply_gain ? 2 : 0) + (out8 ? 1 : 0);
    if (!configured[v]) {
        SB_CUDA(cudaFuncSetAttribute(fn[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[v] = true;
    }
    const int grid = fts_grid(a.n_tiles, sm_count);           // the schedule order and ring plan in a.desc were made for this grid
    void *params[] = {const_cast<FeatherTmaArgs *>(&a)};
    SB_CUDA(cudaLaunchKernel(fn[v], dim3(grid), dim3(SB_FTS_THREADS), params, smem, s));
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
