// kernels_feather_tma.cu — the feather frame kernel as a persistent, warp-specialised streaming kernel.
//
// Same arithmetic as k_feather_fused_px1 (kernels_fused.cu): remap + gain + convertTo(16S) +
// FeatherBlender::feed over all cameras + blend + convertTo(8U), one launch per frame.  What changes
// is how the bytes move.  In the one-pixel-per-thread kernel every warp waits for its tile mask, then
// for its table entries, then for its taps: three dependent DRAM round trips, and the profile is
// latency-bound (long-scoreboard stalls, < 20 % of HBM bandwidth).  Here
//   * the panorama is cut into 32x16 tiles; per (tile, camera) the sequence-constant table is stored
//     TILE-MAJOR as one contiguous 4 KB block, and the bounding box of the source pixels the tile
//     samples is known per calibration (a 64-byte descriptor per tile);
//   * each CTA is persistent (one per SM) and walks tiles round-robin; a PRODUCER warp streams, three
//     tiles ahead, the table blocks (one cp.async.bulk / TMA each, mbarrier complete_tx) and the source
//     boxes (16-byte cp.async chunks, lanes in parallel, mbarrier arrive.noinc) into a 4-stage ring;
//   * 16 CONSUMER warps compute pixels purely from shared memory (table entry -> box-relative tap
//     offset -> aligned LDS + funnel shift -> byte dot products) and hand the stage back through an
//     "empty" mbarrier.  No thread ever waits on a global load.
#include <algorithm>
#include <climits>

#include "sb_device.cuh"
#include "sb_fused.h"
#include "sb_tma.cuh"

namespace sb {
using namespace sbd;
using namespace sbt;

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
            const unsigned need_end = (unsigned)(mxx + 1) * 3u + 3u;      // last needed byte + the aligned-word slack of load_6bytes
            const unsigned pitch = ((need_end + 15u) & ~15u) - xlo;
            unsigned n_rows = (unsigned)(mxy - mny + 1);
            if (n_rows > (unsigned)SB_FTS_MAX_ROWS || n_rows * pitch > (unsigned)SB_FTS_BOX_BYTES) n_rows = SB_FTS_DIRECT;
            r = make_uint4(xlo | ((unsigned)mny << 16), need_end | (n_rows << 24), pitch, every ? 0x80000000u : 0u);
        }
        rec[blockIdx.y * ntx + blockIdx.x] = r;
    }
}

// pass 2: tile-major entries, tap position relative to the tile's source box:
//   x = box byte offset of tap (y0, x0) | (x1 == x0) << 26 | (y1 == y0) << 27     y = fx | fy << 5 | dist << 16
// (boxes too large for a shared-memory slot: x = x0 | y0 << 13 | flags, the row-major format, taps gathered directly)
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
                t.x = (r.y >> 24) == (unsigned)SB_FTS_DIRECT ? s.x : (((y0 - ylo) * pitch + x0 * 3u - xlo) | (s.x & (3u << 26)));
                t.y = s.y;
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

// pass 3: one 64-byte descriptor per panorama tile: {n_cams, 0, 0, 0} + per camera slot (ascending
// camera index = feed order) {xlo | ylo << 16, need_end | n_rows << 24, pitch, cam | table block index << 4 | full weight << 31}.
// *status: bit 0 = some tile has more than SB_FTT_MAXC cameras, bit 1 = block index overflow.
__global__ void k_fts_descriptors(FtsSetup a, uint4 *desc, int *status)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.n_tiles) return;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    uint4 out[1 + SB_FTT_MAXC];
    for (int k = 0; k <= SB_FTT_MAXC; ++k) out[k] = make_uint4(0u, 0u, 0u, 0u);
    int k = 0, bad = 0;
    for (int i = 0; i < a.n; ++i) {
        const int rx = tx - a.cam[i].tx0, ry = ty - a.cam[i].ty0;
        if ((unsigned)rx >= (unsigned)a.cam[i].ntx || (unsigned)ry >= (unsigned)a.cam[i].nty) continue;
        const unsigned block = (unsigned)(ry * a.cam[i].ntx + rx);
        uint4 r = a.cam[i].rec[block];
        const unsigned n_rows = r.y >> 24;
        if (n_rows == 0u) continue;
        if (block >= (1u << 27)) bad |= 2;
        if (k == SB_FTT_MAXC) { bad |= 1; break; }
        r.w = (unsigned)i | (block << 4) | (r.w & 0x80000000u);
        out[1 + k++] = r;
    }
    out[0].x = (unsigned)k;
    // one camera at full weight over the whole tile, box staged in shared memory: the consumers take the short path
    out[0].z = (k == 1 && (out[1].w >> 31) && (out[1].y >> 24) != (unsigned)SB_FTS_DIRECT) ? 1u : 0u;
    for (int j = 0; j <= SB_FTT_MAXC; ++j) desc[(size_t)tile * (1 + SB_FTT_MAXC) + j] = out[j];
    if (bad) atomicOr(status, bad);
}

int launch_fts_descriptors(const FtsSetup &a, uint4 *desc, int *status, cudaStream_t s)
{
    k_fts_descriptors<<<div_up(a.n_tiles, 128), 128, 0, s>>>(a, desc, status);
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ frame kernel
// Shared memory: a ring of SB_FTS_SLOTS camera slots (table block + source box) allocated to tiles in
// order, as many as the tile has cameras (0..SB_FTT_MAXC; 1.2 on average), and a ring of SB_FTT_STAGES
// tile entries (descriptor + full/empty barriers).  ~10 tiles are in flight per CTA.
struct FtsSlot {
    uint2 tab[SB_FTT_H][SB_FTT_W];                          // 4 KB
    unsigned char box[SB_FTS_BOX_BYTES];
};
struct FtsSmem {
    FtsSlot slot[SB_FTS_SLOTS];
    uint4 desc[SB_FTT_STAGES][1 + SB_FTT_MAXC];             // [0] = {n_cams, first slot, 0, 0}
    uint64_t full[SB_FTT_STAGES], empty[SB_FTT_STAGES];
    int nc_hist[SB_FTS_PRODUCER_WARPS][SB_FTT_STAGES];      // per producer warp: cameras of its tiles in flight
};

// exposure gain of camera `c` at panorama pixel (X, Y): the scalar of GainCompensator or the resized block map
__device__ __forceinline__ float fts_gain(const FeatherTmaCam &c, int X, int Y)
{
    return c.gmap ? __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.gmap) + (size_t)(Y - c.dy) * c.gmstep) + (X - c.dx)) : c.gain;
}

// lo = bytes 0..3, hi = bytes 4..7 of the 6 tap bytes at shared-memory byte address `a`
__device__ __forceinline__ void lds_6bytes(uint32_t a, unsigned &lo, unsigned &hi)
{
    const unsigned o = a & 3u;
    const uint32_t b = a - o;
    unsigned w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(b));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(b));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(b));
    lo = __funnelshift_r(w0, w1, o * 8);
    hi = __funnelshift_r(w1, w2, o * 8);
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
        // ------------------------------------------------ producer warps
        const int pw = warp - SB_FTS_CONSUMER_WARPS;
        // Tile seq (0, 1, 2 ... of this CTA) is fetched by producer warp seq % PRODUCER_WARPS, so the per-tile issue
        // latency (descriptor shuffles, barrier ops, address arithmetic: ~1.4 us of dependent instructions) overlaps
        // across warps.  Every warp tracks the slot ring for ALL tiles (it only needs each tile's camera count).
        int stage = 0, head = 0, used = 0;                  // tile entry of this tile; next free slot; slots held by tiles in flight
        int seq = 0, oldest = 0;                            // tiles issued / retired by this CTA
        const uint4 *dbase = a.desc;
        unsigned nc_next = __ldg(&dbase[(size_t)blockIdx.x * (1 + SB_FTT_MAXC)].x);   // camera count, one tile ahead
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += G, ++seq) {
            const int nc = (int)nc_next;
            if (tile + G < a.n_tiles) nc_next = __ldg(&dbase[(size_t)(tile + G) * (1 + SB_FTT_MAXC)].x);
            const bool mine = seq % SB_FTS_PRODUCER_WARPS == pw;
            uint4 d = make_uint4(0u, 0u, 0u, 0u);
            if (mine && lane <= SB_FTT_MAXC) d = __ldg(dbase + (size_t)tile * (1 + SB_FTT_MAXC) + lane);
            // retire the oldest tiles until a tile entry and nc slots are free (tiles complete in order)
            while (seq - oldest >= SB_FTT_STAGES || used + nc > SB_FTS_SLOTS) {
                const int e = oldest % SB_FTT_STAGES;
                mbar_wait(&sm.empty[e], (unsigned)(oldest / SB_FTT_STAGES) & 1u);   // every consumer warp is done with it
                used -= sm.nc_hist[pw][e];
                ++oldest;
            }
            sm.nc_hist[pw][stage] = nc;                     // (every lane stores the same value)
            if (mine) {
                if (lane == 0) d.y = (unsigned)head;
                if (lane <= SB_FTT_MAXC) sm.desc[stage][lane] = d;
                // table blocks: one bulk copy (TMA) per camera slot, completion by expect_tx
                if (lane == 0) {
                    if (nc == 0) mbar_arrive(&sm.full[stage]);
                    else mbar_arrive_expect_tx(&sm.full[stage], (unsigned)nc * (unsigned)(SB_FTT_W * SB_FTT_H * sizeof(uint2)));
                }
                __syncwarp();
                for (int k = 0; k < nc; ++k) {
                    const unsigned dx_ = __shfl_sync(0xffffffffu, d.x, k + 1), dy_ = __shfl_sync(0xffffffffu, d.y, k + 1);
                    const unsigned dw_ = __shfl_sync(0xffffffffu, d.w, k + 1), pitch = __shfl_sync(0xffffffffu, d.z, k + 1);
                    const FeatherTmaCam &c = a.cam[dw_ & 15u];
                    const int slot = head + k < SB_FTS_SLOTS ? head + k : head + k - SB_FTS_SLOTS;
                    if (lane == 0)
                        bulk_g2s(&sm.slot[slot].tab[0][0], c.tiles + (size_t)((dw_ >> 4) & 0x7ffffffu) * (SB_FTT_W * SB_FTT_H),
                                 SB_FTT_W * SB_FTT_H * sizeof(uint2), &sm.full[stage]);
                    unsigned n_rows = dy_ >> 24;
                    if (n_rows == (unsigned)SB_FTS_DIRECT) n_rows = 0u;
                    // source box: 16-byte cp.async chunks, all lanes (row length clamped to the pitch of the source image)
                    const unsigned xlo = dx_ & 0xffffu, ylo = dx_ >> 16, need_end = dy_ & 0xffffffu;
                    const unsigned cpr = (min((need_end + 15u) & ~15u, c.sstep) - xlo) >> 4;       // chunks per row
                    const float inv = __frcp_rn((float)cpr);
                    const unsigned n_chunks = n_rows * cpr;
                    const uint32_t box = smem_u32(&sm.slot[slot].box[0]);
                    const uint8_t *g = c.src + (size_t)ylo * c.sstep + xlo;
                    for (unsigned ch = lane; ch < n_chunks; ch += 32u) {
                        unsigned r = (unsigned)__float2int_rz(__fmul_rn((float)ch + 0.5f, inv));    // ch / cpr for ch < 1024 ...
                        if (r * cpr > ch) --r;                                                      // ... made exact
                        else if ((r + 1u) * cpr <= ch) ++r;
                        const unsigned col = ch - r * cpr;
                        cp_async_16(box + r * pitch + col * 16u, g + (size_t)r * c.sstep + col * 16u);
                    }
                }
                cp_async_mbar_arrive_noinc(&sm.full[stage]);    // one arrival per lane when its chunks have landed
            }
            head = head + nc < SB_FTS_SLOTS ? head + nc : head + nc - SB_FTS_SLOTS;
            used += nc;
            if (++stage == SB_FTT_STAGES) stage = 0;
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    // warp w covers columns (w % SEGS) * 32 .. +31 of tile rows (w / SEGS) + RPP * p, p = 0 .. PX-1
    constexpr int SEGS = SB_FTT_W / 32, RPP = SB_FTS_CONSUMER_WARPS / SEGS, PX = SB_FTT_H / RPP;
    const int lx = (warp % SEGS) * 32 + lane, ly = warp / SEGS;
    const int gx = G % a.tiles_x, gy = G / a.tiles_x;       // tile -> (tx, ty) advanced incrementally
    int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
    int stage = 0;
    unsigned parity = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += G) {
        mbar_wait(&sm.full[stage], parity);
        const int nc = (int)sm.desc[stage][0].x, first = (int)sm.desc[stage][0].y;
        const int X = tx * SB_FTT_W + lx, Y0 = ty * SB_FTT_H + ly;
        if (sm.desc[stage][0].z) {
            // Short path (block-uniform; ~80 % of a ring panorama): ONE camera with weight exactly 1.0f on every pixel
            // of the tile.  Then dst = short(p * 1.0f) = p, dst_w = 1.0f, and normalizeUsingWeightMap gives
            // short(p / (1.0f + 1e-5f)) = p - 1 for p in 1..255 and 0 for p = 0 (the quotient lies strictly between
            // p - 1 and p); the mask is 255.  No float op is needed at all.
            const uint4 rec = sm.desc[stage][1];
            const unsigned pitch = rec.z;
            const FtsSlot &sl = sm.slot[first];
            const uint32_t box = smem_u32(&sl.box[0]);
            int v[PX][3];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const uint2 te = sl.tab[ly + RPP * p][lx];
                const uint2 bw = __ldg(a.bilin_lut + (te.y & 1023u));
                const uint32_t r0 = box + (te.x & 0x3ffffu);
                const uint32_t r1 = (te.x & (1u << 27)) ? r0 : r0 + pitch;
                unsigned lo0, hi0, lo1, hi1;
                lds_6bytes(r0, lo0, hi0);
                lds_6bytes(r1, lo1, hi1);
                if (te.x & (1u << 26)) {                    // x1 == x0 at the image edge: repeat the pixel
                    hi0 = lo0 >> 8; lo0 = (lo0 & 0x00ffffffu) | (lo0 << 24);
                    hi1 = lo1 >> 8; lo1 = (lo1 & 0x00ffffffu) | (lo1 << 24);
                }
                bilinear_rgb(lo0, hi0, lo1, hi1, bw, v[p][0], v[p][1], v[p][2]);
                if (GAIN) {                                 // saturate_cast<uchar>(p * gain)
                    const float g = fts_gain(a.cam[rec.w & 15u], X, Y0 + RPP * p);
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[p][k] = min(max(__float2int_rn(__fmul_rn((float)v[p][k], g)), 0), 255);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            unsigned char *orow = reinterpret_cast<unsigned char *>(a.out) + (size_t)Y0 * a.out_step + (size_t)X * (OUT8 ? 3 : 6);
            uint8_t *mrow_ = a.out_mask ? a.out_mask + (size_t)Y0 * a.mask_step + X : nullptr;
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const int o0 = NOBLEND ? v[p][0] : v[p][0] - min(v[p][0], 1), o1 = NOBLEND ? v[p][1] : v[p][1] - min(v[p][1], 1),
                          o2 = NOBLEND ? v[p][2] : v[p][2] - min(v[p][2], 1);
                if (OUT8) {
                    orow[0] = (uint8_t)o0; orow[1] = (uint8_t)o1; orow[2] = (uint8_t)o2;
                } else {
                    short *o = reinterpret_cast<short *>(orow);
                    o[0] = (short)o0; o[1] = (short)o1; o[2] = (short)o2;
                }
                if (mrow_) { *mrow_ = 255; mrow_ += RPP * a.mask_step; }
                orow += RPP * a.out_step;
            }
            tx += gx; ty += gy;
            if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
            if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
            continue;
        }
        int acc[PX][3];
        float wsum[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) { acc[p][0] = acc[p][1] = acc[p][2] = 0; wsum[p] = 0.f; }
        if (NOBLEND) {
            // which camera slot supplies each pixel (the last one in feed order with a non-zero mask), OR of the masks
            int sel[PX];
            unsigned mor[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) { sel[p] = -1; mor[p] = 0u; }
            for (int k = 0; k < nc; ++k) {
                const FtsSlot &sl = sm.slot[first + k < SB_FTS_SLOTS ? first + k : first + k - SB_FTS_SLOTS];
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const unsigned m = sl.tab[ly + RPP * p][lx].y >> 16;
                    mor[p] |= m;
                    if (m) sel[p] = k;
                }
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (sel[p] < 0) continue;
                const uint4 rec = sm.desc[stage][1 + sel[p]];
                const FtsSlot &sl = sm.slot[first + sel[p] < SB_FTS_SLOTS ? first + sel[p] : first + sel[p] - SB_FTS_SLOTS];
                const uint2 te = sl.tab[ly + RPP * p][lx];
                const uint2 bw = __ldg(a.bilin_lut + (te.y & 1023u));
                unsigned lo0, hi0, lo1, hi1;
                if ((rec.y >> 24) != (unsigned)SB_FTS_DIRECT) {
                    const uint32_t r0 = smem_u32(&sl.box[0]) + (te.x & 0x3ffffu);
                    const uint32_t r1 = (te.x & (1u << 27)) ? r0 : r0 + rec.z;
                    lds_6bytes(r0, lo0, hi0);
                    lds_6bytes(r1, lo1, hi1);
                    if (te.x & (1u << 26)) {
                        hi0 = lo0 >> 8; lo0 = (lo0 & 0x00ffffffu) | (lo0 << 24);
                        hi1 = lo1 >> 8; lo1 = (lo1 & 0x00ffffffu) | (lo1 << 24);
                    }
                } else {
                    const FeatherTmaCam &c = a.cam[rec.w & 15u];
                    const unsigned x0 = te.x & 0x1fffu, y0 = (te.x >> 13) & 0x1fffu;
                    const uint8_t *g0 = c.src + (size_t)(y0 * c.sstep);
                    const uint8_t *g1 = (te.x & (1u << 27)) ? g0 : g0 + c.sstep;
                    const unsigned x1 = x0 + 1u - ((te.x >> 26) & 1u);
                    load_tap_row(g0, x0, x1, lo0, hi0);
                    load_tap_row(g1, x0, x1, lo1, hi1);
                }
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
                unsigned char *orow = reinterpret_cast<unsigned char *>(a.out) + (size_t)Y0 * a.out_step + (size_t)X * (OUT8 ? 3 : 6);
                uint8_t *mrow_ = a.out_mask ? a.out_mask + (size_t)Y0 * a.mask_step + X : nullptr;
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    if (Y0 + RPP * p >= a.ph) break;
                    if (OUT8) {
                        orow[0] = (uint8_t)acc[p][0]; orow[1] = (uint8_t)acc[p][1]; orow[2] = (uint8_t)acc[p][2];
                    } else {
                        short *o = reinterpret_cast<short *>(orow);
                        o[0] = (short)acc[p][0]; o[1] = (short)acc[p][1]; o[2] = (short)acc[p][2];
                    }
                    if (mrow_) { *mrow_ = (uint8_t)mor[p]; mrow_ += RPP * a.mask_step; }
                    orow += RPP * a.out_step;
                }
            }
            tx += gx; ty += gy;
            if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
            if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
            continue;
        }
        for (int k = 0; k < nc; ++k) {                      // ascending camera index = feed order (float weight sums)
            const uint4 rec = sm.desc[stage][1 + k];
            const unsigned pitch = rec.z;
            const bool direct = (rec.y >> 24) == (unsigned)SB_FTS_DIRECT;    // block-uniform
            const FtsSlot &sl = sm.slot[first + k < SB_FTS_SLOTS ? first + k : first + k - SB_FTS_SLOTS];
            const uint32_t box = smem_u32(&sl.box[0]);
            uint2 te[PX], bw[PX];
            unsigned lo0[PX], hi0[PX], lo1[PX], hi1[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                te[p] = sl.tab[ly + RPP * p][lx];
                bw[p] = __ldg(a.bilin_lut + (te[p].y & 1023u));
                if (!direct) {
                    const uint32_t r0 = box + (te[p].x & 0x3ffffu);
                    const uint32_t r1 = (te[p].x & (1u << 27)) ? r0 : r0 + pitch;
                    lds_6bytes(r0, lo0[p], hi0[p]);
                    lds_6bytes(r1, lo1[p], hi1[p]);
                    if (te[p].x & (1u << 26)) {             // x1 == x0 at the image edge: repeat the pixel
                        hi0[p] = lo0[p] >> 8; lo0[p] = (lo0[p] & 0x00ffffffu) | (lo0[p] << 24);
                        hi1[p] = lo1[p] >> 8; lo1[p] = (lo1[p] & 0x00ffffffu) | (lo1[p] << 24);
                    }
                } else {                                    // box too large for the slot: gather from global memory
                    lo0[p] = hi0[p] = lo1[p] = hi1[p] = 0u;
                    if ((te[p].y >> 16) != 0u) {
                        const FeatherTmaCam &c = a.cam[rec.w & 15u];
                        const unsigned x0 = te[p].x & 0x1fffu, y0 = (te[p].x >> 13) & 0x1fffu;
                        const uint8_t *g0 = c.src + (size_t)(y0 * c.sstep);
                        const uint8_t *g1 = (te[p].x & (1u << 27)) ? g0 : g0 + c.sstep;
                        const unsigned x1 = x0 + 1u - ((te[p].x >> 26) & 1u);
                        load_tap_row(g0, x0, x1, lo0[p], hi0[p]);
                        load_tap_row(g1, x0, x1, lo1[p], hi1[p]);
                    }
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
            unsigned char *orow = reinterpret_cast<unsigned char *>(a.out) + (size_t)Y0 * a.out_step + (size_t)X * (OUT8 ? 3 : 6);
            uint8_t *mrow_ = a.out_mask ? a.out_mask + (size_t)Y0 * a.mask_step + X : nullptr;
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (Y0 + RPP * p >= a.ph) break;
                const int m = wsum[p] > SB_WEIGHT_EPS ? 255 : 0;
                const SharedDiv div(__fadd_rn(wsum[p], SB_WEIGHT_EPS));            // in [1e-5, n + 1e-5]: fast-path range
                const int o0 = m ? __float2int_rz(div((float)acc[p][0])) : 0, o1 = m ? __float2int_rz(div((float)acc[p][1])) : 0,
                          o2 = m ? __float2int_rz(div((float)acc[p][2])) : 0;
                if (OUT8) {
                    orow[0] = (uint8_t)o0; orow[1] = (uint8_t)o1; orow[2] = (uint8_t)o2;
                } else {
                    short *o = reinterpret_cast<short *>(orow);
                    o[0] = (short)o0; o[1] = (short)o1; o[2] = (short)o2;
                }
                if (mrow_) { *mrow_ = (uint8_t)m; mrow_ += RPP * a.mask_step; }
                orow += RPP * a.out_step;
            }
        }
        tx += gx; ty += gy;
        if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
        if (++stage == SB_FTT_STAGES) { stage = 0; parity ^= 1u; }
    }
}

int launch_feather_stream(const FeatherTmaArgs &a, bool apply_gain, bool out8, int sm_count, cudaStream_t s)
{
    SB_ASSERT(a.sharpness > 0.f && a.bilin_lut && a.desc && a.n_tiles > 0 && a.n <= 16);
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
    const int grid = std::min(a.n_tiles, SB_FTS_CTAS_PER_SM * sm_count);
    void *params[] = {const_cast<FeatherTmaArgs *>(&a)};
    SB_CUDA(cudaLaunchKernel(fn[v], dim3(grid), dim3(SB_FTS_THREADS), params, smem, s));
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
