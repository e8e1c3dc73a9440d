// kernels_fstream2.cu — the feather / no-blend frame kernel, second generation (design notes: sb_fs2.h).
//
// Reference semantics replaced by ONE launch per frame set:
//   RotationWarperBase::warp -> cv::remap INTER_LINEAR BORDER_REFLECT      warpers_inl.hpp:88-99
//   GainCompensator::apply / BlocksGainCompensator::apply                  exposure_compensate.cpp:150-153, 225-246
//   convertTo(CV_16S), FeatherBlender::feed x n, FeatherBlender::blend     blenders.cpp:136-155, 383-424
//   Blender::feed / Blender::blend (no blending)                           blenders.cpp:81-112
//   result.convertTo(CV_8U)                                                stitcher.cpp:313
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <vector>

#include "sb_device.cuh"
#include "sb_fs2.h"
#include "sb_tma.cuh"

namespace sb {
using namespace sbd;
using namespace sbt;

#define SB_WEIGHT_EPS 1e-5f

// ------------------------------------------------------------------------------------ setup kernels
// pass 1: per (camera, tile) the source bounding box of the weighted entries of the row-major feather table
//   table.x = x0 | y0 << 13 | (x1 == x0) << 26 | (y1 == y0) << 27 | unrepresentable << 28,  table.y = fx | fy << 5 | dist << 16
__global__ void __launch_bounds__(256)
k_fs2_bbox(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, float sharpness, Fs2Box *boxes)
{
    __shared__ int red[4][8];
    bool all_one = true;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
    for (int e = tid; e < FS2_W * FS2_H; e += blockDim.x) {
        const int lx = e % FS2_W, ly = e / FS2_W;
        const int x = (tx0 + (int)blockIdx.x) * FS2_W + lx - dx, y = (ty0 + (int)blockIdx.y) * FS2_H + ly - dy;
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
        Fs2Box b{};
        b.mnx = mnx; b.mny = mny; b.mxx = mxx; b.mxy = mxy; b.all_one = every;
        boxes[blockIdx.y * ntx + blockIdx.x] = b;
    }
}

// pass 2: the tile-major block of every (camera, tile): 1024 tap entries (sb_fs2.h) + 1024 weight-index bytes.
// An x-folded pair (x1 == x0 at the source edge) is stored with fx = 0, a y-folded one with fy = 0: the weights of the
// second column / row are then zero, sum w*p is the same integer, and the consumer has one code path (the bytes it
// multiplies by zero lie inside the staged box or right behind it in shared memory).
__global__ void __launch_bounds__(256)
k_fs2_entries(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, const Fs2Place *place,
              unsigned dist_cap, unsigned char *blocks)
{
    const Fs2Place pl = place[blockIdx.y * ntx + blockIdx.x];
    unsigned char *blk = blocks + ((size_t)blockIdx.y * ntx + blockIdx.x) * FS2_BLOCK_BYTES;
    uint32_t *ent = reinterpret_cast<uint32_t *>(blk);
    unsigned char *plane = blk + FS2_ENT_BYTES;
    for (int e = threadIdx.x; e < FS2_W * FS2_H; e += blockDim.x) {
        const int lx = e % FS2_W, ly = e / FS2_W;
        const int x = (tx0 + (int)blockIdx.x) * FS2_W + lx - dx, y = (ty0 + (int)blockIdx.y) * FS2_H + ly - dy;
        uint32_t t = 0u;
        unsigned d = 0u;
        if (pl.valid && (unsigned)x < (unsigned)ww && (unsigned)y < (unsigned)wh) {
            const uint2 s = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(table) + (size_t)y * tstep)[x];
            d = min(s.y >> 16, dist_cap);
            if (d != 0u) {
                const unsigned x0 = s.x & 0x1fffu, y0 = (s.x >> 13) & 0x1fffu;
                unsigned fxy = s.y & 1023u;
                if (s.x & (1u << 26)) fxy &= ~31u;
                if (s.x & (1u << 27)) fxy &= 31u;
                const unsigned off = (y0 - (unsigned)pl.ylo) * (unsigned)pl.pitch + x0 * 3u - (unsigned)pl.xlo;
                t = (fxy << 22) | ((off >> 2) << 5) | ((off & 3u) << 3);
            }
        }
        ent[e] = t;
        plane[e] = (unsigned char)d;
    }
}

int launch_fs2_bbox(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                    float sharpness, Fs2Box *boxes, cudaStream_t s)
{
    k_fs2_bbox<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, ww, wh, dx, dy, tx0, ty0, ntx, sharpness, boxes);
    SB_LAUNCHED();
    return SB_OK;
}
int launch_fs2_entries(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                       const Fs2Place *place, unsigned dist_cap, unsigned char *blocks, cudaStream_t s)
{
    k_fs2_entries<<<dim3(ntx, nty), 256, 0, s>>>(table, tstep, ww, wh, dx, dy, tx0, ty0, ntx, place, dist_cap, blocks);
    SB_LAUNCHED();
    return SB_OK;
}
int fs2_grid(int n_tiles, int sm_count) { return std::min(n_tiles, FS2_CTAS_PER_SM * sm_count); }

// ------------------------------------------------------------------------------------ host-side setup
// Per calibration: boxes of every (camera, tile) -> box width classes -> tile-major blocks -> per-tile descriptors in
// schedule order with the shared-memory ring plan.
//   descriptor (FS2_DESC_RECS x 16 bytes per panorama tile): everything the producer warp needs, ready to issue
//   [0]     = {n_cams | short path << 2 | edge tile << 3 | copies << 4 | ring units << 8, X0 | Y0 << 16, ring start unit,
//              lag behind the consumed tiles | (bytes the copies deliver >> 4) << 16}   (see the ring plan below)
//   [1 + k] = per camera slot (ascending camera index = feed order), for the consumers
//             {byte offset of the slot inside the tile's ring space, box pitch, -, camera}
//   [4 + j] = copy j of the tile: {destination byte offset inside the tile's ring space,
//             bulk copy: byte offset of the table block in the camera's blocks, bytes, camera
//             tensor copy: box x (byte) | box y << 16, tensor map index (camera * FS2_NCLS + class), bit 31 set}
int fs2_build(const Fs2CamSetup *cams, int n, int pw, int ph, float sharpness, int sm_count, bool gain_maps, DevBuf &desc_out, Fs2Plan *plan, cudaStream_t s,
              bool drop_empty)
{
    *plan = Fs2Plan{};
    const unsigned cls_w[FS2_NCLS] = {96u, 160u, 224u, 256u};      // pitches of 32 (mod 128) bytes keep the 4 tile rows of a warp on distinct banks
    for (int c = 0; c < FS2_NCLS; ++c) plan->cls_w[c] = cls_w[c];
    // block gain maps travel with the tile (one more tensor copy per camera slot) when the signed 16-bit tile coordinates reach
    bool gain_tma = gain_maps;
    for (int i = 0; i < n; ++i) gain_tma = gain_tma && cams[i].ww < 32768 && cams[i].wh < 32768 && pw < 32768 && ph < 32768;
    plan->gain_tma = gain_tma;
    // the weight index is one byte: the weight must saturate by distance 255 (Blender::NO: the mask byte itself)
    if (!(sharpness > 0.f) || fminf((float)(255.f * sharpness), 1.f) != 1.f) return SB_OK;
    const int tiles_x = div_up(pw, FS2_W), tiles_y = div_up(ph, FS2_H);
    int n_tiles = tiles_x * tiles_y;
    std::vector<std::vector<Fs2Box>> boxes(n);
    std::vector<std::vector<Fs2Place>> places(n);
    DevBuf tmp;
    for (int i = 0; i < n; ++i) {
        const Fs2CamSetup &c = cams[i];
        const size_t nt = (size_t)c.ntx * c.nty;
        boxes[i].resize(nt);
        SB_TRY(tmp.ensure(nt * sizeof(Fs2Box)));
        SB_TRY(launch_fs2_bbox(c.table, c.tstep, c.ww, c.wh, c.dx, c.dy, c.tx0, c.ty0, c.ntx, c.nty, sharpness, static_cast<Fs2Box *>(tmp.p), s));
        SB_CUDA(cudaMemcpyAsync(boxes[i].data(), tmp.p, nt * sizeof(Fs2Box), cudaMemcpyDeviceToHost, s));
        SB_CUDA(cudaStreamSynchronize(s));
        places[i].assign(nt, Fs2Place{0, 0, 0, 0});
        for (size_t t = 0; t < nt; ++t) {
            const Fs2Box &b = boxes[i][t];
            if (b.mxx < b.mnx) continue;
            const int xlo = (b.mnx * 3) & ~15, need_w = (b.mxx + 1) * 3 - xlo, rows = b.mxy - b.mny + 1;
            int cls = 0;
            while (cls < FS2_NCLS && (int)cls_w[cls] < need_w) ++cls;
            if (cls == FS2_NCLS || rows > FS2_ROWS * FS2_MAX_OPS) return SB_OK;       // a box the ring cannot stage: plan->ok stays false
            places[i][t] = Fs2Place{xlo, b.mny, (int)cls_w[cls], 1 | (cls << 4) | (div_up(rows, FS2_ROWS) << 8)};
        }
        SB_TRY(tmp.ensure(nt * sizeof(Fs2Place)));
        SB_CUDA(cudaMemcpyAsync(tmp.p, places[i].data(), nt * sizeof(Fs2Place), cudaMemcpyHostToDevice, s));
        SB_TRY(launch_fs2_entries(c.table, c.tstep, c.ww, c.wh, c.dx, c.dy, c.tx0, c.ty0, c.ntx, c.nty, static_cast<const Fs2Place *>(tmp.p),
                                  255u, c.blocks, s));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    const size_t per = FS2_DESC_RECS;
    std::vector<uint4> d((size_t)n_tiles * per, make_uint4(0u, 0u, 0u, 0u));
    double table_bytes = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        uint4 *o = &d[(size_t)t * per];
        int k = 0, one = 0;
        int slot_cam[FS2_MAXC];
        size_t slot_blk[FS2_MAXC];
        for (int i = 0; i < n; ++i) {
            const int rx = tx - cams[i].tx0, ry = ty - cams[i].ty0;
            if ((unsigned)rx >= (unsigned)cams[i].ntx || (unsigned)ry >= (unsigned)cams[i].nty) continue;
            const size_t blk = (size_t)ry * cams[i].ntx + rx;
            if (!(places[i][blk].valid & 1)) continue;
            if (k == FS2_MAXC) return SB_OK;                                          // more cameras than slots: plan->ok stays false
            slot_cam[k] = i; slot_blk[k] = blk; one = boxes[i][blk].all_one;
            ++k;
        }
        const unsigned short_path = (k == 1 && one) ? 1u : 0u;
        const unsigned blk_bytes = short_path ? FS2_ENT_BYTES : FS2_BLOCK_BYTES;
        unsigned units = 0, tx_bytes = 0, n_ops = 0;
        for (int j = 0; j < k; ++j) {
            const Fs2Place &pl = places[slot_cam[j]][slot_blk[j]];
            const unsigned cls = (unsigned)(pl.valid >> 4) & 15u, nops = (unsigned)pl.valid >> 8;
            const unsigned bytes = blk_bytes + nops * FS2_ROWS * (unsigned)pl.pitch, slot_off = units * 128u;
            if (slot_blk[j] * (size_t)FS2_BLOCK_BYTES >= (1ull << 32)) return SB_OK;
            const unsigned box_bytes = nops * FS2_ROWS * (unsigned)pl.pitch, gain_off = blk_bytes + ((box_bytes + 127u) & ~127u);
            o[1 + j] = make_uint4(slot_off, (unsigned)pl.pitch, gain_off, (unsigned)slot_cam[j]);
            // copy list: the table block (bulk copy), then the source box, FS2_ROWS rows per tensor copy
            o[1 + FS2_MAXC + n_ops++] = make_uint4(slot_off, (unsigned)(slot_blk[j] * FS2_BLOCK_BYTES), blk_bytes, (unsigned)slot_cam[j]);
            for (unsigned q = 0; q < nops; ++q)
                o[1 + FS2_MAXC + n_ops++] = make_uint4(slot_off + blk_bytes + q * FS2_ROWS * (unsigned)pl.pitch,
                                                       ((unsigned)pl.xlo & 0xffffu) | ((unsigned)(pl.ylo + (int)(q * FS2_ROWS)) << 16),
                                                       (unsigned)slot_cam[j] * FS2_NCLS + cls, 0x80000000u);
            unsigned slot_bytes = bytes;
            if (gain_tma) {     // the 32x32 floats of the camera's resized gain map under this tile (out-of-range: zero, never used)
                // (coordinates in the gain map PADDED by one tile all round, so that the 32x32 box always lies inside the tensor: a
                // tensor copy that overlaps its tensor by a single element in x and y raised an illegal-instruction fault on
                // B200 - scripts/microbench/tma_f32.cu, box at (-31, -31) - although partly outside boxes are fine in general)
                const int gx = tx * FS2_W - cams[slot_cam[j]].dx + FS2_W, gy = ty * FS2_H - cams[slot_cam[j]].dy + FS2_H;
                if (gx < 0 || gy < 0) return fail(SB_ERR_ASSERT, "fs2_build: gain tile of a camera that does not touch the tile");
                // the inner coordinate of a tensor copy must be 16-byte aligned (x = -31 floats faults, -8 and -700 do not): the box
                // is FS2_GAIN_W = 36 floats wide and starts at the multiple of 4 below gx; the consumers skip the first gx & 3 columns
                const int gxa = gx & ~3;
                o[1 + j].z |= (unsigned)(gx - gxa) << 28;
                o[1 + FS2_MAXC + n_ops++] = make_uint4(slot_off + gain_off, ((unsigned)gxa & 0xffffu) | ((unsigned)gy << 16), (unsigned)slot_cam[j], 0xc0000000u);
                slot_bytes = gain_off + FS2_GAIN_BYTES;
                tx_bytes += FS2_GAIN_BYTES;
            }
            units += (slot_bytes + 127u) >> 7;
            tx_bytes += bytes;
            table_bytes += blk_bytes;
        }
        const unsigned X0 = (unsigned)(tx * FS2_W), Y0 = (unsigned)(ty * FS2_H);
        const unsigned edge = ((int)X0 + FS2_W > pw || (int)Y0 + FS2_H > ph) ? 1u : 0u;
        o[0] = make_uint4((unsigned)k | (short_path << 2) | (edge << 3) | (n_ops << 4) | (units << 8), X0 | (Y0 << 16), 0u, (tx_bytes >> 4) << 16);
    }
    // Schedule order: CTA b takes positions b, b + G, ... of the descriptor array, so listing the tiles by descending cost
    // (blended tiles with 3, 2, 1 cameras, then the single-camera short-path tiles, then empty ones; row-major inside a
    // class) deals every CTA the same number of tiles of each class, the expensive ones first ...
    std::vector<int> order(n_tiles);
    for (int t = 0; t < n_tiles; ++t) order[t] = t;
    auto cost = [&](int t) {
        const unsigned x = d[(size_t)t * per].x;
        const int nc = (int)(x & 3u);
        return nc == 0 ? 0 : (x & 4u) ? 1 : 1 + nc;
    };
    std::stable_sort(order.begin(), order.end(), [&](int l, int r) { return cost(l) > cost(r); });
    if (drop_empty) {            // output planes: tiles no camera feeds are not part of any plane - they leave the schedule
        while (!order.empty() && cost(order.back()) == 0) order.pop_back();
        if (order.empty()) return SB_OK;
    }
    const int n_all = n_tiles;
    n_tiles = (int)order.size();
    (void)n_all;
    const int grid = fs2_grid(n_tiles, sm_count);
    // ... except that every CTA's FIRST tiles are cheap ones: the first wave of copies is what the consumers wait for at
    // kernel start, and a one-camera tile is 2-3x fewer bytes
    static const int sched_mode = getenv("SB_FS2_SCHED") ? atoi(getenv("SB_FS2_SCHED")) : 1;      // 0: the descending-cost order of the first version
    if (sched_mode == 1 && n_tiles >= 2 * grid * FS2_GROUPS) {
        // interleaved: rounds (one tile per CTA) of blended tiles spread evenly among the rounds of one-camera tiles, so that the
        // bytes a CTA's ring must hold ahead of its consumers are the average of the two kinds instead of the blended maximum
        // (measured: app6 +2 %, C2 +0.4 % over the descending-cost order; the first FS2_GROUPS rounds stay cheap ones)
        std::vector<int> ex, ch;
        for (int t : order) (cost(t) >= 2 ? ex : ch).push_back(t);
        const int ne = div_up((int)ex.size(), grid), nc = div_up((int)ch.size(), grid);
        // Round p of the CTA's sequence is consumed by group p % FS2_GROUPS, so an expensive round goes where the group has had
        // the fewest so far (a plain 1-in-3 interleave handed every blended tile to the same group: its tiles then retire
        // late, ring space is handed back in order, and the producers stall - seen in the pipeline trace), paced so that
        // expensive rounds stay evenly spread; the first FS2_GROUPS rounds are cheap ones.
        const int rounds = ne + nc;
        std::vector<int> merged;
        int used_e = 0, used_c = 0, per_group[FS2_GROUPS] = {};
        for (int r = 0; r < rounds; ++r) {
            const int g = r % FS2_GROUPS;
            int least = per_group[0];
            for (int q = 1; q < FS2_GROUPS; ++q) least = std::min(least, per_group[q]);
            const bool behind = (long long)used_e * rounds < (long long)r * ne;         // fewer expensive rounds placed than an even spread would have
            bool take_e = used_e < ne && r >= std::min(FS2_GROUPS, nc) && behind && per_group[g] <= least;
            if (used_c >= nc) take_e = true;
            const std::vector<int> &src = take_e ? ex : ch;
            const size_t c0 = (size_t)(take_e ? used_e : used_c) * grid;
            for (size_t k = c0; k < std::min(src.size(), c0 + (size_t)grid); ++k) merged.push_back(src[k]);
            per_group[g] += cost(src[c0]);                  // a group's load in units of a one-camera tile (2 / 3 / 4: blended with 1 / 2 / 3 cameras)
            if (take_e) ++used_e; else ++used_c;
        }
        order.swap(merged);
    } else
    if (n_tiles >= 2 * grid * FS2_GROUPS) std::rotate(order.begin(), order.end() - (size_t)grid * FS2_GROUPS, order.end());
    // position t of the schedule belongs to CTA t % grid as its tile number t / grid; the CTA's descriptors are contiguous
    const int per_cta = div_up(n_tiles, grid);
    std::vector<uint4> o((size_t)grid * per_cta * per, make_uint4(0u, 0u, 0u, 0u));
    auto at = [&](int t) -> uint4 * { return &o[((size_t)(t % grid) * per_cta + (size_t)(t / grid)) * per]; };
    for (int t = 0; t < n_tiles; ++t)
        for (size_t j = 0; j < per; ++j) at(t)[j] = d[(size_t)order[t] * per + j];
    // The shared-memory ring plan.  CTA b walks positions b, b + G, ...; a tile takes `units` contiguous 128-byte ring units
    // at the head, wrapping to 0 when the end of the ring is too short; space is handed back in tile order.  Both are a pure
    // function of the tile sizes, so the start unit of every tile and how far behind it the CTA's consumed tiles must be
    // before its copies may land are computed here, once.  A multi-frame launch walks the same tile sequence frame after
    // frame without emptying the ring, so the plan is made for three frames in a row: frame 0 starts from an empty ring
    // (plan 0), and if frames 1 and 2 come out the same the ring has reached its cycle (plan 1, for every later frame).
    //   [0].z = start unit (plan 0) | start unit (plan 1) << 16
    //   [0].w = lag (plan 0) | lag (plan 1) << 8 | bytes >> 4 << 16;  lag: copies of tile number s (counted over the frames)
    //           may land once the CTA's tiles up to number s - lag - 1 are consumed (also covers the reuse of the stage entry)
    // The producer warps then need no bookkeeping at all.
    constexpr int RING_UNITS = FS2_RING_BYTES / 128;
    bool steady = true;
    for (int b = 0; b < grid; ++b) {
        const int mine = (n_tiles - b + grid - 1) / grid;
        std::vector<int> start(3 * (size_t)mine), lag(3 * (size_t)mine);
        int hist[FS2_STAGES];
        int head = 0, tail = 0, oldest = 0;
        for (int g = 0; g < 3 * mine; ++g) {
            const int units = (int)(at(b + (g % mine) * grid)->x >> 8);
            int st;
            for (;;) {
                if (g - oldest < FS2_STAGES) {
                    if (g == oldest) head = tail = 0;       // nothing in flight
                    if (head >= tail) {                     // in use: [tail, head)
                        if (head + units <= RING_UNITS) { st = head; break; }
                        if (units < tail) { st = 0; break; }
                    } else if (head + units < tail) {       // in use: [tail, end) and [0, head)
                        st = head; break;
                    }
                }
                ++oldest;
                tail = oldest < g ? hist[oldest % FS2_STAGES] : head;
            }
            head = st + units;
            hist[g % FS2_STAGES] = st;
            start[g] = st; lag[g] = g - oldest;
        }
        for (int q = 0; q < mine; ++q) {
            uint4 &d0 = *at(b + q * grid);
            steady = steady && start[mine + q] == start[2 * mine + q] && lag[mine + q] == lag[2 * mine + q];
            d0.z = (unsigned)start[q] | ((unsigned)start[mine + q] << 16);
            d0.w |= (unsigned)lag[q] | ((unsigned)lag[mine + q] << 8);
        }
    }
    plan->steady = steady;
    SB_TRY(desc_out.ensure(o.size() * sizeof(uint4)));
    SB_CUDA(cudaMemcpyAsync(desc_out.p, o.data(), o.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaStreamSynchronize(s));
    plan->n_tiles = n_tiles; plan->grid = grid; plan->per_cta = per_cta; plan->table_bytes = table_bytes; plan->ok = true;
    return SB_OK;
}

// ------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_tmap_encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmap_encode tmap_encoder()
{
    static PFN_tmap_encode fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmap_encode>(p);
    }
    return fn;
}
// The 8UC3 source image as a 2-D tensor of bytes {row bytes, rows}; a copy brings cls_w[c] x FS2_ROWS bytes densely
// packed into shared memory, out-of-range bytes read as zero (they are only ever multiplied by zero weights).
int fs2_encode_tmaps(const void *src, size_t sstep, int sw, int sh, const unsigned cls_w[FS2_NCLS], CUtensorMap out[FS2_NCLS])
{
    PFN_tmap_encode enc = tmap_encoder();
    if (!enc) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)sw * 3u, (cuuint64_t)sh};
    const cuuint64_t gstride[1] = {(cuuint64_t)sstep};
    const cuuint32_t estr[2] = {1u, 1u};
    for (int c = 0; c < FS2_NCLS; ++c) {
        const cuuint32_t box[2] = {cls_w[c], (cuuint32_t)FS2_ROWS};
        const CUresult r = enc(&out[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2u, const_cast<void *>(src), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d source, pitch %zu, box %u", (int)r, sw, sh, sstep, cls_w[c]);
    }
    return SB_OK;
}

// an image of 32-bit pixels (RGBX Gaussian level) as a 2-D tensor, box = box_w x box_h pixels
int tmap_encode_u32(const void *base, size_t step, int w, int h, int box_w, int box_h, CUtensorMap *out)
{
    PFN_tmap_encode enc = tmap_encoder();
    if (!enc) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t gstride[1] = {(cuuint64_t)step};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1u, 1u};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2u, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d image of 32-bit pixels, pitch %zu", (int)r, w, h, step);
    return SB_OK;
}

// a camera's resized block gain map (CV_32FC1 of the warped size + one tile of zero padding on the top and left) as a 2-D float tensor, box = one tile
int fs2_encode_gain_tmap(const float *gmap, size_t step, int w, int h, CUtensorMap *out)
{
    PFN_tmap_encode enc = tmap_encoder();
    if (!enc) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t gstride[1] = {(cuuint64_t)step};
    const cuuint32_t box[2] = {(cuuint32_t)FS2_GAIN_W, (cuuint32_t)FS2_H}, estr[2] = {1u, 1u};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2u, const_cast<float *>(gmap), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d gain map, pitch %zu", (int)r, w, h, step);
    return SB_OK;
}

// ------------------------------------------------------------------------------------ frame kernel
struct Fs2Smem {
    unsigned char ring[FS2_RING_BYTES];
    uint4 desc[FS2_STAGES][1 + FS2_MAXC];                   // stage copy of the tile descriptor, rewritten by the producer:
                                                            //   [0] = {flags, X0 | Y0 << 16, output byte offset, mask byte offset}
                                                            //   [1 + k] = {slot offset in the ring, box pitch, -, camera}
    uint64_t full[FS2_STAGES], empty[FS2_STAGES];
};
static_assert(sizeof(Fs2Smem) * FS2_CTAS_PER_SM + 1024 * FS2_CTAS_PER_SM <= 232448, "shared memory of one SM");

// Bilinear product weights of a table entry (fx = bits 22-26, fy = bits 27-31), times 64, as the two operands of the
// 16-bit x 8-bit dot products: wx = w00 | w01 << 16, wy = w10 | w11 << 16 with w = a*b*64, a in {32 - fx, fx},
// b in {32 - fy, fy} (sb_device.cuh: bilin_weights).  The factor 64 puts the result byte of (sum ab*p + 512) >> 10 into
// byte 2 of the 32-bit sum.  They are computed, not looked up: a 1024-entry table in shared memory costs 3.4 bank
// conflicts per load (the index is as good as random across a warp), a third of the kernel's shared-memory traffic.
//   t = (32 - fx) * 64 | fx * 64 << 16 = fx * (0xffff * 64) + 2048;   wx = t * (32 - fy);   wy = t * fy
// (fx, fy) = (0, 0) would need w00 = 2^16: it becomes 0xffff, which gives the same byte for every 8-bit p, with and
// without the "- 1" of the full-weight path:  (65535 p + 32768) >> 16 == p  and  (65535 p - 32768) >> 16 == p - 1 for
// p in 1..255.
__device__ __forceinline__ void fs2_weights(uint32_t e, unsigned &wx, unsigned &wy)
{
    const unsigned fy = e >> 27, fx = (e >> 22) & 31u;
    const unsigned t = fx * (0xffffu * 64u) + 2048u;
    wx = t * (32u - fy);
    wy = t * fy;
    if (wx == 0x10000u) wx = 0xffffu;
}

// One output pixel of one camera: three 32-bit sums s_c = 64 * (sum ab*p_c) + BIAS; the pixel value is byte 2.
// e = table entry (sb_fs2.h), box = byte offset of the staged source box in shared memory, pitch = its row pitch.
template <int BIAS>
__device__ __forceinline__ void fs2_pixel(const unsigned char *sm, uint32_t box, uint32_t pitch, uint32_t e, unsigned &s0, unsigned &s1, unsigned &s2)
{
#ifdef SB_FS2_EXPERIMENT_NOCONFLICT
    e &= ~0x7ffe0u;                                          // tuning experiment: every tap at box offset 0 (broadcast, no bank conflicts; results are garbage)
#endif
    uint2 bw;
    fs2_weights(e, bw.x, bw.y);
    const unsigned char *r0 = sm + box + ((e >> 3) & 0xfffcu);
    const unsigned char *r1 = r0 + pitch;
    const unsigned a0 = *reinterpret_cast<const uint32_t *>(r0), a1 = *reinterpret_cast<const uint32_t *>(r0 + 4),
                   a2 = *reinterpret_cast<const uint32_t *>(r0 + 8);
    const unsigned b0 = *reinterpret_cast<const uint32_t *>(r1), b1 = *reinterpret_cast<const uint32_t *>(r1 + 4),
                   b2 = *reinterpret_cast<const uint32_t *>(r1 + 8);
    const unsigned lo0 = __funnelshift_r(a0, a1, e), hi0 = __funnelshift_r(a1, a2, e);     // (shift amount = e mod 32 = 8 * byte shift)
    const unsigned lo1 = __funnelshift_r(b0, b1, e), hi1 = __funnelshift_r(b1, b2, e);
    const unsigned m0 = __byte_perm(lo0, hi0, 0x5241), m1 = __byte_perm(lo1, hi1, 0x5241);     // [G0 G1 B0 B1] per row
    const unsigned p0 = __byte_perm(lo0, lo1, 0x7430), p1 = __byte_perm(m0, m1, 0x5410), p2 = __byte_perm(m0, m1, 0x7632);
    s0 = __dp2a_hi(bw.y, p0, __dp2a_lo(bw.x, p0, (unsigned)BIAS));
    s1 = __dp2a_hi(bw.y, p1, __dp2a_lo(bw.x, p1, (unsigned)BIAS));
    s2 = __dp2a_hi(bw.y, p2, __dp2a_lo(bw.x, p2, (unsigned)BIAS));
}

// exposure gain of camera `c` at panorama pixel (X, Y): the scalar of GainCompensator or the resized block map
__device__ __forceinline__ float fs2_gain(const Fs2Cam &c, int X, int Y)
{
    return c.gmap ? __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(c.gmap) + (size_t)(Y - c.dy) * c.gmstep) + (X - c.dx)) : c.gain;
}
// ... of the thread's 4 pixels: from the gain tile staged with the tile (rec = the camera slot's stage record), else as above
__device__ __forceinline__ void fs2_gain4(const Fs2Args &a, const unsigned char *smb, const uint4 &rec, uint32_t gain_idx, int X, int Y, int nx, float (&g)[4])
{
    const Fs2Cam &c = a.cam[rec.w & 15u];
    if (a.gain_tma) {                                       // rec.z = byte offset of the gain tile in the slot | columns to skip << 28
        const float *t = reinterpret_cast<const float *>(smb + rec.x + (rec.z & 0x0fffffffu) + gain_idx) + (rec.z >> 28);
        g[0] = t[0]; g[1] = t[1]; g[2] = t[2]; g[3] = t[3];
        return;
    }
    if (!c.gmap) { g[0] = g[1] = g[2] = g[3] = c.gain; return; }      // GainCompensator: one scalar per camera
#pragma unroll
    for (int p = 0; p < 4; ++p) g[p] = p >= nx ? 1.f : fs2_gain(c, X + p, Y);
}

// saturate_cast<uchar>(v * gain) for an 8-bit v: cvRound then clamp == clamp then round-half-even (rounding is monotonic), and
// both conversions ride the FP32 pipe through the 2^23 constant instead of I2F / F2I (a quarter-rate unit that the 12 gain
// products of a thread's 4 pixels would otherwise saturate): float(v) = bits(2^23 | v) - 2^23, rint(x) = low byte of bits(x + 2^23).
// max before min sends a NaN product to 0 like cvRound's INT_MIN.
__device__ __forceinline__ unsigned fs2_apply_gain(unsigned v, float g)
{
    const float f = __fsub_rn(__uint_as_float(v | 0x4b000000u), 8388608.f);
    const float x = fminf(fmaxf(__fmul_rn(f, g), 0.f), 255.f);
    return __float_as_uint(__fadd_rn(x, 8388608.f)) & 0xffu;
}
// the same for a value sitting in byte 2 of a dot-product sum (fs2_pixel), result in byte 0
__device__ __forceinline__ unsigned fs2_apply_gain_b2(unsigned s, float g)
{
    const float f = __fsub_rn(__uint_as_float(__byte_perm(s, 0x4b000000u, 0x7442)), 8388608.f);
    const float x = fminf(fmaxf(__fmul_rn(f, g), 0.f), 255.f);
    return __float_as_uint(__fadd_rn(x, 8388608.f)) & 0xffu;
}

// crop_app_fill: camera 0's warped pixel (0, 0), sampled from its frame in global memory (rare: only pixels no camera covers)
template <bool GAIN>
__device__ __noinline__ uint3 fs2_fill_pixel(const Fs2Args &a, const uint8_t *src0, unsigned sstep0)      // (by value: references would pin the caller's pixel registers to the stack)
{
    const unsigned tex = a.fill_tex[0], fxy = a.fill_tex[1] & 1023u;
    const unsigned x0 = tex & 0x1fffu, y0 = (tex >> 13) & 0x1fffu;
    const uint8_t *g0 = src0 + (size_t)y0 * sstep0;
    const uint8_t *g1 = (tex & (1u << 27)) ? g0 : g0 + sstep0;
    const unsigned x1 = x0 + 1u - ((tex >> 26) & 1u);
    unsigned lo0, hi0, lo1, hi1;
    load_tap_row(g0, x0, x1, lo0, hi0);
    load_tap_row(g1, x0, x1, lo1, hi1);
    int v0, v1, v2;
    bilinear_rgb(lo0, hi0, lo1, hi1, bilin_weights((int)(fxy & 31u), (int)(fxy >> 5)), v0, v1, v2);
    if (GAIN) {
        const float g = fs2_gain(a.cam[0], a.cam[0].dx, a.cam[0].dy);
        v0 = (int)fs2_apply_gain((unsigned)v0, g); v1 = (int)fs2_apply_gain((unsigned)v1, g); v2 = (int)fs2_apply_gain((unsigned)v2, g);
    }
    return make_uint3((unsigned)v0 << 16, (unsigned)v1 << 16, (unsigned)v2 << 16);
}

// The 4 pixels x 3 channels of one thread, each value in byte B of its register (the byte above it zero), as the 12 (8UC3)
// or 24 (16SC3) output bytes.  Full tiles: three 32-bit (64-bit) stores per thread; edge tiles: per pixel, guarded.
// (Tried in round 2: staging the warp's 4 rows in shared memory so that consecutive lanes store consecutive words made the
// C2 frame 1.7 us SLOWER - the extra shared-memory traffic costs more than the partial-sector stores.)
template <bool OUT8, int B>
__device__ __forceinline__ void fs2_store_quad(unsigned char *o, const unsigned (&v)[4][3], bool edge, int nx, bool row_ok)
{
#ifdef SB_FS2_EXPERIMENT_NOSTORE
    if (nx >= 0) return;                                     // tuning experiment: no output stores at all
#endif
    if (!edge) {
        if (OUT8) {
            constexpr unsigned S = (unsigned)B | ((unsigned)(B + 4) << 4);
            uint32_t *q = reinterpret_cast<uint32_t *>(o);
            q[0] = __byte_perm(__byte_perm(v[0][0], v[0][1], S), __byte_perm(v[0][2], v[1][0], S), 0x5410);
            q[1] = __byte_perm(__byte_perm(v[1][1], v[1][2], S), __byte_perm(v[2][0], v[2][1], S), 0x5410);
            q[2] = __byte_perm(__byte_perm(v[2][2], v[3][0], S), __byte_perm(v[3][1], v[3][2], S), 0x5410);
        } else {
            constexpr unsigned S = (unsigned)B | ((unsigned)(B + 1) << 4) | ((unsigned)(B + 4) << 8) | ((unsigned)(B + 5) << 12);
            uint2 *q = reinterpret_cast<uint2 *>(o);
            q[0] = make_uint2(__byte_perm(v[0][0], v[0][1], S), __byte_perm(v[0][2], v[1][0], S));
            q[1] = make_uint2(__byte_perm(v[1][1], v[1][2], S), __byte_perm(v[2][0], v[2][1], S));
            q[2] = make_uint2(__byte_perm(v[2][2], v[3][0], S), __byte_perm(v[3][1], v[3][2], S));
        }
        return;
    }
    if (!row_ok) return;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (p >= nx) break;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned val = (v[p][k] >> (8 * B)) & 0xffu;
            if (OUT8) o[p * 3 + k] = (unsigned char)val;
            else reinterpret_cast<short *>(o)[p * 3 + k] = (short)val;
        }
    }
}
// ... as 4 RGBX words (multi-band Gaussian level 0): one 16-byte store per thread
template <int B>
__device__ __forceinline__ void fs2_store_rgbx(unsigned char *o, const unsigned (&v)[4][3], bool edge, int nx, bool row_ok)
{
    constexpr unsigned S = (unsigned)B | ((unsigned)(B + 4) << 4);           // [v0.B, v1.B, -, -]
    constexpr unsigned T = 0x0010u | ((unsigned)(B + 4) << 8) | ((unsigned)(4 + ((B + 1) & 3)) << 12);   // [lo.0, lo.1, v2.B, 0 (a masked-off byte of v2)]
    uint4 w;
    w.x = __byte_perm(__byte_perm(v[0][0], v[0][1], S), v[0][2] & (0xffu << (8 * B)), T);
    w.y = __byte_perm(__byte_perm(v[1][0], v[1][1], S), v[1][2] & (0xffu << (8 * B)), T);
    w.z = __byte_perm(__byte_perm(v[2][0], v[2][1], S), v[2][2] & (0xffu << (8 * B)), T);
    w.w = __byte_perm(__byte_perm(v[3][0], v[3][1], S), v[3][2] & (0xffu << (8 * B)), T);
    if (!edge) { *reinterpret_cast<uint4 *>(o) = w; return; }
    if (!row_ok) return;
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (p < nx) reinterpret_cast<uint32_t *>(o)[p] = ws[p];
}
__device__ __forceinline__ void fs2_store_mask(uint8_t *m, unsigned word, bool edge, int nx, bool row_ok)
{
#ifdef SB_FS2_EXPERIMENT_NOSTORE
    if (nx >= 0) return;
#endif
    if (!edge) { *reinterpret_cast<uint32_t *>(m) = word; return; }
    if (!row_ok) return;
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (p < nx) m[p] = (uint8_t)(word >> (8 * p));
}

// Pipeline trace (debugging aid; builds with -DSB_FS2_TRACE only - scripts/build_variant.sh trace "-DSB_FS2_TRACE" - and then
// active when the environment variable SB_FS2_TRACE names a file): per CTA and tile the timestamps copies issued / consumer
// group starts waiting / data landed / tile finished / producer reaches the tile / producer done with it.
__device__ __forceinline__ unsigned long long fs2_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef SB_FS2_TRACE
#define FS2_TRACE(SLOT, SEQ)                                                                        \
    do {                                                                                            \
        if (a.trace && (SEQ) < FS2_TRACE_TILES) a.trace[((size_t)blockIdx.x * FS2_TRACE_TILES + (SEQ)) * 8 + (SLOT)] = fs2_now(); \
    } while (0)
#define FS2_DEBUG_SUPPLY_ONLY (a.debug == 1)
#else
#define FS2_TRACE(SLOT, SEQ) do { } while (0)
#define FS2_DEBUG_SUPPLY_ONLY false
#endif
constexpr int FS2_TRACE_TILES = 64;

// NOBLEND: Blender::feed / blend without blending (blenders.cpp:81-112): the pixel of the LAST camera (feed order)
// whose mask is non-zero, dst_mask = OR of the mask bytes, 0 where no camera has a mask.
// OUT: 0 = CV_16SC3 panorama, 1 = CV_8UC3 panorama, 2 (with NOBLEND) = the multi-band path's first stage: every "camera" is
// its own output plane (Gaussian level 0 of the camera's padded feed rect as RGBX words, Fs2Cam::mb_out), the "panorama" only
// the coordinate system the planes are stacked in (capi_compositor.cu: mb_fs2_setup).
template <bool GAIN, int OUT, bool NOBLEND>
__global__ void __launch_bounds__(FS2_THREADS, FS2_CTAS_PER_SM)
k_fs2(const __grid_constant__ Fs2Args a)
{
    constexpr bool OUT8 = OUT == 1;
    constexpr bool PLANES = OUT == 2;
    constexpr unsigned PX = OUT == 0 ? 6u : OUT == 1 ? 3u : 4u;
    static_assert(!PLANES || NOBLEND, "per-camera output planes exist only without blending");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Fs2Smem &sm = *reinterpret_cast<Fs2Smem *>(smem_raw);
    const unsigned char *const smb = smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    pdl_release();                                          // programmatic dependent launch: the next kernel of the stream may be set up from here on
    if (tid == 0) {
        for (int s = 0; s < FS2_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);                      // the producer's arrive (+ the copies' expect_tx)
            mbar_init(&sm.empty[s], FS2_GROUP_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= FS2_GROUPS * FS2_GROUP_WARPS) {
        // ------------------------------------------------ producer warps
        // The CTA walks its tiles (schedule positions blockIdx.x, + G, + 2G ...; their descriptors are contiguous) in order;
        // producer warp w fetches tiles w, w + P, ...  Where a tile lands in the ring and how many of the CTA's tiles must
        // have been consumed before its copies may land come from the per-calibration ring plan in the descriptor.
        // Lane l issues copy l of the tile: per camera one bulk copy (table block) and one tensor copy per FS2_ROWS box
        // rows.  A TMA instruction is warp-uniform, so the lanes' copies leave one after the other (~70 cycles each):
        // the parallelism that keeps the ring full comes from the P warps.
        const int pw = warp - FS2_GROUPS * FS2_GROUP_WARPS;
        const uint32_t ring0 = smem_u32(&sm.ring[0]);
        const uint4 *const dp0 = a.desc + (size_t)blockIdx.x * a.per_cta * FS2_DESC_RECS;
        const int n_mine = (a.n_tiles - (int)blockIdx.x + G - 1) / G;
        const int total = n_mine * max(a.n_frames, 1);      // a multi-frame launch walks the same tiles once per frame set
        // lane l holds record l of the tile's descriptor (one coalesced 256-byte load, a tile ahead of its use)
        auto load_rec = [&](int g) { return lane < FS2_DESC_RECS ? __ldg(dp0 + (size_t)(g % n_mine) * FS2_DESC_RECS + lane) : make_uint4(0u, 0u, 0u, 0u); };
        uint4 nx = make_uint4(0u, 0u, 0u, 0u);
        if (pw < total) nx = load_rec(pw);
        pdl_wait();                                         // the predecessor grid may still be running up to here (the descriptors are per-calibration data)
        int retired = 0;
        for (int gs = pw; gs < total; gs += FS2_PRODUCERS) {     // gs: tile number counted over the frames
            const uint4 d = nx;
            if (gs + FS2_PRODUCERS < total) nx = load_rec(gs + FS2_PRODUCERS);
            const int f = gs / n_mine, seq = gs - f * n_mine;
            const int stage = gs % FS2_STAGES;
            if (lane == 0) FS2_TRACE(4, gs);
            const unsigned d0x = __shfl_sync(0xffffffffu, d.x, 0), d0z = __shfl_sync(0xffffffffu, d.z, 0), d0w = __shfl_sync(0xffffffffu, d.w, 0);
            // plan 0: the ring was empty when this frame started (first frame; every frame when the plan is not cyclic and
            // the ring is drained in between); plan 1: the cyclic plan of the frames after the first
            const bool cyc = f > 0 && a.steady;
            int need = gs - (int)((cyc ? d0w >> 8 : d0w) & 0xffu);
            if (f > 0 && !a.steady) need = max(need, f * n_mine);
            for (; retired < need; ++retired)              // groups finish tiles out of order: every warp awaits every tile once, in order
                mbar_wait_hw(&sm.empty[retired % FS2_STAGES], (unsigned)(retired / FS2_STAGES) & 1u);
            const Fs2Frame *fr = a.frames ? a.frames + f : nullptr;
            const CUtensorMap *tmaps = fr ? fr->tmap : a.tmap;
            if (fr && seq < FS2_PRODUCERS && lane < a.n * FS2_NCLS)      // first use of this frame's tensor maps by this warp: they were
                asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmaps + lane) : "memory");   // written through the generic proxy
            const uint32_t tile_off = ((cyc ? d0z >> 16 : d0z) & 0xffffu) * 128u;          // ring byte offset of the tile
            if (lane <= FS2_MAXC) {
                uint4 ds = d;
                if (lane == 0) {                            // tile origin -> byte offsets of its first pixel in the panorama and the mask
                    const unsigned X0 = d.y & 0xffffu, Y0 = d.y >> 16;
                    ds.z = Y0 * a.out_step + X0 * PX; ds.w = Y0 * a.mask_step + X0;
                } else {
                    ds.x += tile_off;                       // the slot's byte offset in the ring
                }
                sm.desc[stage][lane] = ds;
            }
            __syncwarp();
            const int n_ops = (int)((d0x >> 4) & 15u);
            if (lane == 0) {
                if (n_ops == 0) mbar_arrive(&sm.full[stage]);
                else mbar_arrive_expect_tx(&sm.full[stage], (d0w >> 16) << 4);
                FS2_TRACE(0, gs);
            }
            __syncwarp();
            // copies: lane 1 + FS2_MAXC + j issues copy j of the tile's list
            if (lane > FS2_MAXC && lane <= FS2_MAXC + n_ops) {
                const uint32_t dst = ring0 + tile_off + d.x;
                if (d.w & 0x80000000u) {                    // tensor copy: a source box, or (bit 30) a tile of the camera's gain map
                    // (ONE copy instruction for both kinds: ptxas turned a second call site into a predicated UTMALDG that
                    // faulted as an illegal instruction on B200)
                    const CUtensorMap *m = (d.w & 0x40000000u) ? a.gtmap + d.z : tmaps + d.z;
                    tma_load_2d(dst, m, (int)(d.y & 0xffffu), (int)(d.y >> 16), &sm.full[stage]);
                }
                else bulk_g2s_addr(dst, a.cam[d.w].blocks + d.y, d.z, &sm.full[stage]);
            }
            if (lane == 0) FS2_TRACE(5, gs);
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    // group g = warp / 8 works on tiles seq = g, g + GROUPS, ... of the CTA's sequence; inside a group warp w covers tile
    // rows 4w .. 4w+3, lane = (row << 3 | quad), a thread owns pixels 4*quad .. 4*quad+3 of its row
    const int grp = warp / FS2_GROUP_WARPS, slab = warp % FS2_GROUP_WARPS;
    const int ty = slab * 4 + (lane >> 3), tx = (lane & 7) * 4;
    const uint32_t gain_idx = (uint32_t)(ty * FS2_GAIN_W + tx) * 4u;     // this thread's 4 gains inside a staged gain tile
    const uint32_t tab_off = (uint32_t)(ty * FS2_W + tx) * 4u, plane_off = (uint32_t)FS2_ENT_BYTES + (uint32_t)(ty * FS2_W + tx);
    const unsigned out_t = (unsigned)ty * a.out_step + (unsigned)tx * PX, mask_t = (unsigned)ty * a.mask_step + (unsigned)tx;
    const int n_mine = (a.n_tiles - (int)blockIdx.x + G - 1) / G;
    const int total = n_mine * max(a.n_frames, 1);
    unsigned char *out_f = reinterpret_cast<unsigned char *>(a.out);
    uint8_t *mask_f = a.out_mask;
    const uint8_t *src0 = a.src0;
    unsigned sstep0 = a.sstep0;
    int next_frame_at = a.frames ? 0 : total;               // tile number at which the next frame set's output pointers are due
    pdl_wait();                                             // (before the first output store, and before shared memory a copy of this grid lands in is read)
    for (int seq = grp; seq < total; seq += FS2_GROUPS) {   // seq: tile number counted over the frames
        if (seq >= next_frame_at) {
            const Fs2Frame *fr = a.frames + seq / n_mine;
            out_f = reinterpret_cast<unsigned char *>(fr->out); mask_f = fr->out_mask;
            src0 = fr->src0; sstep0 = fr->sstep0;
            next_frame_at = (seq / n_mine + 1) * n_mine;
        }
        const int stage = seq % FS2_STAGES;
        if (tid % (FS2_GROUP_WARPS * 32) == 0) FS2_TRACE(1, seq);
        mbar_wait_hw(&sm.full[stage], (unsigned)(seq / FS2_STAGES) & 1u);
        if (tid % (FS2_GROUP_WARPS * 32) == 0) FS2_TRACE(2, seq);
        if (FS2_DEBUG_SUPPLY_ONLY) {                         // tuning experiment (trace builds): the supply side alone (results are garbage)
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            continue;
        }
        const uint4 d0 = sm.desc[stage][0];
        const int nc = (int)(d0.x & 3u);
        const bool edge = (d0.x & 8u) != 0u;                // block-uniform: the tile crosses the panorama's right / bottom edge
        const int X = (int)(d0.y & 0xffffu) + tx, Y = (int)(d0.y >> 16) + ty;
        int nx = edge ? min(max(a.pw - X, 0), 4) : 4;
        bool row_ok = !edge || Y < a.ph;
        unsigned char *o = out_f + (out_t + d0.z);
        uint8_t *const mo = mask_f ? mask_f + (mask_t + d0.w) : nullptr;
        bool plane_edge = false;
        if (PLANES) {                                        // the tile's only camera = the output plane; clip to the plane's own size
            const Fs2Cam &pc = a.cam[sm.desc[stage][1].w & 15u];
            const int y = Y - pc.dy;
            o = pc.mb_out + (size_t)y * pc.mb_step + (size_t)X * 4u;
            plane_edge = (int)(d0.y & 0xffffu) + FS2_W > pc.ow || (int)(d0.y >> 16) - pc.dy + FS2_H > pc.oh;      // block-uniform
            nx = min(max(pc.ow - X, 0), 4);
            row_ok = y < pc.oh;
        }
        unsigned v[4][3];
        if (d0.x & 4u) {
            // Short path (block-uniform; ~2/3 of a ring panorama): ONE camera with weight exactly 1.0f on every pixel of
            // the tile.  Then dst = short(p * 1.0f) = p, dst_w = 1.0f, and normalizeUsingWeightMap gives
            // short(p / (1.0f + 1e-5f)) = p - 1 for p in 1..255 and 0 for p = 0 (the quotient lies strictly between
            // p - 1 and p); the mask is 255.  No float op at all: max(64 * sum ab*p - 64 * 512, 0), byte 2.
            const uint4 rec = sm.desc[stage][1];
            const uint4 e4 = *reinterpret_cast<const uint4 *>(smb + rec.x + tab_off);
            const uint32_t box = rec.x + (uint32_t)FS2_ENT_BYTES;
            const uint32_t ee[4] = {e4.x, e4.y, e4.z, e4.w};
            constexpr int BIAS = (!GAIN && !NOBLEND) ? -32768 : 32768;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                fs2_pixel<BIAS>(smb, box, rec.y, ee[p], v[p][0], v[p][1], v[p][2]);
                if (!GAIN && !NOBLEND) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[p][k] = (unsigned)max((int)v[p][k], 0);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);   // this warp no longer reads the stage
            if (GAIN) {
                float g4[4];
                fs2_gain4(a, smb, rec, gain_idx, X, Y, nx, g4);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float g = g4[p];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        unsigned q = fs2_apply_gain_b2(v[p][k], g);
                        if (!NOBLEND) q -= min(q, 1u);
                        v[p][k] = q;
                    }
                }
                if (PLANES) fs2_store_rgbx<0>(o, v, plane_edge, nx, row_ok);
                else fs2_store_quad<OUT8, 0>(o, v, edge, nx, row_ok);
            } else {
                if (PLANES) fs2_store_rgbx<2>(o, v, plane_edge, nx, row_ok);
                else fs2_store_quad<OUT8, 2>(o, v, edge, nx, row_ok);
            }
            if (!PLANES && mo) fs2_store_mask(mo, 0xffffffffu, edge, nx, row_ok);
        } else if (NOBLEND) {
            // which camera slot supplies each pixel (the last one in feed order with a non-zero mask), OR of the masks
            int sel[4] = {-1, -1, -1, -1};
            unsigned mor = 0u;
            for (int k = 0; k < nc; ++k) {
                const unsigned m4 = *reinterpret_cast<const uint32_t *>(smb + sm.desc[stage][1 + k].x + plane_off);
                mor |= m4;
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if ((m4 >> (8 * p)) & 0xffu) sel[p] = k;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                v[p][0] = v[p][1] = v[p][2] = 0u;
                if (sel[p] < 0) {
                    // feedSizeRemap gathers unconditionally (APP64:165-172): where no camera feeds the look-up tables are
                    // zero, i.e. image 0, pixel (0, 0) of its warped, gain-compensated image
                    if (a.fill_on) {
                        const uint3 f = fs2_fill_pixel<GAIN>(a, src0, sstep0);
                        v[p][0] = f.x; v[p][1] = f.y; v[p][2] = f.z;
                    }
                    continue;
                }
                const uint4 rec = sm.desc[stage][1 + sel[p]];
                const uint32_t e = *reinterpret_cast<const uint32_t *>(smb + rec.x + tab_off + 4u * p);
                fs2_pixel<32768>(smb, rec.x + (uint32_t)FS2_BLOCK_BYTES, rec.y, e, v[p][0], v[p][1], v[p][2]);
                if (GAIN) {
                    const float g = a.gain_tma ? reinterpret_cast<const float *>(smb + rec.x + (rec.z & 0x0fffffffu) + gain_idx)[(rec.z >> 28) + p]
                                               : fs2_gain(a.cam[rec.w & 15u], X + p, Y);
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[p][k] = fs2_apply_gain_b2(v[p][k], g) << 16;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            if (PLANES) {
                if (nc) fs2_store_rgbx<2>(o, v, plane_edge, nx, row_ok);
            } else {
                fs2_store_quad<OUT8, 2>(o, v, edge, nx, row_ok);
                if (mo) fs2_store_mask(mo, mor, edge, nx, row_ok);
            }
        } else {
            // FeatherBlender::feed over the cameras of the tile in feed order (ascending camera index): per channel
            // dst += short(src * w) as integers, dst_w += w in float.  float(v): byte 2 of the sum under the exponent of
            // 2^23, minus 2^23; short(t): bits of (t + 2^23) rounded toward zero, summed as integers.
            float wsum[4] = {0.f, 0.f, 0.f, 0.f};
            unsigned acc[4][3];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0u;
            for (int k = 0; k < nc; ++k) {
                const uint4 rec = sm.desc[stage][1 + k];
                const uint4 e4 = *reinterpret_cast<const uint4 *>(smb + rec.x + tab_off);
                const unsigned d4 = *reinterpret_cast<const uint32_t *>(smb + rec.x + plane_off);
                const uint32_t box = rec.x + (uint32_t)FS2_BLOCK_BYTES;
                const uint32_t ee[4] = {e4.x, e4.y, e4.z, e4.w};
                float g4[4] = {1.f, 1.f, 1.f, 1.f};
                if (GAIN) {
                    const Fs2Cam &fc = a.cam[rec.w & 15u];
                    if (a.gain_tma || !fc.gmap) {
                        fs2_gain4(a, smb, rec, gain_idx, X, Y, nx, g4);
                    } else {                                // block gain map in global memory: only where the camera carries weight (inside its rect)
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            if ((d4 >> (8 * p)) & 0xffu) g4[p] = fs2_gain(fc, X + p, Y);
                    }
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const unsigned dist = (d4 >> (8 * p)) & 0xffu;      // 0: short(p * 0) == 0 and dst_w += 0 -> contributes nothing
                    const float w = fminf(__fmul_rn((float)dist, a.sharpness), 1.f);       // createWeightMap
                    wsum[p] = __fadd_rn(wsum[p], w);
                    unsigned s0, s1, s2;
                    fs2_pixel<32768>(smb, box, rec.y, ee[p], s0, s1, s2);
                    float f0, f1, f2;
                    if (GAIN) {                             // saturate_cast<uchar>(p * gain)
                        const float g = g4[p];
                        // (the rounded product stays a float: bits(x + 2^23) - 2^23)
                        f0 = __fsub_rn(__uint_as_float(fs2_apply_gain_b2(s0, g) | 0x4b000000u), 8388608.f);
                        f1 = __fsub_rn(__uint_as_float(fs2_apply_gain_b2(s1, g) | 0x4b000000u), 8388608.f);
                        f2 = __fsub_rn(__uint_as_float(fs2_apply_gain_b2(s2, g) | 0x4b000000u), 8388608.f);
                    } else {
                        f0 = __fsub_rn(__uint_as_float(__byte_perm(s0, 0x4b000000u, 0x7442)), 8388608.f);
                        f1 = __fsub_rn(__uint_as_float(__byte_perm(s1, 0x4b000000u, 0x7442)), 8388608.f);
                        f2 = __fsub_rn(__uint_as_float(__byte_perm(s2, 0x4b000000u, 0x7442)), 8388608.f);
                    }
                    acc[p][0] += __float_as_uint(__fadd_rz(__fmul_rn(f0, w), 8388608.f));      // static_cast<short>(src * w)
                    acc[p][1] += __float_as_uint(__fadd_rz(__fmul_rn(f1, w), 8388608.f));
                    acc[p][2] += __float_as_uint(__fadd_rz(__fmul_rn(f2, w), 8388608.f));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            // FeatherBlender::blend: normalizeUsingWeightMap, mask = weight > eps, zero unmasked (the sums are already
            // zero there: every w <= 1e-5 makes every short(src * w) zero), convertTo(8U)
            unsigned mword = 0u;
            const unsigned magic = (unsigned)nc * 0x4b000000u;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (wsum[p] > SB_WEIGHT_EPS) mword |= 0xffu << (8 * p);
                const SharedDiv div(__fadd_rn(wsum[p], SB_WEIGHT_EPS));                    // in [1e-5, n + 1e-5]: fast-path range
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float s = __fsub_rn(__uint_as_float((acc[p][k] - magic) | 0x4b000000u), 8388608.f);
                    v[p][k] = __float_as_uint(__fadd_rz(div(s), 8388608.f));
                }
            }
            fs2_store_quad<OUT8, 0>(o, v, edge, nx, row_ok);
            if (mo) fs2_store_mask(mo, mword, edge, nx, row_ok);
        }
        if (tid % (FS2_GROUP_WARPS * 32) == 0) FS2_TRACE(3, seq);
    }
}

// ---- pipeline trace plumbing (SB_FS2_TRACE=file): a pool of 64 buffers, one per launch, dumped on request
static constexpr int FS2_TRACE_POOL = 64;
static unsigned long long *g_trace_pool[FS2_TRACE_POOL] = {};
static int g_trace_grid = 0, g_trace_next = 0;
static bool fs2_trace_enabled()
{
    static const bool on = getenv("SB_FS2_TRACE") != nullptr;
    return on;
}
static int fs2_debug_mode()
{
    static const int m = getenv("SB_FS2_DEBUG") ? atoi(getenv("SB_FS2_DEBUG")) : 0;
    return m;
}
static unsigned long long *fs2_trace_buffer(int grid)
{
    if (!g_trace_pool[0]) {                                  // (first launch: never inside a stream capture, batches warm up eagerly)
        g_trace_grid = grid;
        const size_t bytes = (size_t)grid * FS2_TRACE_TILES * 8 * sizeof(unsigned long long);
        for (auto &p : g_trace_pool) { cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes); }
    }
    return g_trace_pool[g_trace_next++ % FS2_TRACE_POOL];
}
// writes the pool (launch-major) to the file SB_FS2_TRACE names; returns the number of launches since the last dump
int fs2_trace_dump()
{
    if (!fs2_trace_enabled() || !g_trace_pool[0]) return 0;
    cudaDeviceSynchronize();
    const size_t bytes = (size_t)g_trace_grid * FS2_TRACE_TILES * 8 * sizeof(unsigned long long);
    std::vector<unsigned long long> h(bytes / 8);
    FILE *f = fopen(getenv("SB_FS2_TRACE"), "wb");
    if (!f) return -1;
    const int n = std::min(g_trace_next, FS2_TRACE_POOL);
    for (int k = 0; k < n; ++k) {
        cudaMemcpy(h.data(), g_trace_pool[k], bytes, cudaMemcpyDeviceToHost);
        fwrite(h.data(), 1, bytes, f);
        cudaMemset(g_trace_pool[k], 0, bytes);
    }
    fclose(f);
    g_trace_next = 0;
    return n;
}

int launch_fs2(const Fs2Args &a, bool apply_gain, int out_mode, int grid, cudaStream_t s)
{
    const bool out8 = out_mode == 1, planes = out_mode == 2;
    SB_ASSERT(a.sharpness > 0.f && a.desc && a.n_tiles > 0 && a.n <= FS2_TMAP_CAMS);
    SB_ASSERT(a.pw < 65536 && a.ph < 65536);
    if (planes) {
        SB_ASSERT(a.no_blend && !a.out_mask && !a.frames);
        for (int i = 0; i < a.n; ++i) SB_ASSERT(a.cam[i].mb_out && reinterpret_cast<uintptr_t>(a.cam[i].mb_out) % 16 == 0 && a.cam[i].mb_step % 16 == 0);
    } else {
        SB_ASSERT((unsigned long long)a.ph * a.out_step < (1ull << 32) && (unsigned long long)a.ph * a.mask_step < (1ull << 32));
        SB_ASSERT(reinterpret_cast<uintptr_t>(a.out) % 8 == 0 && a.out_step % (out8 ? 4 : 8) == 0);
        SB_ASSERT(!a.out_mask || (reinterpret_cast<uintptr_t>(a.out_mask) % 4 == 0 && a.mask_step % 4 == 0));
    }
    const size_t smem = sizeof(Fs2Smem);
    static bool configured_dev[64][10] = {};                 // the attribute is per device (context)
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    bool *configured = configured_dev[dev & 63];
    const void *fn[10] = {(const void *)k_fs2<false, 0, false>, (const void *)k_fs2<false, 1, false>,
                          (const void *)k_fs2<true, 0, false>, (const void *)k_fs2<true, 1, false>,
                          (const void *)k_fs2<false, 0, true>, (const void *)k_fs2<false, 1, true>,
                          (const void *)k_fs2<true, 0, true>, (const void *)k_fs2<true, 1, true>,
                          (const void *)k_fs2<false, 2, true>, (const void *)k_fs2<true, 2, true>};
    const int v = planes ? 8 + (apply_gain ? 1 : 0) : (a.no_blend ? 4 : 0) + (apply_gain ? 2 : 0) + (out8 ? 1 : 0);
    if (!configured[v]) {
        SB_CUDA(cudaFuncSetAttribute(fn[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[v] = true;
    }
    if (fs2_trace_enabled()) {                               // debugging aid: launch k writes its timestamps to trace buffer k % 64
        Fs2Args at = a;
        at.trace = fs2_trace_buffer(grid);
        at.debug = fs2_debug_mode();
        void *params[] = {&at};
        SB_CUDA(cudaLaunchKernel(fn[v], dim3(grid), dim3(FS2_THREADS), params, smem, s));
        SB_LAUNCHED();
        return SB_OK;
    }
    if (fs2_debug_mode()) {
        Fs2Args ad = a;
        ad.debug = fs2_debug_mode();
        void *params[] = {&ad};
        SB_CUDA(cudaLaunchKernel(fn[v], dim3(grid), dim3(FS2_THREADS), params, smem, s));
        SB_LAUNCHED();
        return SB_OK;
    }
    void *params[] = {const_cast<Fs2Args *>(&a)};
    SB_CUDA(launch_pdl_c(fn[v], dim3(grid), dim3(FS2_THREADS), params, smem, s));
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
