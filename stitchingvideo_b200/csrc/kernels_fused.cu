// kernels_fused.cu — panorama-centric fused kernels of the compositor.
//
// The reference accumulates camera by camera into the panorama (read-modify-write per feed,
// blenders.cpp:123-147, 300-356) and then normalises, collapses and crops in further passes.
// `dst += short(...)` wraps mod 2^16, so the sum over cameras is order independent and every
// zero-weight term is exactly 0: the same result is obtained by visiting each PANORAMA pixel once,
// summing the contributions of the cameras whose weight there is non-zero, and finishing the pixel
// (normalise -> [collapse add] -> mask -> convertTo(8U)) in registers.  No accumulator ever exists
// in HBM: no zero-fill, no read-modify-write, no separate normalise / crop passes.
//
//   k_feather_fused : remap (pre-quantised fixed-point map table, 8 B/px incl. the feather distance)
//                     + gain + convertTo(16S) + FeatherBlender::feed over all cameras + blend +
//                     convertTo(8U)  — ONE launch per frame.
//   k_band_fused    : one Laplacian band of MultiBandBlender for all cameras: Laplacian formed on
//                     the fly from the cameras' Gaussian levels, weighted sum, normalise, collapse
//                     add of the coarser restored band, and at band 0 crop + mask + convertTo(8U).
//
// HBM-bound integer/byte work: each thread owns 4 consecutive panorama pixels so that the weight
// sums load as float4 and the panorama stores as 3 x 32-bit (8U) or 3 x 64-bit (16S) words.
#include <climits>

#include "sb_device.cuh"
#include "sb_fused.h"
#include "sb_pyr.cuh"
#include "sb_warp.cuh"

namespace sb {
using namespace sbd;

#define SB_WEIGHT_EPS 1e-5f

__device__ __forceinline__ bool in_spans(const int span[4], int x0, int x1)   // [x0, x1) overlaps a span?
{
    return (x0 < span[1] && x1 > span[0]) || (x0 < span[3] && x1 > span[2]);
}

// ------------------------------------------------------------------------------------ feather
// setup: one table entry per warped pixel = what cv::remap's map conversion would compute every
// frame (sx, sy, fx, fy; Appendix A1) + the L1 distance behind the feather weight.
template <int KIND>
__global__ void __launch_bounds__(256)
k_build_feather_table(ProjParams p, int tl_x, int tl_y, const float *dist, size_t dstep, int w, int h, uint2 *table, size_t tstep, int tpad)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float mx, my;
    map_backward<KIND>(p, (float)(tl_x + x), (float)(tl_y + y), mx, my);
    const int fsx = cvround(__fmul_rn(mx, 32.f)), fsy = cvround(__fmul_rn(my, 32.f));
    const int sx = sat_s16(fsx >> 5), sy = sat_s16(fsy >> 5);
    const float d = crow<float>(dist, dstep, y)[x];
    const unsigned di = d >= 65535.f ? 65535u : (unsigned)d;       // exact integers below the 8192 saturation
    uint2 t;
    t.x = ((unsigned)sx & 0xffffu) | ((unsigned)sy << 16);
    t.y = (unsigned)(fsx & 31) | ((unsigned)(fsy & 31) << 5) | (di << 16);
    reinterpret_cast<uint2 *>(reinterpret_cast<char *>(table) + (size_t)y * tstep)[x + tpad] = t;
}

int launch_build_feather_table(const ProjParams &p, int tl_x, int tl_y, const DImage &dist, uint2 *table, size_t tstep, int tpad, cudaStream_t s)
{
    SB_ASSERT(dist.type == SB_32FC1);
    dim3 block(32, 8), grid(div_up(dist.cols, 32), div_up(dist.rows, 8));
#define SB_FT(K) k_build_feather_table<K><<<grid, block, 0, s>>>(p, tl_x, tl_y, dist.ptr<float>(), dist.step, dist.cols, dist.rows, table, tstep, tpad)
    switch (p.kind) {
    case SB_WARP_PLANE: SB_FT(SB_WARP_PLANE); break;
    case SB_WARP_CYLINDRICAL: SB_FT(SB_WARP_CYLINDRICAL); break;
    case SB_WARP_SPHERICAL: SB_FT(SB_WARP_SPHERICAL); break;
    default: return fail(SB_ERR_BAD_ARG, "unsupported projector kind %d", p.kind);
    }
#undef SB_FT
    SB_LAUNCHED();
    return SB_OK;
}

// setup: source bounding box per (panorama tile, camera) — sequence-constant, so the per-frame
// kernel can stage exactly the bytes it will sample into shared memory with 16-byte loads.
__global__ void __launch_bounds__(256) k_feather_tile_bbox(FeatherCam c, int pw, int ph, int tiles_x, int4 *bbox)
{
    __shared__ int red[4][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X0 = blockIdx.x * SB_FT_W + (tid & 7) * 4, Y = blockIdx.y * SB_FT_H + (tid >> 3);
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
    const int y = Y - c.dy;
    if (Y < ph && (unsigned)y < (unsigned)c.wh) {
        const uint2 *trow = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + c.tpad;
        for (int j = 0; j < 4; ++j) {
            const int x = X0 + j - c.dx;
            if (X0 + j >= pw || (unsigned)x >= (unsigned)c.ww) continue;
            const uint2 t = trow[x];
            if ((t.y >> 16) == 0u) continue;
            const int sx = (short)(t.x & 0xffffu), sy = (int)t.x >> 16;
            mnx = min(mnx, max(sx, 0)); mxx = max(mxx, min(sx + 1, c.sw - 1));
            mny = min(mny, max(sy, 0)); mxy = max(mxy, min(sy + 1, c.sh - 1));
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            mnx = min(mnx, red[0][w]); mny = min(mny, red[1][w]); mxx = max(mxx, red[2][w]); mxy = max(mxy, red[3][w]);
        }
        bbox[blockIdx.y * tiles_x + blockIdx.x] = (mxx < mnx) ? make_int4(0, 0, -1, -1) : make_int4(mnx, mny, mxx, mxy);
    }
}

int launch_feather_tile_bbox(const FeatherCam &c, int pw, int ph, int4 *bbox, cudaStream_t s)
{
    dim3 block(256), grid(div_up(pw, SB_FT_W), div_up(ph, SB_FT_H));
    k_feather_tile_bbox<<<grid, block, 0, s>>>(c, pw, ph, grid.x, bbox);
    SB_LAUNCHED();
    return SB_OK;
}

__global__ void k_feather_tile_mask(const int4 *bbox, int n_tiles, int cam_index, uint32_t *tile_cams)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles && bbox[i].z >= bbox[i].x) tile_cams[i] |= 1u << cam_index;
}

int launch_feather_tile_mask(const int4 *bbox, int n_tiles, int cam_index, uint32_t *tile_cams, cudaStream_t s)
{
    k_feather_tile_mask<<<div_up(n_tiles, 256), 256, 0, s>>>(bbox, n_tiles, cam_index, tile_cams);
    SB_LAUNCHED();
    return SB_OK;
}

// One launch per frame.  A non-zero distance implies the pixel lies inside the warped all-255 mask,
// i.e. its source coordinate rounds into the image: sx in [-1, sw-1], sy in [-1, sh-1], where
// BORDER_REFLECT coincides with clamping.  Every table weight carries the factor 32, so
// (sum w*p + 2^14) >> 15 == (v + 512) >> 10 with v the sum of 5-bit products; OpenCV's (0,0) entry
// {32767,0,0,1} equals an exact copy for 8-bit data, as does {32768,0,0,0}.  With 8-bit sources and
// weights in [0,1] every intermediate stays inside [0, 255*n]: no saturation or wrap can fire.
//
// The gather is the expensive part (12 byte taps per sample): the source box each tile needs is known
// per calibration, so the block first stages it in shared memory with coalesced 16-byte loads and
// then samples with conflict-free byte LDS (lanes are 12 bytes apart: 3 words, coprime with 32 banks).
template <bool GAIN, bool OUT8>
__global__ void __launch_bounds__(256)
k_feather_fused(const __grid_constant__ FeatherFusedArgs a)
{
    __shared__ __align__(16) uint8_t stage[SB_STAGE_BYTES];
    const int tile = blockIdx.y * a.tiles_x + blockIdx.x;
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * SB_FT_W + (tid & 7) * 4;       // 8 threads x 4 px across, 32 rows down
    const int Y = blockIdx.y * SB_FT_H + (tid >> 3);
    const bool active = X0 < a.pw && Y < a.ph;
    int acc[4][3];
    float wsum[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = 0; wsum[j] = 0.f; }

    for (int i = 0; i < a.n; ++i) {                       // ascending index = feed order (float weight sums)
        const FeatherCam &c = a.cam[i];
        const int4 bb = __ldg(c.bbox + tile);             // block-uniform
        if (bb.z < bb.x) continue;
        const int bw = bb.z - bb.x + 1, bh = bb.w - bb.y + 1;
        const uint8_t *g0 = c.src + (size_t)bb.y * c.sstep + bb.x * 3;
        const unsigned o0 = (unsigned)(reinterpret_cast<uintptr_t>(g0) & 15), os = (unsigned)(c.sstep & 15);
        const int nch = (bw * 3 + 15 + 15) >> 4;          // 16-byte chunks per staged row (upper bound over row alignments)
        const int pitch = nch * 16;
        const bool staged = pitch * bh <= SB_STAGE_BYTES;
        if (staged) {
            for (int idx = tid; idx < bh * nch; idx += 256) {
                const int r = idx / nch, ch = idx - r * nch;
                const uint8_t *gr = g0 + (size_t)r * c.sstep;
                const unsigned orow = (unsigned)(reinterpret_cast<uintptr_t>(gr) & 15);
                if (ch * 16 < (int)orow + bw * 3)
                    *reinterpret_cast<uint4 *>(stage + r * pitch + ch * 16) = __ldg(reinterpret_cast<const uint4 *>(gr - orow) + ch);
            }
            __syncthreads();
        }
        const int y = Y - c.dy;
        const int xb = X0 - c.dx;
        if (active && (unsigned)y < (unsigned)c.wh && xb + 3 >= 0 && xb < c.ww) {
            // the quad's 4 table entries are 32-byte aligned (tpad) and lie inside the padded table row
            const uint4 *tq = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + ((xb + c.tpad) >> 1);
            const uint4 ta = __ldg(tq), tb = __ldg(tq + 1);
            const uint2 te[4] = {make_uint2(ta.x, ta.y), make_uint2(ta.z, ta.w), make_uint2(tb.x, tb.y), make_uint2(tb.z, tb.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = xb + j;
                if ((unsigned)x >= (unsigned)c.ww) continue;
                const uint2 t = te[j];
                const unsigned dist = t.y >> 16;
                if (dist == 0u) continue;                   // weight 0: short(p * 0) == 0 and dst_w += 0
                // createWeightMap: threshold(dist * sharpness, 1, THRESH_TRUNC)
                const float w = fminf(__fmul_rn((float)dist, a.sharpness), 1.f);
                wsum[j] = __fadd_rn(wsum[j], w);
                const int sx = (short)(t.x & 0xffffu), sy = (int)t.x >> 16;
                const int fx = t.y & 31, fy = (t.y >> 5) & 31, ax = 32 - fx, ay = 32 - fy;
                const int x0 = max(sx, 0), x1 = min(sx + 1, c.sw - 1), y0 = max(sy, 0), y1 = min(sy + 1, c.sh - 1);
                int v[3];
                if (staged) {
                    const int r0 = y0 - bb.y, r1 = y1 - bb.y;
                    const uint8_t *s0 = stage + r0 * pitch + ((o0 + r0 * os) & 15) - bb.x * 3;
                    const uint8_t *s1 = stage + r1 * pitch + ((o0 + r1 * os) & 15) - bb.x * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int h0 = (int)s0[x0 * 3 + k] * ax + (int)s0[x1 * 3 + k] * fx;
                        const int h1 = (int)s1[x0 * 3 + k] * ax + (int)s1[x1 * 3 + k] * fx;
                        v[k] = (h0 * ay + h1 * fy + 512) >> 10;
                    }
                } else {
                    const uint8_t *r0 = c.src + (size_t)y0 * c.sstep, *r1 = c.src + (size_t)y1 * c.sstep;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int h0 = (int)__ldg(r0 + x0 * 3 + k) * ax + (int)__ldg(r0 + x1 * 3 + k) * fx;
                        const int h1 = (int)__ldg(r1 + x0 * 3 + k) * ax + (int)__ldg(r1 + x1 * 3 + k) * fx;
                        v[k] = (h0 * ay + h1 * fy + 512) >> 10;
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    int p = v[k];
                    if (GAIN) p = min(max(__float2int_rn(__fmul_rn((float)p, c.gain)), 0), 255);   // saturate_cast<uchar>
                    acc[j][k] += __float2int_rz(__fmul_rn((float)p, w));                            // static_cast<short>(src * w)
                }
            }
        }
        if (staged) __syncthreads();                      // the stage buffer is reused by the next camera
    }
    if (!active) return;
    // FeatherBlender::blend: normalizeUsingWeightMap, mask = weight > eps, zero unmasked, convertTo(8U)
    int o[4][3], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        m[j] = wsum[j] > SB_WEIGHT_EPS ? 255 : 0;
        const SharedDiv div(__fadd_rn(wsum[j], SB_WEIGHT_EPS));        // in [1e-5, n + 1e-5]: fast-path range
#pragma unroll
        for (int k = 0; k < 3; ++k) o[j][k] = m[j] ? __float2int_rz(div((float)acc[j][k])) : 0;
    }
    const bool full = X0 + 4 <= a.pw;
    if (OUT8) {
        uint8_t *orow = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X0 * 3;
        if (full && ((reinterpret_cast<uintptr_t>(orow) & 3) == 0)) {
            uint32_t *q = reinterpret_cast<uint32_t *>(orow);
            q[0] = o[0][0] | (o[0][1] << 8) | (o[0][2] << 16) | (o[1][0] << 24);
            q[1] = o[1][1] | (o[1][2] << 8) | (o[2][0] << 16) | (o[2][1] << 24);
            q[2] = o[2][2] | (o[3][0] << 8) | (o[3][1] << 16) | (o[3][2] << 24);
        } else {
            for (int j = 0; j < 4 && X0 + j < a.pw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (uint8_t)o[j][k];
        }
    } else {
        short *orow = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X0 * 3;
        for (int j = 0; j < 4 && X0 + j < a.pw; ++j)
            for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (short)o[j][k];
    }
    if (a.out_mask) {
        uint8_t *mrow_ = a.out_mask + (size_t)Y * a.mask_step + X0;
        if (full && ((reinterpret_cast<uintptr_t>(mrow_) & 3) == 0))
            *reinterpret_cast<uint32_t *>(mrow_) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
        else
            for (int j = 0; j < 4 && X0 + j < a.pw; ++j) mrow_[j] = (uint8_t)m[j];
    }
}

// Variant with one panorama pixel per thread: a warp's 32 lanes sample 32 neighbouring pixels, so each
// byte-tap request touches 3-4 sectors instead of ~20 (lanes 3 B apart rather than 4 px = 12 B apart).
// One load fetches the tile's camera bitmask; the table entries of every contributing camera are
// requested before any is consumed, so the DRAM latencies of the cameras overlap.
template <bool GAIN, bool OUT8>
__global__ void __launch_bounds__(256)
k_feather_fused_px1(const __grid_constant__ FeatherFusedArgs a)
{
    const int X = blockIdx.x * 32 + threadIdx.x;
    const int Y = blockIdx.y * 8 + threadIdx.y;
    if (X >= a.pw || Y >= a.ph) return;
    // tiles are SB_FT_W x SB_FT_H = 32 x 32 and blocks 32 x 8: the mask is block-uniform
    uint32_t cams = __ldg(a.tile_cams + (blockIdx.y >> 2) * a.tiles_x + blockIdx.x);
    constexpr int MAXC = 4;
    uint2 t[MAXC];
    int ci[MAXC];
    int nc = 0;
#pragma unroll
    for (int q = 0; q < MAXC; ++q) {
        t[q] = make_uint2(0u, 0u);
        ci[q] = 0;
        if (cams) {
            const int i = __ffs(cams) - 1;
            cams &= cams - 1;
            const FeatherCam &c = a.cam[i];
            const int y = Y - c.dy, x = X - c.dx;
            ci[q] = i;
            nc = q + 1;
            if ((unsigned)y < (unsigned)c.wh && (unsigned)x < (unsigned)c.ww)
                t[q] = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + x + c.tpad);
        }
    }
    int acc0 = 0, acc1 = 0, acc2 = 0;
    float wsum = 0.f;
    auto contribute = [&](const FeatherCam &c, const uint2 te) {
        const unsigned dist = te.y >> 16;
        if (dist == 0u) return;                             // weight 0: short(p * 0) == 0 and dst_w += 0
        const float w = fminf(__fmul_rn((float)dist, a.sharpness), 1.f);      // createWeightMap
        wsum = __fadd_rn(wsum, w);
        const int sx = (short)(te.x & 0xffffu), sy = (int)te.x >> 16;
        const int fx = te.y & 31, fy = (te.y >> 5) & 31, ax = 32 - fx, ay = 32 - fy;
        const int x0 = max(sx, 0), x1 = min(sx + 1, c.sw - 1), y0 = max(sy, 0), y1 = min(sy + 1, c.sh - 1);
        const uint8_t *r0 = c.src + (size_t)y0 * c.sstep, *r1 = c.src + (size_t)y1 * c.sstep;
        unsigned q00, q01, q10, q11;                        // packed RGB of the four taps
        if (x1 == x0 + 1) {                                 // the pair is 6 contiguous bytes: 2-3 word loads per row
            load_pixel_pair_8uc3(r0 + x0 * 3, q00, q01);
            load_pixel_pair_8uc3(r1 + x0 * 3, q10, q11);
        } else {                                            // clamped at the image edge
            const uint8_t *p00 = r0 + x0 * 3, *p01 = r0 + x1 * 3, *p10 = r1 + x0 * 3, *p11 = r1 + x1 * 3;
            q00 = __ldg(p00) | (__ldg(p00 + 1) << 8) | (__ldg(p00 + 2) << 16);
            q01 = __ldg(p01) | (__ldg(p01 + 1) << 8) | (__ldg(p01 + 2) << 16);
            q10 = __ldg(p10) | (__ldg(p10 + 1) << 8) | (__ldg(p10 + 2) << 16);
            q11 = __ldg(p11) | (__ldg(p11 + 1) << 8) | (__ldg(p11 + 2) << 16);
        }
        int v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int h0 = (int)((q00 >> (8 * k)) & 0xff) * ax + (int)((q01 >> (8 * k)) & 0xff) * fx;
            const int h1 = (int)((q10 >> (8 * k)) & 0xff) * ax + (int)((q11 >> (8 * k)) & 0xff) * fx;
            v[k] = (h0 * ay + h1 * fy + 512) >> 10;
            if (GAIN) v[k] = min(max(__float2int_rn(__fmul_rn((float)v[k], c.gain)), 0), 255);   // saturate_cast<uchar>
        }
        acc0 += __float2int_rz(__fmul_rn((float)v[0], w));                                          // static_cast<short>(src * w)
        acc1 += __float2int_rz(__fmul_rn((float)v[1], w));
        acc2 += __float2int_rz(__fmul_rn((float)v[2], w));
    };
#pragma unroll
    for (int q = 0; q < MAXC; ++q)
        if (q < nc) contribute(a.cam[ci[q]], t[q]);
    for (; cams; cams &= cams - 1) {                        // more than MAXC overlapping cameras: rest in order
        const FeatherCam &c = a.cam[__ffs(cams) - 1];
        const int y = Y - c.dy, x = X - c.dx;
        if ((unsigned)y < (unsigned)c.wh && (unsigned)x < (unsigned)c.ww)
            contribute(c, __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(c.table) + (size_t)y * c.tstep) + x + c.tpad));
    }
    const int m = wsum > SB_WEIGHT_EPS ? 255 : 0;
    const SharedDiv div(__fadd_rn(wsum, SB_WEIGHT_EPS));
    const int o0 = m ? __float2int_rz(div((float)acc0)) : 0, o1 = m ? __float2int_rz(div((float)acc1)) : 0,
              o2 = m ? __float2int_rz(div((float)acc2)) : 0;
    if (OUT8) {
        uint8_t *o = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X * 3;
        o[0] = (uint8_t)o0; o[1] = (uint8_t)o1; o[2] = (uint8_t)o2;
    } else {
        short *o = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X * 3;
        o[0] = (short)o0; o[1] = (short)o1; o[2] = (short)o2;
    }
    if (a.out_mask) a.out_mask[(size_t)Y * a.mask_step + X] = (uint8_t)m;
}

int launch_feather_fused(const FeatherFusedArgs &a, bool apply_gain, bool out8, cudaStream_t s)
{
    SB_ASSERT(a.sharpness > 0.f);
    SB_ASSERT(div_up(a.pw, SB_FT_W) == a.tiles_x);
    if (a.variant == 1) {
        dim3 block(32, 8), grid(div_up(a.pw, 32), div_up(a.ph, 8));
        if (apply_gain) { if (out8) k_feather_fused_px1<true, true><<<grid, block, 0, s>>>(a); else k_feather_fused_px1<true, false><<<grid, block, 0, s>>>(a); }
        else            { if (out8) k_feather_fused_px1<false, true><<<grid, block, 0, s>>>(a); else k_feather_fused_px1<false, false><<<grid, block, 0, s>>>(a); }
        SB_LAUNCHED();
        return SB_OK;
    }
    dim3 block(256), grid(div_up(a.pw, SB_FT_W), div_up(a.ph, SB_FT_H));
    if (apply_gain) { if (out8) k_feather_fused<true, true><<<grid, block, 0, s>>>(a); else k_feather_fused<true, false><<<grid, block, 0, s>>>(a); }
    else            { if (out8) k_feather_fused<false, true><<<grid, block, 0, s>>>(a); else k_feather_fused<false, false><<<grid, block, 0, s>>>(a); }
    SB_LAUNCHED();
    return SB_OK;
}

// device self-test: SharedDiv vs __fdiv_rn over pseudo-random operands in the ranges the kernels use
__global__ void k_selftest_division(unsigned long long n, unsigned seed, unsigned long long *bad)
{
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        unsigned g = h * 747796405u + 2891336453u;
        g ^= g >> 16; g *= 2246822519u; g ^= g >> 13;
        // divisor: any float with exponent in the accepted range; numerator: integers and general floats
        float d = __uint_as_float(((87u + (h % 81u)) << 23) | (g & 0x7fffffu));
        float a_ = (i & 1) ? (float)((int)(g >> 8) % 65536 - 32768) : __uint_as_float((g & 0x80000000u) | ((87u + ((g >> 3) % 81u)) << 23) | (h & 0x7fffffu));
        if (!div_fast_ok(d, false) || !div_fast_ok(a_, true)) continue;
        SharedDiv sd(d);
        if (__float_as_uint(sd(a_)) != __float_as_uint(__fdiv_rn(a_, d))) ++local;
    }
    if (local) atomicAdd(bad, local);
}

int selftest_division(unsigned long long n, unsigned seed, unsigned long long *mismatches)
{
    unsigned long long *d_bad = nullptr;
    SB_CUDA(cudaMalloc(&d_bad, sizeof *d_bad));
    SB_CUDA(cudaMemset(d_bad, 0, sizeof *d_bad));
    k_selftest_division<<<148 * 8, 256>>>(n, seed, d_bad);
    SB_LAUNCHED();
    SB_CUDA(cudaMemcpy(mismatches, d_bad, sizeof *d_bad, cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    return SB_OK;
}

// ------------------------------------------------------------------------------------ multi-band
__device__ __forceinline__ int band_weighted(int lap, float w) { return (int)trunc_short(__fmul_rn((float)lap, w)); }
__device__ __forceinline__ int band_weighted(int lap, short w) { return (int)(short)((lap * (int)w) >> 8); }
__device__ __forceinline__ int band_normalized(short p, float w) { return trunc_short(__fdiv_rn((float)p, __fadd_rn(w, SB_WEIGHT_EPS))); }
__device__ __forceinline__ int band_normalized(short p, short w)
{
    const int wi = (int)w + 1;
    return wi == 0 ? 0 : (int)(short)((((int)p) << 8) / wi);
}
__device__ __forceinline__ bool band_masked(float w) { return w > SB_WEIGHT_EPS; }
__device__ __forceinline__ bool band_masked(short w) { return w > 0; }

// One band l for all cameras.  HAS_FINER_SRC: the cameras have a coarser Gaussian level (l < n);
// HAS_COARSE_R: a restored coarser band exists to be added (l < n); FINAL: l == 0.
template <typename WT, bool NOT_TOP, bool FINAL, bool OUT8>
__global__ void __launch_bounds__(128)
k_band_fused(const __grid_constant__ BandFusedArgs a)
{
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    const int lw = FINAL ? a.out_w : a.lw, lh = FINAL ? a.out_h : a.lh;    // band 0 is cropped to dst_roi_final_
    if (X0 >= lw || Y >= lh) return;
    int acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0;

    for (int i = 0; i < a.n; ++i) {
        const BandCam &c = a.cam[i];
        const int y = Y - c.ry;
        if ((unsigned)y >= (unsigned)c.rh) continue;
        if (!in_spans(c.span, X0, X0 + 4)) continue;
        const WT *wrow = crow<WT>(c.weight, c.wstep, y);
        const short *frow = crow<short>(c.fine, c.fstep, y);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = X0 + j - c.rx;
            if ((unsigned)x >= (unsigned)c.rw) continue;
            const WT w = wrow[x];
            if (w == (WT)0) continue;                     // short(lap * 0) == 0 and (lap * 0) >> 8 == 0
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int lap = frow[x * 3 + k];
                if (NOT_TOP) {
                    const int up = up_cast<short>(pyr_up_sum<short, 3>(c.coarse, c.cstep, c.rw >> 1, c.rh >> 1, y, x, k));
                    lap = sat_s16(lap - up);              // subtract(pyr[i], tmp, pyr[i]) saturates
                }
                acc[j][k] += band_weighted(lap, w);
            }
        }
    }
    const WT *ws = crow<WT>(a.wsum, a.wsum_step, Y) + X0;
    int o[4][3], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool in = X0 + j < lw;
        const WT w = in ? ws[j] : (WT)0;
        m[j] = band_masked(w) ? 255 : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int v = band_normalized((short)acc[j][k], w);
            if (NOT_TOP && in)   // restoreImageFromLaplacePyr: add(pyrUp(pyr[i+1]), pyr[i]) saturates
                v = sat_s16(up_cast<short>(pyr_up_sum<short, 3>(a.coarse_r, a.coarse_r_step, a.lw >> 1, a.lh >> 1, Y, X0 + j, k)) + v);
            if (FINAL) { v = m[j] ? v : 0; if (OUT8) v = sat_u8(v); }
            o[j][k] = v;
        }
    }
    if (FINAL && OUT8) {
        uint8_t *orow = reinterpret_cast<uint8_t *>(a.out) + (size_t)Y * a.out_step + X0 * 3;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(orow) & 3) == 0)) {
            uint32_t *q = reinterpret_cast<uint32_t *>(orow);
            q[0] = o[0][0] | (o[0][1] << 8) | (o[0][2] << 16) | (o[1][0] << 24);
            q[1] = o[1][1] | (o[1][2] << 8) | (o[2][0] << 16) | (o[2][1] << 24);
            q[2] = o[2][2] | (o[3][0] << 8) | (o[3][1] << 16) | (o[3][2] << 24);
        } else {
            for (int j = 0; j < 4 && X0 + j < lw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (uint8_t)o[j][k];
        }
    } else {
        short *orow = reinterpret_cast<short *>(reinterpret_cast<char *>(a.out) + (size_t)Y * a.out_step) + X0 * 3;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(orow) & 7) == 0)) {
            uint2 *q = reinterpret_cast<uint2 *>(orow);
            auto pk = [](int lo, int hi) { return (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16); };
            q[0] = make_uint2(pk(o[0][0], o[0][1]), pk(o[0][2], o[1][0]));
            q[1] = make_uint2(pk(o[1][1], o[1][2]), pk(o[2][0], o[2][1]));
            q[2] = make_uint2(pk(o[2][2], o[3][0]), pk(o[3][1], o[3][2]));
        } else {
            for (int j = 0; j < 4 && X0 + j < lw; ++j)
                for (int k = 0; k < 3; ++k) orow[j * 3 + k] = (short)o[j][k];
        }
    }
    if (FINAL && a.out_mask) {
        uint8_t *mrow_ = a.out_mask + (size_t)Y * a.mask_step + X0;
        if (X0 + 4 <= lw && ((reinterpret_cast<uintptr_t>(mrow_) & 3) == 0))
            *reinterpret_cast<uint32_t *>(mrow_) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
        else
            for (int j = 0; j < 4 && X0 + j < lw; ++j) mrow_[j] = (uint8_t)m[j];
    }
}

int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s)
{
    const int lw = final_band ? a.out_w : a.lw, lh = final_band ? a.out_h : a.lh;
    dim3 block(32, 4), grid(div_up(div_up(lw, 4), 32), div_up(lh, 4));
#define SB_BF(WT, NT, FIN, O8) k_band_fused<WT, NT, FIN, O8><<<grid, block, 0, s>>>(a)
#define SB_BF_W(WT)                                                                          \
    do {                                                                                     \
        if (final_band) { if (not_top) { if (out8) SB_BF(WT, true, true, true); else SB_BF(WT, true, true, false); } \
                          else { if (out8) SB_BF(WT, false, true, true); else SB_BF(WT, false, true, false); } }     \
        else { if (not_top) SB_BF(WT, true, false, false); else SB_BF(WT, false, false, false); }                    \
    } while (0)
    if (float_weights) SB_BF_W(float); else SB_BF_W(short);
#undef SB_BF_W
#undef SB_BF
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------------------------ spans
// flags[x] = 1 if column x of the weight image holds any non-zero weight (setup time only)
template <typename WT> __global__ void k_column_nonzero(const WT *w, size_t wstep, int cols, int rows, int *flags)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cols) return;
    int any = 0;
    for (int y = 0; y < rows && !any; ++y) any |= crow<WT>(w, wstep, y)[x] != (WT)0;
    flags[x] = any;
}

int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s)
{
    SB_ASSERT(w.type == SB_32FC1 || w.type == SB_16SC1);
    if (w.type == SB_32FC1) k_column_nonzero<float><<<div_up(w.cols, 128), 128, 0, s>>>(w.ptr<float>(), w.step, w.cols, w.rows, flags);
    else k_column_nonzero<short><<<div_up(w.cols, 128), 128, 0, s>>>(w.ptr<short>(), w.step, w.cols, w.rows, flags);
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
