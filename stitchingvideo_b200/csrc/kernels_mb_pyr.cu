// kernels_mb_pyr.cu — cv::pyrDown on the RGBX Gaussian levels of the multi-band path, tile-staged by tensor copy.
//
// Reference: createLaplacePyr's pyrDown chain (blenders.cpp:435-452 -> cv::pyrDown): 5x5 binomial
// kernel, BORDER_REFLECT_101, (sum + 128) >> 8 on 8-bit values (after warp + gain every Gaussian level of the path is
// 8-bit; the CV_16S levels of the reference hold the same numbers).
//
// k_mb_pyr_down (kernels_mb.cu) computes a 2x2 output block per thread from a 7x7 window it gathers through L1: every
// input pixel is requested 12 times, the loads of a thread are dependent on nothing but still issue one by one, and the
// kernel sat at 0.28-0.32 of the HBM roofline with 40 % occupancy (profiles/r01_mb_v6, r02_mb_c3).  Here a CTA owns a
// 64 x 32 output tile: ONE tensor copy (cp.async.bulk.tensor.2d) brings its 136 x 67 input pixels into shared memory -
// out-of-image elements arrive as zero and the <= 2 reflected columns / rows are patched in shared memory - and each
// thread walks 11 staged rows for a 2-wide, 4-high output column: horizontal 5-tap sums once per input row (3 shared
// loads), vertical sums from registers.  Several CTAs per SM overlap one tile's copy with another's arithmetic.
#include <algorithm>

#include "sb_device.cuh"
#include "sb_fs2.h"
#include "sb_mb.h"
#include "sb_tma.cuh"

namespace sb {
using namespace sbd;
using namespace sbt;

__device__ __forceinline__ unsigned pt_lanes02(unsigned v) { return v & 0x00ff00ffu; }
__device__ __forceinline__ unsigned pt_lane1(unsigned v) { return __byte_perm(v, 0u, 0x4441); }

__global__ void __launch_bounds__(256) k_mb_pyr_down_tma(const __grid_constant__ MbPyrTmaArgs a)
{
    __shared__ __align__(128) uint32_t tile[MB_PT_IH][MB_PT_IW];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_release();
    int k = 0;
    while (k + 1 < a.list.n_seg && (int)blockIdx.x >= a.list.seg[k + 1].first) ++k;
    const MbPyrSeg sg = a.list.seg[k];
    const MbPyrCam &c = a.list.p.cam[sg.cam];
    const int local = (int)blockIdx.x - sg.first, bx = sg.tx0 + local % sg.ntx, by = local / sg.ntx;
    const int dw = (c.sw + 1) >> 1, dh = (c.sh + 1) >> 1;
    const int X0 = bx * MB_PT_OW, Y0 = by * MB_PT_OH;
    const int ix0 = 2 * X0 - 4, iy0 = 2 * Y0 - 2;           // input coordinates of tile[0][0] (16-byte aligned column)
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&bar, (unsigned)sizeof(tile));
        tma_load_2d(smem_u32(&tile[0][0]), &a.map[sg.cam], ix0, iy0, &bar);
    }
    __syncthreads();                                        // (the barrier's initialisation before anyone waits on it)
    mbar_wait_hw(&bar, 0u);

    // BORDER_REFLECT_101: the window of an output pixel reaches 2 pixels past the image at most
    if (ix0 < 0 || ix0 + MB_PT_IW > c.sw) {
        for (int e = tid; e < MB_PT_IH * 4; e += 256) {
            const int r = e >> 2, q = e & 3;
            const int x = q < 2 ? q - 2 : c.sw + (q - 2);   // -2, -1, sw, sw + 1
            const int j = x - ix0, y = iy0 + r;
            if (j >= 0 && j < MB_PT_IW && y >= 0 && y < c.sh) tile[r][j] = tile[r][reflect101(x, c.sw) - ix0];
        }
        __syncthreads();
    }
    if (iy0 < 0 || iy0 + MB_PT_IH > c.sh) {
        for (int e = tid; e < MB_PT_IW * 4; e += 256) {
            const int j = e >> 2, q = e & 3;
            const int y = q < 2 ? q - 2 : c.sh + (q - 2);
            const int r = y - iy0;
            if (r >= 0 && r < MB_PT_IH) tile[r][j] = tile[reflect101(y, c.sh) - iy0][j];
        }
        __syncthreads();
    }

    const int cp = tid & 31, rg = tid >> 5;
    constexpr int R = MB_PT_OH / 8, NR = 2 * R + 3;         // output rows per thread, staged rows they need
    const int X = X0 + 2 * cp, Yb = Y0 + R * rg;
    if (X >= dw || Yb >= dh) return;
    if (!((X + 1 >= c.ox[0] && X < c.ox[1]) || (X + 1 >= c.ox[2] && X < c.ox[3]))) return;
    unsigned ha02[NR], ha1[NR], hb02[NR], hb1[NR];          // horizontal sums of output columns X and X + 1, per staged row
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const uint32_t *row = &tile[2 * R * rg + i][4 * cp + 2];
        const uint2 p = *reinterpret_cast<const uint2 *>(row);
        const uint4 q = *reinterpret_cast<const uint4 *>(row + 2);
        const unsigned v6 = row[6];
        const unsigned e0 = pt_lanes02(p.x), e1 = pt_lanes02(p.y), e2 = pt_lanes02(q.x), e3 = pt_lanes02(q.y), e4 = pt_lanes02(q.z),
                       e5 = pt_lanes02(q.w), e6 = pt_lanes02(v6);
        const unsigned g0 = pt_lane1(p.x), g1 = pt_lane1(p.y), g2 = pt_lane1(q.x), g3 = pt_lane1(q.y), g4 = pt_lane1(q.z), g5 = pt_lane1(q.w),
                       g6 = pt_lane1(v6);
        ha02[i] = e2 * 6u + (e1 + e3) * 4u + e0 + e4; hb02[i] = e4 * 6u + (e3 + e5) * 4u + e2 + e6;
        ha1[i] = g2 * 6u + (g1 + g3) * 4u + g0 + g4;  hb1[i] = g4 * 6u + (g3 + g5) * 4u + g2 + g6;
    }
    // vertical: <= 256 * 255 (+128) per 16-bit lane, no carry between lanes; (s + 128) >> 8
#pragma unroll
    for (int oy = 0; oy < R; ++oy) {
        const int y = Yb + oy;
        if (y >= dh) break;
        const unsigned *A02 = ha02 + 2 * oy, *A1 = ha1 + 2 * oy, *B02 = hb02 + 2 * oy, *B1 = hb1 + 2 * oy;
        const unsigned sa02 = A02[2] * 6u + (A02[1] + A02[3]) * 4u + A02[0] + A02[4] + 0x00800080u;
        const unsigned sa1 = A1[2] * 6u + (A1[1] + A1[3]) * 4u + A1[0] + A1[4] + 128u;
        const unsigned sb02 = B02[2] * 6u + (B02[1] + B02[3]) * 4u + B02[0] + B02[4] + 0x00800080u;
        const unsigned sb1 = B1[2] * 6u + (B1[1] + B1[3]) * 4u + B1[0] + B1[4] + 128u;
        const unsigned oa = ((sa02 >> 8) & 0x00ff00ffu) | ((sa1 >> 8) << 8), ob = ((sb02 >> 8) & 0x00ff00ffu) | ((sb1 >> 8) << 8);
        uint32_t *d = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(c.dst) + (size_t)y * c.dstep) + X;
        if (X + 1 < dw) *reinterpret_cast<uint2 *>(d) = make_uint2(oa, ob);
        else *d = oa;
    }
}

// a.list.p (cameras, wanted output runs) and a.map filled by the caller; builds the tile list and launches
int launch_mb_pyr_down_tma(MbPyrTmaArgs &a, cudaStream_t s)
{
    a.list.n_seg = 0;
    int total = 0;
    for (int i = 0; i < a.list.p.n; ++i) {
        const MbPyrCam &c = a.list.p.cam[i];
        const int dw = (c.sw + 1) >> 1, dh = (c.sh + 1) >> 1;
        const int rows = div_up(dh, MB_PT_OH);
        int t0[2], t1[2], nr = 0;
        for (int r = 0; r < 2; ++r) {
            const int lo = std::max(0, c.ox[2 * r]), hi = std::min(dw, c.ox[2 * r + 1]);
            if (hi <= lo) continue;
            t0[nr] = lo / MB_PT_OW; t1[nr] = div_up(hi, MB_PT_OW); ++nr;
        }
        if (nr == 2 && t0[1] <= t1[0]) { t1[0] = std::max(t1[0], t1[1]); nr = 1; }      // (runs are ascending)
        for (int r = 0; r < nr; ++r) {
            a.list.seg[a.list.n_seg++] = MbPyrSeg{i, t0[r], t1[r] - t0[r], total};
            total += (t1[r] - t0[r]) * rows;
        }
    }
    if (total == 0) return SB_OK;
    SB_CUDA(launch_pdl(k_mb_pyr_down_tma, dim3(total), dim3(256), 0, s, a));
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
