// sb_mb.h — argument blocks of the multi-band fast path (kernels_mb.cu).
#pragma once
#include <cuda.h>          // CUtensorMap (types only)
#include "sb_fused.h"

namespace sb {

struct MbWarpCam {
    const uint8_t *src;    // 8UC3 source frame
    size_t sstep;
    const uint2 *table;    // resolved bilinear taps per padded-rect pixel (sequence-constant)
    size_t tstep;
    uint32_t *g0;          // Gaussian level 0 of the padded warped image, RGBX bytes
    size_t gstep;
    int rw, rh;            // padded feed rect (blenders.cpp:241-269)
    float gain;
    const float *gmap;     // SB_COMP_GAIN_BLOCKS: per padded-rect pixel gain (null: the scalar gain)
    size_t gmstep;
    int cx[4];             // columns of the rect to produce: [cx[0], cx[1]) and [cx[2], cx[3]) (empty runs have lo >= hi)
};
struct MbWarpArgs {
    int n;
    const uint2 *bilin_lut;   // 1024 x {hi, lo} bilinear product weights (sb_device.cuh)
    MbWarpCam cam[SB_MAX_CAMERAS];
};

struct MbPyrCam {
    const uint32_t *src;
    size_t sstep;
    int sw, sh;
    uint32_t *dst;         // ((sw+1)/2, (sh+1)/2)
    size_t dstep;
    int ox[4];             // output columns to produce: [ox[0], ox[1]) and [ox[2], ox[3])
};
struct MbPyrArgs {
    int n;
    MbPyrCam cam[SB_MAX_CAMERAS];
};

struct MbPyrSeg { int cam, tx0, ntx, first; };          // tiles tx0 .. tx0 + ntx - 1 (64 output columns each) of one camera, all tile rows
struct MbPyrListArgs {
    MbPyrArgs p;
    int n_seg;
    MbPyrSeg seg[2 * SB_MAX_CAMERAS];
};

// tile-staged pyrDown (kernels_mb_pyr.cu): a CTA = 64 x 32 outputs from one tensor copy of 136 x 67 inputs
#ifndef SB_CFG_PT_OH
#define SB_CFG_PT_OH 32
#endif
constexpr int MB_PT_OW = 64, MB_PT_OH = SB_CFG_PT_OH, MB_PT_IW = 2 * MB_PT_OW + 8, MB_PT_IH = 2 * MB_PT_OH + 3;
struct alignas(64) MbPyrTmaArgs {
    CUtensorMap map[SB_MAX_CAMERAS];                        // level l of camera i as a 2-D tensor of 32-bit pixels, box MB_PT_IW x MB_PT_IH
    MbPyrListArgs list;                                     // cameras + the tile list (here: MB_PT_OW-column tiles, MB_PT_OH-row tile rows)
};
int launch_mb_pyr_down_tma(MbPyrTmaArgs &a, cudaStream_t s);

struct MbBandCam {
    const uint32_t *fine;    // Gaussian level l (RGBX), rect-local
    unsigned fstep;          // (row pitches as 32-bit values: one IMAD.WIDE.U32 per row address instead of a 64-bit multiply)
    const uint32_t *coarse;  // Gaussian level l+1 (null at the top level)
    unsigned cstep;
    const void *weight;      // weight pyramid level l (float or short), sequence-constant
    unsigned wstep;
    int rx, ry, rw, rh;      // feed rect at this level in panorama-level coordinates
};
struct MbBandGeom {
    int n;
    MbBandCam cam[SB_MAX_CAMERAS];
    int lw, lh;              // size of band l (padded panorama >> l)
};
struct MbBandArgs {
    MbBandGeom g;
    const uint32_t *tile_mask;   // per 32x8 tile of band l: cameras with non-zero weight there
    int tiles_x;
    const void *wsum;        // dst_band_weights_[l]
    unsigned wsum_step;
    const short4 *coarse_r;  // restored band l+1 (CV_16SC3 values in 8-byte pixels); null at the top level
    unsigned coarse_r_step;
    void *out;               // restored band l (short4 pixels), or the final panorama at band 0
    unsigned out_step;
    uint8_t *out_mask;
    unsigned mask_step;
    int out_w, out_h;        // band 0 only: dst_roi_final_ size
    int x_begin, x_end;      // band columns to produce (strip mode; whole band: 0, lw)
};

// several coarse levels in one cooperative launch (kernels_mb.cu "multi-level launches")
constexpr int SB_MB_MAX_FUSED_LEVELS = 6;
struct MbPyrTailArgs {
    int n_levels;                                          // level[k]: Gaussian level l0 + k -> l0 + k + 1
    MbPyrArgs level[SB_MB_MAX_FUSED_LEVELS];
    int items[SB_MB_MAX_FUSED_LEVELS];                     // 32x8-thread work items of the level (all cameras)
    int first_item[SB_MB_MAX_FUSED_LEVELS][SB_MAX_CAMERAS];
    int tiles_x[SB_MB_MAX_FUSED_LEVELS][SB_MAX_CAMERAS];
};
struct MbBandHeadArgs {
    int n_levels;                                          // level[0] = the coarsest band of the run, then finer ones
    int top_is_top;                                        // level[0] is the top of the pyramid (no coarser level)
    MbBandArgs level[SB_MB_MAX_FUSED_LEVELS];
    int items[SB_MB_MAX_FUSED_LEVELS];
    int tiles_x[SB_MB_MAX_FUSED_LEVELS];
};
// every coarse level in ONE ordinary launch: stages drawn from a work counter in order, stage s waits for stage s - 1's count
struct MbCoarseArgs {
    MbPyrTailArgs down;                                    // Gaussian levels l0 -> ... (n_levels may be 0)
    MbBandHeadArgs band;                                   // then the bands, coarsest first
    int first[2 * SB_MB_MAX_FUSED_LEVELS + 1];             // first work item of each stage (filled by the launcher)
    int total;
    unsigned long long *sync;                              // per slot: [0] work counter, [1 + s] items of stage s done; only ever grow
    unsigned long long base;                               // value of the work counter when this launch starts
    unsigned long long want[2 * SB_MB_MAX_FUSED_LEVELS];   // value of stage s's counter when all its items of this launch are done
};
int launch_mb_coarse(MbCoarseArgs &a, unsigned long long *totals, bool float_weights, int sm_count, bool alone, cudaStream_t s);
int launch_mb_pyr_tail(const MbPyrTailArgs &a, int sm_count, cudaStream_t s);
int launch_mb_band_head(const MbBandHeadArgs &a, bool float_weights, int sm_count, cudaStream_t s);

// ---- streaming form of the warp stage (kernels_mb_stream.cu, machinery of sb_stream.cuh) ----
struct MbStreamCam {
    const uint8_t *src;    // 8UC3 source frame, 16-byte aligned, sstep a multiple of 16
    unsigned sstep;
    const uint2 *tiles;    // tile-major tap entries of this camera's padded rect
    uint32_t *g0;          // Gaussian level 0 of the padded warped image, RGBX bytes
    unsigned gstep;
    int rw, rh;
    float gain;
    const float *gmap;     // SB_COMP_GAIN_BLOCKS: per padded-rect pixel gain (null: the scalar gain)
    unsigned gmstep;
};
struct MbStreamArgs {
    int n;
    MbStreamCam cam[SB_MAX_CAMERAS];
    const uint4 *desc;     // per listed tile: 1 + SB_FTT_MAXC records, in schedule order, with the ring plan
    const uint2 *bilin_lut;
    int n_tiles;
};
struct MbsSetup {
    const uint4 *rec[SB_MAX_CAMERAS];   // per camera, per tile of its padded rect: source box record
    int ntx[SB_MAX_CAMERAS];
};
int launch_mbs_camera_tiles(const uint2 *table, size_t tstep, int rw, int rh, int ntx, int nty, uint4 *rec, uint2 *tiles, cudaStream_t s);
// list_dev[t] = camera | tile block index << 4 for the tiles that have to be produced; grid = fts_grid()
int launch_mbs_descriptors(const MbsSetup &a, const unsigned *list_dev, int n_tiles, uint4 *desc, int grid, cudaStream_t s);
int launch_mb_warp_stream(const MbStreamArgs &a, bool apply_gain, int sm_count, cudaStream_t s);
// the padded rect's taps in the row-major format of the k_fs2 setup (mask 255 inside the column runs cx)
int launch_mbs_feather_format(const uint2 *table, size_t tstep, int rw, int rh, const int cx[4], uint2 *out, size_t ostep, cudaStream_t s);

int launch_mb_tap_table(const ProjParams &p, int tl_x, int tl_y, int ww, int wh, int left, int top, int sw, int sh,
                        uint2 *table, size_t tstep, int rw, int rh, cudaStream_t s, const DImage *xmap = nullptr, const DImage *ymap = nullptr);
int launch_mb_tile_mask(const MbBandGeom &g, bool float_weights, int lw, int lh, const void *wsum, size_t wsum_step, uint32_t *mask, cudaStream_t s);
int launch_mb_warp(const MbWarpArgs &a, bool apply_gain, int max_rw, int max_rh, cudaStream_t s);
int launch_mb_pyr_down(const MbPyrArgs &a, int max_dw, int max_dh, cudaStream_t s);
int launch_mb_pyr_down_list(MbPyrListArgs &a, cudaStream_t s);      // the same over the compacted tile list (fills a.seg)
int launch_mb_band(const MbBandArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);

}  // namespace sb
