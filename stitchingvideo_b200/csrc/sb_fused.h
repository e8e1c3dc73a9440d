// sb_fused.h — argument blocks of the panorama-centric fused kernels (kernels_fused.cu).
#pragma once
#include "sb_kernels.h"

namespace sb {

constexpr int SB_MAX_CAMERAS = SB_MAX_COMPOSITOR_CAMERAS;

// Per-camera, per-warped-pixel table of the feather path (sequence-constant, 8 bytes/pixel): the
// resolved taps of cv::remap's fixed-point map entry (A1) and the L1 distance createWeightMap turns
// into the weight.
//   x: x0 | y0 << 13 | (x1 == x0) << 26 | (y1 == y0) << 27 | unrepresentable << 28
//   y: fx | fy << 5 (low 10 bits = bilinear product table index) | dist << 16
// dist == 0 marks a zero-weight pixel (outside the warped mask): the kernel skips it, exactly.
struct FeatherCam {
    const uint8_t *src;
    size_t sstep;
    const uint2 *table;    // row y of the warped image starts at table + y*tstep
    size_t tstep;          // bytes per table row
    int ww, wh;            // warped size
    int dx, dy;            // warped corner in panorama coordinates
    float gain;
    const float *gmap;     // SB_COMP_GAIN_BLOCKS: gain per warped pixel (null: the scalar gain)
    size_t gmstep;
};

constexpr int SB_FT_W = 32, SB_FT_H = 8;         // panorama pixels per block of k_feather_fused_px1 = camera-mask tile

struct FeatherFusedArgs {
    int n;
    FeatherCam cam[SB_MAX_CAMERAS];
    int tiles_x;
    const uint32_t *tile_cams;   // per panorama tile: bitmask of the cameras with non-zero weight there
    const uint2 *bilin_lut;      // 1024 x {hi, lo} bilinear product weights (sb_device.cuh)
    float sharpness;
    void *out;             // 8UC3 or 16SC3 panorama
    size_t out_step;
    uint8_t *out_mask;     // may be null
    size_t mask_step;
    int pw, ph;
};

// ---- persistent, warp-specialised streaming feather kernel (kernels_feather_tma.cu) ----
// Shape knobs (compile-time, -DSB_CFG_*: round 1 timed 32x16 / 32x32 / 32x64 tiles with 8 or 16 consumer warps; 32x32 x 16 won)
#ifndef SB_CFG_FTT_H
#define SB_CFG_FTT_H 32
#endif
#ifndef SB_CFG_FTS_CONSUMER_WARPS
#define SB_CFG_FTS_CONSUMER_WARPS 16
#endif
#ifndef SB_CFG_FTS_PRODUCER_WARPS
#define SB_CFG_FTS_PRODUCER_WARPS 4
#endif
#ifndef SB_CFG_FTS_CTAS_PER_SM
#define SB_CFG_FTS_CTAS_PER_SM 2
#endif
#ifndef SB_CFG_FTS_RING_KB
#define SB_CFG_FTS_RING_KB 100
#endif
constexpr int SB_FTT_W = 32, SB_FTT_H = SB_CFG_FTT_H;   // panorama tile: one contiguous table block (8 B per pixel) per camera
constexpr int SB_FTT_TAB_BYTES = SB_FTT_W * SB_FTT_H * 8;
constexpr int SB_FTT_MAXC = 3;                   // cameras with weight inside one tile (more -> k_feather_fused_px1)
constexpr int SB_FTT_STAGES = 16;                // tile entries (descriptor + barriers) in flight per CTA
constexpr int SB_FTS_RING_BYTES = SB_CFG_FTS_RING_KB * 1024;   // shared-memory ring: per (tile, camera) table block + source box, byte-granular
constexpr int SB_FTS_BOX_BYTES = 12288;          // largest source box staged in shared memory; larger boxes are gathered directly
constexpr int SB_FTS_MAX_ROWS = 96;              // box rows (< SB_FTS_DIRECT)
constexpr int SB_FTS_DIRECT = 0xff;              // n_rows marker: no box, taps come from global memory
constexpr int SB_FTS_CONSUMER_WARPS = SB_CFG_FTS_CONSUMER_WARPS;   // pixel threads = 32 x this
constexpr int SB_FTS_PRODUCER_WARPS = SB_CFG_FTS_PRODUCER_WARPS;   // all walk the tile sequence; warp w fetches tiles w, w + P, ...
constexpr int SB_FTS_CTAS_PER_SM = SB_CFG_FTS_CTAS_PER_SM;
static_assert(SB_FTT_MAXC * (SB_FTT_TAB_BYTES + SB_FTS_BOX_BYTES) <= SB_FTS_RING_BYTES, "one tile must fit the ring");
constexpr int SB_FTS_THREADS = (SB_FTS_CONSUMER_WARPS + SB_FTS_PRODUCER_WARPS) * 32;

struct FeatherTmaCam {
    const uint8_t *src;    // 16-byte aligned, sstep a multiple of 16
    unsigned sstep;
    float gain;
    const float *gmap;     // SB_COMP_GAIN_BLOCKS: gain per warped pixel (null: the scalar gain)
    unsigned gmstep;
    int dx, dy;            // warped corner in panorama coordinates (gain map lookup)
    const uint2 *tiles;    // tile-major table blocks of this camera
};
struct FeatherTmaArgs {
    int n;
    FeatherTmaCam cam[SB_MAX_CAMERAS];
    const uint4 *desc;           // per panorama tile: 1 + SB_FTT_MAXC records (kernels_feather_tma.cu)
    const uint2 *bilin_lut;
    float sharpness;
    int no_blend;                // Blender::NO: the table's distance field is the mask byte; last camera with a non-zero mask wins
    void *out;
    size_t out_step;
    uint8_t *out_mask;
    size_t mask_step;
    int pw, ph, tiles_x, n_tiles;
};
struct FtsSetupCam {
    const uint4 *rec;      // per tile block of this camera: source box record
    int tx0, ty0, ntx, nty;
};
struct FtsSetup {
    int n;
    FtsSetupCam cam[SB_MAX_CAMERAS];
    int tiles_x, n_tiles;
};
// setup: box records + tile-major entries of one camera
int launch_fts_camera_tiles(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                            float sharpness, uint4 *rec, uint2 *tiles, cudaStream_t s);
// setup: the per-tile descriptors; *status != 0 -> the streaming kernel cannot be used for this calibration
int launch_fts_descriptors(const FtsSetup &a, uint4 *desc, int *status, int grid, cudaStream_t s);   // grid = fts_grid(): the schedule and ring plan are per grid size
int fts_grid(int n_tiles, int sm_count);
// schedule order (descending cost) + ring plan of a descriptor array, in place
int fts_schedule(uint4 *desc, int n_tiles, int grid, cudaStream_t s);
int launch_feather_stream(const FeatherTmaArgs &a, bool apply_gain, bool out8, int sm_count, cudaStream_t s);

// one camera as seen by k_band_fused at one pyramid level
struct BandCam {
    const short *fine;     // Gaussian level l of the padded warped image (rect-local)
    size_t fstep;
    const short *coarse;   // Gaussian level l+1 (null at the top level)
    size_t cstep;
    const void *weight;    // weight pyramid level l (float or short)
    size_t wstep;
    int rx, ry, rw, rh;    // feed rect at this level, in panorama-level coordinates
    int span[4];
};

struct BandFusedArgs {
    int n;
    BandCam cam[SB_MAX_CAMERAS];
    const void *wsum;      // dst_band_weights_[l]
    size_t wsum_step;
    const short *coarse_r; // restored band l+1 (null at the top level)
    size_t coarse_r_step;
    void *out;             // restored band l (16SC3), or the final panorama at band 0
    size_t out_step;
    uint8_t *out_mask;
    size_t mask_step;
    int lw, lh;            // size of band l (padded panorama >> l)
    int out_w, out_h;      // band 0 only: dst_roi_final_ size
};

int launch_build_bilin_lut(uint2 *lut, cudaStream_t s);
// setup: tile_cams[tile] |= 1 << cam_index where the camera has weight; *bad += weighted entries with unrepresentable taps
int launch_feather_tile_cams(const FeatherCam &c, int cam_index, int pw, int ph, uint32_t *tile_cams, unsigned *bad, cudaStream_t s);
int launch_feather_fused(const FeatherFusedArgs &a, bool apply_gain, bool out8, cudaStream_t s);
// setup: resolved-tap + distance table of one camera (dist: CV_32FC1 output of distanceTransform)
// (xmap / ymap: the warped image's float maps for the projectors whose mapBackward is evaluated on the host; null: recomputed)
int launch_build_feather_table(const ProjParams &p, int tl_x, int tl_y, const DImage &dist, int sw, int sh, uint2 *table, size_t tstep, cudaStream_t s,
                               const DImage *xmap = nullptr, const DImage *ymap = nullptr);
// device self-test of SharedDiv against __fdiv_rn; returns the number of mismatching quotients
int selftest_division(unsigned long long n, unsigned seed, unsigned long long *mismatches);
int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);
int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s);

}  // namespace sb
