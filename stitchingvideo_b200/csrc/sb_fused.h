// sb_fused.h — argument blocks of the panorama-centric fused kernels (kernels_fused.cu).
#pragma once
#include "sb_kernels.h"

namespace sb {

constexpr int SB_MAX_CAMERAS = 12;

// Per-camera, per-warped-pixel table of the feather path (sequence-constant, 8 bytes/pixel):
// the fixed-point map entry cv::remap derives from the float maps (A1: sx, sy after >> 5 and the
// int16 clamp, 5+5 fractional bits) and the L1 distance createWeightMap turns into the weight.
//   x: sx (low 16, signed) | sy (high 16, signed)      y: fx | fy << 5 (low 16) | dist (high 16)
// dist == 0 marks a zero-weight pixel (outside the warped mask): the kernel skips it, exactly.
struct FeatherCam {
    const uint8_t *src;
    size_t sstep;
    int sw, sh;            // source size
    const uint2 *table;    // row y starts at table + y*tstep; entry of warped column x is at index x + tpad
    size_t tstep;          // bytes per table row
    int tpad;              // (dx mod 4): makes the 4 entries of a panorama-aligned pixel quad 32-byte aligned
    int ww, wh;            // warped size
    int dx, dy;            // warped corner in panorama coordinates
    float gain;
    // per panorama tile (SB_FT_W x SB_FT_H): bounding box {x0, y0, x1, y1} (inclusive) of the source
    // pixels the tile's non-zero-weight samples touch; x1 < x0 = the camera does not contribute
    const int4 *bbox;
};

constexpr int SB_FT_W = 32, SB_FT_H = 32;        // panorama pixels per block of k_feather_fused (8 x 32 threads, 4 px each)
constexpr int SB_STAGE_BYTES = 16384;            // shared-memory stage for one camera's source box (pitch * rows must fit)

struct FeatherFusedArgs {
    int n;
    FeatherCam cam[SB_MAX_CAMERAS];
    int tiles_x;
    const uint32_t *tile_cams;   // per panorama tile: bitmask of the cameras whose bbox there is non-empty
    int variant;           // 0: 4 px/thread with the source box staged in shared memory; 1: 1 px/thread, direct taps
    float sharpness;
    void *out;             // 8UC3 or 16SC3 panorama
    size_t out_step;
    uint8_t *out_mask;     // may be null
    size_t mask_step;
    int pw, ph;
};

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

// setup: per panorama tile, the source bounding box of one camera's non-zero-weight samples
int launch_feather_tile_bbox(const FeatherCam &c, int pw, int ph, int4 *bbox, cudaStream_t s);
// setup: tile_cams[tile] |= 1 << cam_index where that camera's bbox is non-empty
int launch_feather_tile_mask(const int4 *bbox, int n_tiles, int cam_index, uint32_t *tile_cams, cudaStream_t s);
int launch_feather_fused(const FeatherFusedArgs &a, bool apply_gain, bool out8, cudaStream_t s);
// setup: fixed-point map + distance table of one camera (dist: CV_32FC1 output of distanceTransform)
int launch_build_feather_table(const ProjParams &p, int tl_x, int tl_y, const DImage &dist, uint2 *table, size_t tstep, int tpad, cudaStream_t s);
// device self-test of SharedDiv against __fdiv_rn; returns the number of mismatching quotients
int selftest_division(unsigned long long n, unsigned seed, unsigned long long *mismatches);
int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);
int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s);

}  // namespace sb
