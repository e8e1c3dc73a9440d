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
    const uint2 *table;
    size_t tstep;          // bytes per table row
    int ww, wh;            // warped size
    int dx, dy;            // warped corner in panorama coordinates
    float gain;
};

struct FeatherFusedArgs {
    int n;
    FeatherCam cam[SB_MAX_CAMERAS];
    const uint32_t *tile_cams;   // per 128-pixel panorama column block: bitmask of cameras with non-zero weight there
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

constexpr int SB_FEATHER_TILE_W = 128;    // panorama pixels per block column of k_feather_fused
int launch_feather_fused(const FeatherFusedArgs &a, bool apply_gain, bool out8, cudaStream_t s);
// setup: fixed-point map + distance table of one camera (dist: CV_32FC1 output of distanceTransform)
int launch_build_feather_table(const ProjParams &p, int tl_x, int tl_y, const DImage &dist, uint2 *table, size_t tstep, cudaStream_t s);
// device self-test of SharedDiv against __fdiv_rn; returns the number of mismatching quotients
int selftest_division(unsigned long long n, unsigned seed, unsigned long long *mismatches);
int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);
int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s);

}  // namespace sb
