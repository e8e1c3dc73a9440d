// sb_fused.h — argument blocks of the panorama-centric fused kernels (kernels_fused.cu).
#pragma once
#include "sb_kernels.h"

namespace sb {

constexpr int SB_MAX_CAMERAS = 12;

// one camera as seen by k_feather_fused: projector, separable trig tables, source frame, weights
struct FusedCam {
    float k_rinv[9];
    float one_minus_t2;
    const float *col_sin, *col_cos, *row_a, *row_b;
    const uint8_t *src;
    size_t sstep;
    int sw, sh;            // source size
    int ww, wh;            // warped size
    int dx, dy;            // warped corner in panorama coordinates
    float gain;
    int apply_gain;
    const float *weight;   // feather weight map (ww x wh), sequence-constant
    size_t wstep;
    int span[4];           // panorama column ranges [s0,s1) U [s2,s3) holding non-zero weights
};

struct FeatherFusedArgs {
    int n;
    FusedCam cam[SB_MAX_CAMERAS];
    const float *wsum;     // dst_weight_map_ (sequence-constant)
    size_t wsum_step;
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

int launch_feather_fused(const FeatherFusedArgs &a, int kind, bool out8, cudaStream_t s);
int launch_band_fused(const BandFusedArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);
int launch_column_nonzero(const DImage &w, int *flags, cudaStream_t s);

}  // namespace sb
