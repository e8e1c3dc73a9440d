// sb_mb.h — argument blocks of the multi-band fast path (kernels_mb.cu).
#pragma once
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
    int cx0, cx1;          // columns of the rect to produce (strip mode; whole rect: 0, rw)
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
    int ox0, ox1;          // output columns to produce (strip mode; whole level: 0, (sw+1)/2)
};
struct MbPyrArgs {
    int n;
    MbPyrCam cam[SB_MAX_CAMERAS];
};

struct MbBandCam {
    const uint32_t *fine;    // Gaussian level l (RGBX), rect-local
    size_t fstep;
    const uint32_t *coarse;  // Gaussian level l+1 (null at the top level)
    size_t cstep;
    const void *weight;      // weight pyramid level l (float or short), sequence-constant
    size_t wstep;
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
    size_t wsum_step;
    const short4 *coarse_r;  // restored band l+1 (CV_16SC3 values in 8-byte pixels); null at the top level
    size_t coarse_r_step;
    void *out;               // restored band l (short4 pixels), or the final panorama at band 0
    size_t out_step;
    uint8_t *out_mask;
    size_t mask_step;
    int out_w, out_h;        // band 0 only: dst_roi_final_ size
    int x_begin, x_end;      // band columns to produce (strip mode; whole band: 0, lw)
};

int launch_mb_tap_table(const ProjParams &p, int tl_x, int tl_y, int ww, int wh, int left, int top, int sw, int sh,
                        uint2 *table, size_t tstep, int rw, int rh, cudaStream_t s);
int launch_mb_tile_mask(const MbBandGeom &g, bool float_weights, int lw, int lh, uint32_t *mask, cudaStream_t s);
int launch_mb_warp(const MbWarpArgs &a, bool apply_gain, int max_rw, int max_rh, cudaStream_t s);
int launch_mb_pyr_down(const MbPyrArgs &a, int max_dw, int max_dh, cudaStream_t s);
int launch_mb_band(const MbBandArgs &a, bool float_weights, bool not_top, bool final_band, bool out8, cudaStream_t s);

}  // namespace sb
