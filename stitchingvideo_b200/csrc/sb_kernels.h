// sb_kernels.h — host-callable launchers of the sm_100a kernels (one per SURVEY.md §8a row).
#pragma once
#include "sb_internal.h"

namespace sb {

// ProjectorBase (warpers.hpp:75-87) as a kernel argument
struct ProjParams {
    int kind;
    float scale;
    float k[9], rinv[9], r_kinv[9], k_rinv[9], t[3];
    float a, b;      // CompressedRectilinear / Panini parameters
};

// ---- warp (kernels_warp.cu) -------------------------------------------------------------------
// a3: buildMaps — one mapBackward per destination pixel of Rect(tl, br)
int launch_build_maps(const ProjParams &p, int tl_x, int tl_y, const DImage &xmap, const DImage &ymap, cudaStream_t s);
// a4: cv::remap 8UC1/8UC3 with CV_32FC1 maps, or with the CV_16SC2 (+ CV_16UC1) fixed-point pair of cv::convertMaps
int launch_remap(const DImage &src, const DImage &dst, const DImage &xmap, const DImage &ymap, int interp, int border,
                 const uint8_t bv[4], cudaStream_t s);

// cv::convertMaps: float maps -> fixed-point pair (map2 unused when nn)
int launch_convert_maps(const DImage &xmap, const DImage &ymap, const DImage &map1, const DImage &map2, bool nn, cudaStream_t s);

// Fused per-frame warp of the compositor: maps recomputed on the fly from separable trig tables
// (a3), fixed-point bilinear remap BORDER_REFLECT (a4), gain (a5), convertTo(CV_16S) (a7) and the
// BORDER_REFLECT padding of MultiBandBlender::feed (a10) in one pass.
struct WarpTables {          // sequence-constant, built once per calibration
    const float *col_sin;    // sinf(u/scale), cosf(u/scale) for u = tl.x .. br.x
    const float *col_cos;
    const float *row_a;      // spherical: sinf(pi - v/scale); cylindrical: v/scale
    const float *row_b;      // spherical: cosf(pi - v/scale); cylindrical: unused
};
int launch_build_warp_tables(const ProjParams &p, int tl_x, int tl_y, int w, int h, float *col_sin, float *col_cos,
                             float *row_a, float *row_b, cudaStream_t s);
// dst: padded rect (16SC3 or 8UC3).  Pixel (px, py) of dst takes warped pixel
// (reflect(px - left, warped_w), reflect(py - top, warped_h)).
int launch_warp_fused(const ProjParams &p, const WarpTables &t, const DImage &src, int warped_w, int warped_h,
                      int left, int top, float gain, bool apply_gain, const DImage &dst, cudaStream_t s);

// ---- exposure (kernels_pointwise.cu) ----------------------------------------------------------
int launch_scale_8u(const DImage &img, float gain, cudaStream_t s);                       // a5
int launch_mul_map_8u(const DImage &img, const DImage &gain_full, cudaStream_t s);        // a6 (after resize)
int launch_resize_linear_32f(const DImage &src, const DImage &dst, cudaStream_t s);       // a6 cv::resize
int launch_convert(const DImage &src, const DImage &dst, cudaStream_t s);                 // a7 8U->16S, 16S->8U, same->same
int launch_copy_make_border(const DImage &src, const DImage &dst, int top, int left, int border, cudaStream_t s);  // a10
int launch_set_zero(const DImage &img, cudaStream_t s);

// ---- once-per-calibration steps (kernels_calib.cu) -----------------------
struct OverlapPair {            // one overlap of GainCompensator::feed (exposure_compensate.cpp:99-124), already cropped
    const uint8_t *img1, *img2;     // 8UC3 rows of the overlap in image i / j
    const uint8_t *mask1, *mask2;   // 8UC1
    unsigned istep1, istep2, mstep1, mstep2;
    int w, h;
    int val1, val2;                 // masks[i].second / masks[j].second
};
// out[pair] = {count, sum1 low, sum1 high, sum2 low, sum2 high}: sums of sqrt(r^2+g^2+b^2) * 2^52 as exact integers
int launch_overlap_stats(const OverlapPair *pairs_dev, int n_pairs, int max_h, unsigned long long *out_dev, cudaStream_t s);
int launch_dilate3x3(const DImage &src, const DImage &dst, cudaStream_t s);               // cv::dilate(src, dst, Mat())
int launch_resize_linear_8u(const DImage &src, const DImage &dst, const DImage *and_mask, cudaStream_t s);   // cv::resize 8UC1 (& mask)

// ---- pyramids (kernels_pyr.cu) ----------------------------------------------------------------
int launch_pyr_down(const DImage &src, const DImage &dst, cudaStream_t s);                // A2 (8U/16S/32F, cn 1|3)
int launch_pyr_up(const DImage &src, const DImage &dst, cudaStream_t s);                  // A3 (8U/16S)
// dst = saturate(fine - pyrUp(coarse)) : one Laplacian level (blenders.cpp:485-486; 8U branch :463-464 -> 16S)
int launch_laplace_level(const DImage &fine, const DImage &coarse, const DImage &dst, cudaStream_t s);
// fine = saturate(pyrUp(coarse) + fine) : one collapse step (blenders.cpp:527-528)
int launch_collapse_level(const DImage &coarse, const DImage &fine_inout, cudaStream_t s);

// ---- blend (kernels_blend.cu) -----------------------------------------------------------------
// mask -> level-0 weights with the BORDER_CONSTANT padding (blenders.cpp:285-295)
int launch_mask_to_weight(const DImage &mask, const DImage &w0, int top, int left, cudaStream_t s);
// a13, fused with the Laplacian: dst(x+ox, y+oy) += trunc(lap * w); dst_w += w, where
// lap = fine - pyrUp(coarse) (coarse.empty(): lap = fine, the top Gaussian level).  fine may be 8UC3.
// dst_w.empty(): the weight sums are sequence-constant and already resident (compositor path).
int launch_lap_accumulate(const DImage &fine, const DImage &coarse, const DImage &w, const DImage &dst,
                          const DImage &dst_w, int ox, int oy, cudaStream_t s);
// dst_w(x+ox, y+oy) += w : the weight-sum half of a13 alone (built once per calibration by the compositor)
int launch_weight_accumulate(const DImage &w, const DImage &dst_w, int ox, int oy, cudaStream_t s);
// w /= sum(roi) with sum(roi) < eps set to 1 in place (FeatherBlender::createWeightMaps, blenders.cpp:176-183)
int launch_weight_normalize(const DImage &w, const DImage &sum, int ox, int oy, cudaStream_t s);
// a14 normalizeUsingWeightMap in place
int launch_normalize(const DImage &weight, const DImage &src, cudaStream_t s);
// a14+a15 fused: fine = saturate(pyrUp(coarse) + normalize(fine, w))
int launch_normalize_collapse(const DImage &coarse, const DImage &w, const DImage &fine_inout, cudaStream_t s);
// a16: crop, mask = w > eps (0/255), zero unmasked; src may alias nothing; out 16SC3 or 8UC3 (saturating)
int launch_finalize(const DImage &src, const DImage &weight, const DImage *src_mask, const DImage &out,
                    const DImage &out_mask, cudaStream_t s);
// a17 feather
int launch_distance_l1(const DImage &mask, const DImage &dist, DevBuf &scratch, cudaStream_t s);
int launch_weight_from_dist(const DImage &dist_inout, float sharpness, cudaStream_t s);
// dst_w.empty(): weight sum already resident.  img may be 8UC3 (widened on load: fused convertTo(CV_16S)).
int launch_feather_accumulate(const DImage &img, const DImage &w, const DImage &dst, const DImage &dst_w, int dx,
                              int dy, cudaStream_t s);
int launch_and_8u(const DImage &a, const DImage &b_inout, cudaStream_t s);   // b &= a
// a18 Blender::feed (no blending); img 16SC3 or 8UC3 (widened)
int launch_masked_copy(const DImage &img, const DImage &mask, const DImage &dst, const DImage &dst_mask, int dx,
                       int dy, cudaStream_t s);

}  // namespace sb
