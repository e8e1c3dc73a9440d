/*
 * stitch_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-frame compositing path of the reference
 * (wangzjpku/StitchingVideo = OpenCV 2.4.11 cv::detail stitching sources) plus
 * the OpenCV 2.4.11 core/imgproc primitive semantics that path relies on
 * (those sources are NOT in /root/reference, only the
 * prebuilt opencv_*2411.dll are).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (stitchingvideo_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md §4).  The restatement is pinned by (a) live cross-checks against
 * cv2 4.13 (same arithmetic) in tests/test_oracle_vs_cv2.py,
 * (b) fixtures generated from cv2 by tests/golden/make_golden.py, and
 * (c) oracle/_ref: the reference's own blenders.cpp / warpers.cpp /
 * exposure_compensate.cpp / util.cpp compiled where they lie against a
 * minimal OpenCV shim (see oracle/ref_shim/README.md).
 *
 * Reference path shorthand:  LIB = /root/reference/stitching/
 *   OpenCV2.4.11-Stitching-64bit/OpenCV2.4.11-Stitching
 */
#ifndef STITCH_ORACLE_H
#define STITCH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* OpenCV type codes: CV_MAKETYPE(depth, cn) = depth + ((cn-1) << 3) */
enum {
    SO_8U = 0, SO_16S = 3, SO_32F = 5,
    SO_8UC1 = 0, SO_8UC3 = 16, SO_16SC1 = 3, SO_16SC3 = 19, SO_32FC1 = 5,
    SO_16UC1 = 2, SO_16SC2 = 11   /* fixed-point remap maps (cv::convertMaps) */
};
/* OpenCV interpolation / border codes (imgproc.hpp) */
enum { SO_INTER_NEAREST = 0, SO_INTER_LINEAR = 1 };
enum {
    SO_BORDER_CONSTANT = 0, SO_BORDER_REPLICATE = 1, SO_BORDER_REFLECT = 2,
    SO_BORDER_WRAP = 3, SO_BORDER_REFLECT_101 = 4
};
/* projector kinds (warpers.hpp) */
enum { SO_WARP_PLANE = 0, SO_WARP_CYLINDRICAL = 1, SO_WARP_SPHERICAL = 2 };
/* Blender::{NO,FEATHER,MULTI_BAND} (blenders.hpp:58) */
enum { SO_BLEND_NO = 0, SO_BLEND_FEATHER = 1, SO_BLEND_MULTI_BAND = 2 };

typedef struct so_mat {
    void  *data;
    int    rows, cols, type;
    size_t step;              /* bytes per row */
} so_mat;

typedef struct so_projector {   /* ProjectorBase, warpers.hpp:75-87 */
    int   kind;
    float scale;
    float k[9], rinv[9], r_kinv[9], k_rinv[9], t[3];
} so_projector;

/* ---- scalar helpers (Appendix A) ---- */
int   so_cvround(float v);                /* cvtss2si: half-even, INT_MIN on overflow/NaN */
short so_trunc_short(float v);            /* static_cast<short>(float) on x86-64: cvttss2si, low 16 bits */
int   so_border_interpolate(int p, int len, int border);
float so_sinf(float x);                   /* portable restatement of glibc 2.39 sinf (no-FMA double arithmetic) */
float so_cosf(float x);

/* ---- warpers: warpers.cpp:50-78,171-212; warpers_inl.hpp:52-300 ---- */
/* 3x3 CV_32F cv::invert / cv::gemm as OpenCV 2.4.11 computes them (used by setCameraParams, warpers.cpp:61-74) */
void so_inv3x3_f32(const float S[9], float D[9]);
void so_mul3x3_f32(const float A[9], const float B[9], float D[9]);
void so_projector_set(so_projector *p, int kind, float scale, const float K[9], const float R[9], const float T[3]);
void so_map_forward(const so_projector *p, float x, float y, float *u, float *v);
void so_map_backward(const so_projector *p, float u, float v, float *x, float *y);
void so_detect_result_roi(const so_projector *p, int src_w, int src_h, int tl[2], int br[2]);
/* xmap/ymap must be (br.y-tl.y+1) x (br.x-tl.x+1) SO_32FC1 */
void so_build_maps(const so_projector *p, const int tl[2], const int br[2], so_mat *xmap, so_mat *ymap);

/* ---- imgproc / core primitives (Appendix A1-A4) ---- */
int  so_remap(const so_mat *src, so_mat *dst, const so_mat *xmap, const so_mat *ymap,
              int interp, int border, const uint8_t border_value[4]);
int  so_copy_make_border(const so_mat *src, so_mat *dst, int top, int bottom, int left, int right, int border);
int  so_pyr_down(const so_mat *src, so_mat *dst);   /* dst = ((cols+1)/2, (rows+1)/2); 8U/16S/32F, cn 1|3 */
int  so_pyr_up(const so_mat *src, so_mat *dst);     /* dst = exactly 2x; 8U/16S cn 1|3 */
int  so_add_16s(const so_mat *a, const so_mat *b, so_mat *dst);        /* saturating */
int  so_subtract_16s(const so_mat *a, const so_mat *b, so_mat *dst);   /* saturating */
int  so_subtract_8u_to_16s(const so_mat *a, const so_mat *b, so_mat *dst);
int  so_convert_8u_16s(const so_mat *src, so_mat *dst);
int  so_convert_16s_8u(const so_mat *src, so_mat *dst);                /* saturate_cast<uchar> */
int  so_scale_8u(so_mat *img, double gain);                            /* Mat::operator*=(double) on 8U */
int  so_distance_l1_3x3(const so_mat *mask, so_mat *dist);             /* distanceTransform(CV_DIST_L1, 3) -> 32F */
int  so_resize_linear_32f(const so_mat *src, so_mat *dst);             /* cv::resize INTER_LINEAR, 32FC1 */

/* ---- exposure_compensate.cpp:150-153, 225-246 ---- */
int  so_gain_apply(so_mat *image, double gain);
int  so_blocks_gain_apply(so_mat *image, const so_mat *gain_map);

/* ---- util.cpp:118-140 ---- */
void so_result_roi(const int *corners_xy, const int *sizes_wh, int n, int roi_xywh[4]);

/* ---- blenders.cpp ---- */
void so_normalize_using_weight_map(const so_mat *weight, so_mat *src);       /* :383-424 */
int  so_create_weight_map(const so_mat *mask, float sharpness, so_mat *weight); /* :427-432 */
/* createLaplacePyr (:435-489): pyr[0..num_levels] are caller-allocated 16SC3 of halving sizes */
int  so_create_laplace_pyr(const so_mat *img, int num_levels, so_mat *pyr);
int  so_restore_image_from_laplace_pyr(so_mat *pyr, int n);                  /* :520-530 */

typedef struct so_blender so_blender;
so_blender *so_blender_create(int kind, int num_bands, int weight_type, float sharpness);
void so_blender_destroy(so_blender *b);
int  so_blender_prepare(so_blender *b, const int *corners_xy, const int *sizes_wh, int n);
int  so_blender_prepare_rect(so_blender *b, int x, int y, int w, int h);
int  so_blender_feed(so_blender *b, const so_mat *img, const so_mat *mask, int tl_x, int tl_y);
/* result size: query after prepare; dst 16SC3, dst_mask 8UC1 caller-allocated */
void so_blender_result_size(const so_blender *b, int *w, int *h);
int  so_blender_num_bands_effective(const so_blender *b);
int  so_blender_blend(so_blender *b, so_mat *dst, so_mat *dst_mask);
/* test hook: float weight pyramids are the one build-dependent quantity;
 * mode 0 = scalar summation order of pyrDown_<FltCast> (default), 1 = SSE-order vertical pass */
void so_set_float_pyrdown_order(int mode);

/* version string / self-identification */
/* ---- once-per-calibration steps next to the path (so_calib.c) ---- */
int  so_solve_lu(int n, double *A, double *b);                              /* cv::solve DECOMP_LU, in place */
void so_gain_overlap_stats(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals,
                           int *N, double *I);                              /* exposure_compensate.cpp:93-126 */
int  so_gain_solve(int n, const int *N, const double *I, double *gains);    /* :128-144 */
int  so_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals, double *gains);
int  so_sep_filter3_f32(const so_mat *src, so_mat *dst, float k0, float k1);
int  so_blocks_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals,
                         int bl_width, int bl_height, so_mat *gain_maps);   /* :165-222 */
int  so_dilate3x3_8u(const so_mat *src, so_mat *dst);
int  so_resize_linear_8u(const so_mat *src, so_mat *dst);
int  so_feather_create_weight_maps(int n, const so_mat *masks, const int *corners_xy, float sharpness, so_mat *weight_maps, int roi_xywh[4]);   /* blenders.cpp:158-186 */
int  so_refine_seam_mask(const so_mat *seam_mask, const so_mat *mask_warped, so_mat *out);   /* stitcher.cpp:291-294 */

const char *so_version(void);

#ifdef __cplusplus
}
#endif
#endif
