/*
 * stitchb200.h — C ABI of libstitchb200.so: the B200-native (sm_100a CUDA) per-frame
 * compositing path of StitchingVideo's OpenCV 2.4.11 cv::detail pipeline
 * (warp -> exposure gain -> blend).  Plain pointers and sizes only.
 *
 * Every entry point replaces one reference interface; citations use
 *   LIB = /root/reference/stitching/OpenCV2.4.11-Stitching-64bit/OpenCV2.4.11-Stitching
 *   INC = LIB/include/opencv2/stitching
 *
 * There is no CPU fallback: a missing device or a failed launch is an error
 * (SB_ERR_CUDA), never a downgrade.  Calls on one handle must be serialised by
 * the caller (the reference objects are stateful); distinct
 * handles are independent and each owns a CUDA stream.
 *
 * Images are `sb_image` PODs — the cv::Mat / gpu::GpuMat overload pair of
 * INC/detail/warpers.hpp:385-414 collapsed into one struct: `device < 0` means
 * `data` is a host pointer (staged through the handle's device buffers),
 * `device >= 0` means a CUDA device pointer on that ordinal (zero copy).
 */
#ifndef STITCHB200_H
#define STITCHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- type / flag codes: numerically identical to OpenCV 2.4.11 ---- */
enum { SB_8U = 0, SB_16U = 2, SB_16S = 3, SB_32F = 5,
       SB_8UC1 = 0, SB_8UC3 = 16, SB_16SC1 = 3, SB_16SC3 = 19, SB_32FC1 = 5,
       SB_16UC1 = 2, SB_16SC2 = 11 };   /* the last two: fixed-point remap maps (cv::convertMaps) only */
enum { SB_INTER_NEAREST = 0, SB_INTER_LINEAR = 1 };
enum { SB_BORDER_CONSTANT = 0, SB_BORDER_REPLICATE = 1, SB_BORDER_REFLECT = 2,
       SB_BORDER_WRAP = 3, SB_BORDER_REFLECT_101 = 4 };

/* status codes mirror the cv::Exception codes the reference throws (SURVEY.md §8b "Errors") */
enum { SB_OK = 0,
       SB_ERR_NO_MEM = -4,      /* CV_StsNoMem */
       SB_ERR_BAD_ARG = -5,     /* CV_StsBadArg: unsupported factory type (blenders.cpp:60, exposure_compensate.cpp:58) */
       SB_ERR_ASSERT = -215,    /* CV_StsAssert: CV_Assert on types/sizes (blenders.cpp:83-84,125-126,238-239; warpers.cpp:52-54) */
       SB_ERR_NOT_IMPL = -213,  /* CV_StsNotImplemented */
       SB_ERR_CUDA = -217 };    /* CV_GpuApiCallError */

typedef struct sb_image {
    void  *data;
    int    rows, cols, type;
    size_t step;        /* bytes per row */
    int    device;      /* -1: host memory; >= 0: CUDA device ordinal */
} sb_image;
typedef struct sb_point { int x, y; } sb_point;
typedef struct sb_size  { int width, height; } sb_size;
typedef struct sb_rect  { int x, y, width, height; } sb_rect;

/* thread-local description of the last failure on this thread */
const char *sb_last_error(void);
const char *sb_version(void);
/* number of kernels this library has launched since load (all handles); bench.py's gpu_launches */
uint64_t sb_kernel_launch_count(void);
int sb_device_count(void);
/* Device self-test: the shared-divisor IEEE division used inside the fused kernels against
 * div.rn.f32 on n pseudo-random operand pairs; *mismatches must come back 0. */
int sb_selftest_division(int device, unsigned long long n, unsigned seed, unsigned long long *mismatches);
/* page-locked host memory for the pipelined compositor path (host<->device copies overlap compute) */
int  sb_host_alloc(void **ptr, size_t bytes);
void sb_host_free(void *ptr);

/* =====================================================================================
 * RotationWarper — INC/detail/warpers.hpp:53-72 (interface), :102-125 (RotationWarperBase),
 * factories INC/warpers.hpp:50-167.  kind selects the projector.
 * ===================================================================================== */
enum { SB_WARP_PLANE = 0,        /* PlaneWarper      INC/detail/warpers.hpp:129-147 */
       SB_WARP_CYLINDRICAL = 1,  /* CylindricalWarper INC/detail/warpers.hpp:177-187 */
       SB_WARP_SPHERICAL = 2,    /* SphericalWarper   INC/detail/warpers.hpp:159-166 */
       /* The remaining projectors of INC/detail/warpers.hpp:190-503.  Their maps use
        * atan2f/asinf/tanf/logf/sinhf... of the host libm, exactly like the reference's host code, and are
        * built on the HOST once per calibration (buildMaps) and uploaded; the per-frame remap is the same CUDA
        * kernel.  Object API only (not the fused compositor). */
       SB_WARP_FISHEYE = 3, SB_WARP_STEREOGRAPHIC = 4,
       SB_WARP_COMPRESSED_RECTILINEAR = 5, SB_WARP_COMPRESSED_RECTILINEAR_PORTRAIT = 6,   /* parameters a, b */
       SB_WARP_PANINI = 7, SB_WARP_PANINI_PORTRAIT = 8,                                   /* parameters a, b */
       SB_WARP_MERCATOR = 9, SB_WARP_TRANSVERSE_MERCATOR = 10,
       SB_WARP_SPHERICAL_PORTRAIT = 11, SB_WARP_CYLINDRICAL_PORTRAIT = 12, SB_WARP_PLANE_PORTRAIT = 13 };
typedef struct sb_warper sb_warper;

/* WarperCreator::create(scale) (INC/warpers.hpp:50-83) */
int   sb_warper_create(int kind, float scale, int device, sb_warper **out);
void  sb_warper_destroy(sb_warper *w);
float sb_warper_get_scale(const sb_warper *w);                       /* warpers.hpp:118 */
int   sb_warper_set_scale(sb_warper *w, float scale);                /* warpers.hpp:119 */
/* PlaneWarper's T overloads (warpers.cpp:81-137): translation used by subsequent calls (default 0) */
int   sb_warper_set_translation(sb_warper *w, const float T[3]);
/* the (A, B) constructor arguments of the CompressedRectilinear / Panini warpers (warpers.hpp:227-299; default 1, 1) */
int   sb_warper_set_ab(sb_warper *w, float a, float b);

/* RotationWarper::warpPoint (warpers_inl.hpp:52-59).  K, R: row-major 3x3 float32 (warpers.cpp:52-54) */
int sb_warper_warp_point(sb_warper *w, const float pt[2], const float K[9], const float R[9], float uv[2]);
/* RotationWarper::warpRoi (warpers_inl.hpp:131-139): Rect(tl, br + 1) */
int sb_warper_warp_roi(sb_warper *w, sb_size src_size, const float K[9], const float R[9], sb_rect *roi);
/* RotationWarper::buildMaps (warpers_inl.hpp:62-85).  Returns Rect(tl, br) in *roi (width = br.x - tl.x);
 * maps are (roi.height+1) x (roi.width+1) CV_32FC1.  The maps stay cached in the handle for
 * sb_warper_remap.  xmap/ymap may be NULL (cache only); if non-NULL with data == NULL they are
 * pointed at the handle-owned device maps (valid until the next build on this handle). */
int sb_warper_build_maps(sb_warper *w, sb_size src_size, const float K[9], const float R[9],
                         sb_image *xmap, sb_image *ymap, sb_rect *roi);
/* RotationWarper::warp (warpers_inl.hpp:88-99): buildMaps + cv::remap.  dst must be
 * (roi.height+1) x (roi.width+1) of src's type, or have data == NULL to receive a handle-owned
 * device image (valid until the next warp/remap on this handle).  *tl = dst_roi.tl(). */
int sb_warper_warp(sb_warper *w, const sb_image *src, const float K[9], const float R[9],
                   int interp_mode, int border_mode, sb_image *dst, sb_point *tl);
/* The cached-map video path of the app (APP64:188-198 caches xmap1/ymap1, APP64:752 remaps every
 * frame): cv::remap(src, dst, cached xmap, cached ymap, interp, border). */
int sb_warper_remap(sb_warper *w, const sb_image *src, int interp_mode, int border_mode, sb_image *dst);
/* RotationWarper::warpBackward (warpers_inl.hpp:102-128) */
int sb_warper_warp_backward(sb_warper *w, const sb_image *src, const float K[9], const float R[9],
                            int interp_mode, int border_mode, sb_size dst_size, sb_image *dst);

/* cv::remap itself (OpenCV 2.4.11 imgproc; call sites warpers_inl.hpp:96,127, APP64:741,752): 8UC1/8UC3,
 * INTER_LINEAR (fixed-point INTER_TAB_SIZE=32) or INTER_NEAREST.  Maps: CV_32FC1 x / y, or the fixed-point pair the
 * app's video front end uses (initUndistortRectifyMap(..., CV_16SC2, ...), APP64:201-238): xmap CV_16SC2 integer
 * coordinates + ymap CV_16UC1 fractions (ymap->data may be NULL for INTER_NEAREST). */
int sb_remap(const sb_image *src, sb_image *dst, const sb_image *xmap, const sb_image *ymap,
             int interp_mode, int border_mode, const uint8_t border_value[4], int device);

/* cv::convertMaps(xmap, ymap, map1, map2, CV_16SC2, nn_interpolation): float maps -> fixed-point pair, i.e. the map
 * conversion cv::remap repeats on every call, done once.  map2 is ignored when nn_interpolation. */
int sb_convert_maps(const sb_image *xmap, const sb_image *ymap, sb_image *map1, sb_image *map2, int nn_interpolation, int device);

/* =====================================================================================
 * ExposureCompensator — INC/detail/exposure_compensate.hpp:51-101.
 * feed() (gain estimation) is calibration, once per sequence (north_star): its result is either
 * handed in (sb_comp_set_gains / sb_comp_set_gain_maps) or estimated by sb_comp_feed, which reduces
 * the overlap statistics on the device and solves the normal equations on the host.
 * ===================================================================================== */
enum { SB_COMP_NO = 0, SB_COMP_GAIN = 1, SB_COMP_GAIN_BLOCKS = 2 };   /* exposure_compensate.hpp:56 */
typedef struct sb_comp sb_comp;
/* ExposureCompensator::createDefault (exposure_compensate.cpp:51-61) */
int  sb_comp_create(int kind, int device, sb_comp **out);
void sb_comp_destroy(sb_comp *c);
/* GainCompensator::gains_ (exposure_compensate.cpp:144, :156-162) */
int  sb_comp_set_gains(sb_comp *c, const double *gains, int n);
int  sb_comp_get_gains(const sb_comp *c, double *gains, int n);
/* BlocksGainCompensator::gain_maps_ (exposure_compensate.cpp:203-221): n host CV_32FC1 maps */
int  sb_comp_set_gain_maps(sb_comp *c, const sb_image *maps, int n);
/* ExposureCompensator::feed(corners, images, masks) (exposure_compensate.cpp:64-71): GainCompensator::feed (:76-147) or
 * BlocksGainCompensator::feed (:165-222).  images CV_8UC3, masks CV_8UC1 (level 255), host or device.  The pair
 * statistics N(i,j) (exact) and I(i,j) = sum sqrt(r^2+g^2+b^2) / N are reduced on the device; the sums are accumulated
 * as exact 128-bit integers (order independent, rounded once), where the reference adds doubles in scan order — gains
 * agree with the reference to ~1e-12 relative, not bit for bit.  A = n x n is dense as in the reference (blocks: n =
 * total block count; the LU is the reference's O(n^3)). */
int  sb_comp_feed(sb_comp *c, const sb_point *corners, const sb_image *images, const sb_image *masks, int n);
/* The host half of feed on its own (exposure_compensate.cpp:128-144): gains from the pair statistics.  Pairs (i <= j)
 * each once, N = max(1, overlap count), Iij / Iji = mean intensity of image i / j inside the overlap.  Up to 512 unknowns
 * it is the reference's dense LU (cv::solve); larger systems (the block compensator: one unknown per block) are solved
 * sparse by preconditioned conjugate gradients to 1e-15 relative residual instead of a dense O(n^3) factorisation.
 * Needs no device. */
int  sb_gain_solve(int n, int n_pairs, const int *pi, const int *pj, const double *N, const double *Iij, const double *Iji, double *gains);
/* BlocksGainCompensator(bl_width = 32, bl_height = 32) ctor arguments (exposure_compensate.hpp:92) */
int  sb_comp_set_block_size(sb_comp *c, int bl_width, int bl_height);
/* number of gains / gain maps the compensator holds after feed or set */
int  sb_comp_num_gains(const sb_comp *c);
/* BlocksGainCompensator::gain_maps_[index] as estimated by sb_comp_feed: size, then a copy into a host CV_32FC1 image */
int  sb_comp_gain_map_size(const sb_comp *c, int index, sb_size *size);
int  sb_comp_get_gain_map(const sb_comp *c, int index, sb_image *map);
/* ExposureCompensator::apply (exposure_compensate.cpp:150-153, 225-246): in place on 8UC3 */
int  sb_comp_apply(sb_comp *c, int index, sb_point corner, sb_image *image, const sb_image *mask);

/* The seam-mask refinement every compose loop of the reference runs per camera (stitcher.cpp:291-294; SAMPLE:731-735):
 *     dilate(masks_warped[i], dilated, Mat());  resize(dilated, seam_mask, mask_warped.size());  out = seam_mask & mask_warped
 * seam_mask: CV_8UC1 at seam-estimation scale; mask_warped, out: CV_8UC1 at compose scale.  Bit-exact. */
int sb_refine_seam_mask(const sb_image *seam_mask, const sb_image *mask_warped, sb_image *out, int device);
/* its two OpenCV primitives on CV_8UC1: cv::dilate(src, dst, Mat()) and cv::resize(src, dst, dst.size(), 0, 0, INTER_LINEAR) */
int sb_dilate3x3(const sb_image *src, sb_image *dst, int device);
int sb_resize_linear_8u(const sb_image *src, sb_image *dst, int device);

/* =====================================================================================
 * Blender / FeatherBlender / MultiBandBlender — INC/detail/blenders.hpp:53-117, blenders.cpp
 * ===================================================================================== */
enum { SB_BLEND_NO = 0, SB_BLEND_FEATHER = 1, SB_BLEND_MULTI_BAND = 2 };   /* blenders.hpp:58 */
typedef struct sb_blender sb_blender;
/* Blender::createDefault (blenders.cpp:52-62) + ctor parameters: FeatherBlender(sharpness = 0.02f)
 * (blenders.hpp:75), MultiBandBlender(try_gpu, num_bands = 5, weight_type = CV_32F) (blenders.hpp:99). */
int  sb_blender_create(int kind, int num_bands, int weight_type, float sharpness, int device, sb_blender **out);
void sb_blender_destroy(sb_blender *b);
int  sb_blender_num_bands(const sb_blender *b);                      /* blenders.hpp:101 */
int  sb_blender_set_num_bands(sb_blender *b, int n);                 /* blenders.hpp:102 */
float sb_blender_sharpness(const sb_blender *b);                     /* blenders.hpp:77 */
int  sb_blender_set_sharpness(sb_blender *b, float s);               /* blenders.hpp:78 */
/* FeatherBlender::createWeightMaps(masks, corners, weight_maps) (blenders.hpp:80-81, blenders.cpp:158-186): the feather
 * weights of a fixed set of images normalised by their sum over the result ROI ("final image can be obtained by simple
 * weighting of the source images").  masks CV_8UC1; weight_maps: n caller-allocated CV_32FC1 images of the masks' sizes
 * (host or device); *dst_roi = resultRoi(corners, masks).  Feather blenders only.  Bit-exact. */
int  sb_blender_create_weight_maps(sb_blender *b, const sb_image *masks, const sb_point *corners, int n, sb_image *weight_maps, sb_rect *dst_roi);
/* Blender::prepare(corners, sizes) (blenders.cpp:65-68) / virtual prepare(Rect) (:71-78,115-120,203-233) */
int  sb_blender_prepare(sb_blender *b, const sb_point *corners, const sb_size *sizes, int n);
int  sb_blender_prepare_rect(sb_blender *b, sb_rect dst_roi);
/* Blender::feed (blenders.cpp:81-102, 123-147, 236-356): img CV_16SC3 (multi-band also CV_8UC3),
 * mask CV_8U.  Asynchronous on the handle's stream when the inputs are device images. */
int  sb_blender_feed(sb_blender *b, const sb_image *img, const sb_image *mask, sb_point tl);
/* size of the image blend() will return (dst_roi_final_ for multi-band) */
int  sb_blender_result_size(const sb_blender *b, sb_size *size);
/* Blender::blend (blenders.cpp:105-112, 150-155, 359-377): dst CV_16SC3, dst_mask CV_8U; either may
 * have data == NULL to receive handle-owned device images.  Synchronises the stream.  Like the
 * reference (which hands its buffers to the caller), prepare must be called again afterwards. */
int  sb_blender_blend(sb_blender *b, sb_image *dst, sb_image *dst_mask);

/* blenders.hpp:122-133 auxiliary functions */
int sb_normalize_using_weight_map(const sb_image *weight, sb_image *src, int device);          /* blenders.cpp:383-424 */
int sb_create_weight_map(const sb_image *mask, float sharpness, sb_image *weight, int device); /* blenders.cpp:427-432 */
/* createLaplacePyr (blenders.cpp:435-489): pyr[0..num_levels] caller-allocated CV_16SC3, sizes halving */
int sb_create_laplace_pyr(const sb_image *img, int num_levels, sb_image *pyr, int device);
int sb_restore_image_from_laplace_pyr(sb_image *pyr, int num_images, int device);              /* blenders.cpp:520-530 */

/* =====================================================================================
 * Compositor — the per-frame loop of Stitcher::composePanorama (LIB/src/stitcher.cpp:221-313) and
 * of the live app's StitchingAll (APP64:724-770) with calibration fixed: everything
 * frame-independent (maps, seam masks, weight pyramids, weight sums, gains) is built once at
 * create time and kept resident in HBM; compose() runs warp -> gain -> convertTo(16S) ->
 * Blender::feed x n -> Blender::blend -> convertTo(8U) for one frame set.
 * ===================================================================================== */
typedef struct sb_compositor sb_compositor;
typedef struct sb_compositor_config {
    int      n_cameras;
    sb_size  src_size;            /* all cameras share one frame size */
    int      warper_kind;         /* SB_WARP_* */
    float    warper_scale;
    const float *K;               /* n x 9 row-major float32 */
    const float *R;               /* n x 9 */
    int      blender_kind;        /* SB_BLEND_* */
    int      num_bands;           /* multi-band */
    int      weight_type;         /* SB_32F or SB_16S (multi-band) */
    float    sharpness;           /* feather */
    int      comp_kind;           /* SB_COMP_NO, SB_COMP_GAIN or SB_COMP_GAIN_BLOCKS */
    const double *gains;          /* n gains (SB_COMP_GAIN) */
    /* optional n seam masks in warped coordinates (host CV_8UC1, size of each camera's warped
     * image), ANDed with the warped all-255 mask exactly as stitcher.cpp:278-294; NULL = none */
    const sb_image *seam_masks;
    int      output_type;         /* SB_8UC3 (result.convertTo(CV_8U), stitcher.cpp:313) or SB_16SC3 */
    /* SB_COMP_GAIN_BLOCKS: n host CV_32FC1 block gain maps (BlocksGainCompensator::gain_maps_,
     * exposure_compensate.cpp:203-221), resized to each warped image with INTER_LINEAR once, as apply() does
     * on every call (:225-246; the live app's BlockApply, APP64:310-331).  Fused paths only. */
    const sb_image *gain_maps;
    /* ---- round-2 additions; all zero = off (a zero-initialised struct behaves as before) ---- */
    /* CompressedRectilinear* / Panini* projector parameters (detail/warpers.hpp:300-420); 0, 0 means the default 1, 1 */
    float    warper_a, warper_b;
    /* Fisheye-undistort stage in front of the warp, the live app's first remap (APP64:201-238, 736-745):
     * n pairs of host maps as initUndistortRectifyMap(..., CV_16SC2, map1, map2) makes them - undistort_map1[i] CV_16SC2,
     * undistort_map2[i] CV_16UC1, both of the frame size.  Per frame and camera: remap(frame, INTER_LINEAR,
     * BORDER_CONSTANT 0) with cv::remap's fixed-point arithmetic, THEN the warp of the 8-bit result - two roundings, exactly
     * as the reference (a single composed map would not be bit-exact).  NULL = no stage. */
    const sb_image *undistort_map1;
    const sb_image *undistort_map2;
    /* Crop margins of the live app's composite (APP64:47, 150-177, 702): the panorama handed back is
     *   (width - crop_left - crop_right) x int(height * (1 - crop_up - crop_down)),
     * its pixel (x, y) = composite(x + crop_left, y + yy), yy = int(rows / (1 - crop_up - crop_down) * crop_up) in float
     * arithmetic as APP64:153.  Cropped-away tiles are never computed.  SB_BLEND_NO and SB_BLEND_FEATHER.
     * crop_app_fill != 0 reproduces feedSizeRemap's unconditional gather (APP64:165-172, SB_BLEND_NO only): a pixel no camera
     * covers takes camera 0's warped pixel (0, 0) - its look-up entries are all zero - instead of 0; the mask is unaffected. */
    float    crop_up, crop_down;          /* fractions of the height */
    int      crop_left, crop_right;       /* pixels */
    int      crop_app_fill;
} sb_compositor_config;
/* At most SB_MAX_COMPOSITOR_CAMERAS cameras per compositor (and per calibration file). */
#define SB_MAX_COMPOSITOR_CAMERAS 12

int  sb_compositor_create(const sb_compositor_config *cfg, int device, sb_compositor **out);

/* One process, several GPUs (SURVEY.md 8e throughput mode, as C++ host code): frame f of a sequence runs on devices[f % n];
 * every device gets its own compositor built from `cfg` (tables replicated), `depth` frame sets in flight per device, one
 * host thread per device inside sb_multi_run, no data-path collective.  sb_multi_run blocks until all n_frames panoramas
 * (srcs[f * n_cameras + i] -> panos[f], pano_masks[f] or NULL) have landed; results equal n_frames calls of
 * sb_compositor_compose.  sb_multi_handle gives the per-device compositor (e.g. to build an sb_batch per device). */
typedef struct sb_multi sb_multi;
int  sb_multi_create(const sb_compositor_config *cfg, int n_devices, const int *devices, int depth, sb_multi **out);
int  sb_multi_size(const sb_multi *m);
sb_compositor *sb_multi_handle(sb_multi *m, int k);
int  sb_multi_run(sb_multi *m, int n_frames, const sb_image *srcs, sb_image *panos, sb_image *pano_masks);
void sb_multi_destroy(sb_multi *m);

/* Calibration-table serialization: everything a (re)calibration hands to the per-frame path —
 * the sb_compositor_config with its K, R, gains / block gain maps and seam masks (host images) — in one checksummed
 * file, written atomically.  The reference keeps these only in process memory (APP64:334-346 PreStitchingStruct), so a
 * restart pays the 1-2 s calibration again (APP64:696-722); with the file a process (or another rank) resumes by
 * sb_calibration_load + sb_compositor_create(sb_calibration_config(cal), ...), which rebuilds the device tables in
 * milliseconds, bit-identical.  Host-only: works without a device. */
typedef struct sb_calibration sb_calibration;
int  sb_calibration_save(const sb_compositor_config *cfg, const char *path);
int  sb_calibration_load(const char *path, sb_calibration **out);
/* the loaded configuration; pointers inside stay valid until sb_calibration_free */
const sb_compositor_config *sb_calibration_config(const sb_calibration *cal);
void sb_calibration_free(sb_calibration *cal);
void sb_compositor_destroy(sb_compositor *c);
/* geometry fixed by the calibration: per-camera warped corner/size and the panorama rect */
int  sb_compositor_pano_size(const sb_compositor *c, sb_size *size);
int  sb_compositor_camera_roi(const sb_compositor *c, int index, sb_rect *roi);
/* One frame set: srcs[n] CV_8UC3 (host or device) -> pano (output_type) + pano_mask (CV_8U, may be NULL).
 * Synchronous. */
int  sb_compositor_compose(sb_compositor *c, const sb_image *srcs, sb_image *pano, sb_image *pano_mask);
/* fused != 0 (default): panorama-centric fused kernels (one pass per band / per frame);
 * fused == 0: the staged, camera-by-camera path shaped like the reference's feed/blend calls.
 * Both give identical results; the staged path is kept as a cross-check.  Values >= 10 are tuning hooks that pick a
 * kernel variant of the fused path (10: CV_16S band kernels / one-pixel-per-thread feather, 11: default fast paths,
 * 12 / 13: multi-band fast path with one launch per pyramid level / with the multi-level launches forced,
 * 14: multi-band fast path with the direct-gather form of the warp stage instead of the streaming one,
 * 15: feather / no-blend with the round-1 streaming kernel instead of its successor,
 * 16: multi-band with every level >= 1 and every band but the last in ONE launch (k_mb_coarse; 4 launches per frame - measured
 *     slower than the default of one launch per level),
 * 17: multi-band with the round-1 streaming kernel as the warp stage instead of k_fs2's output-planes mode,
 * 18: multi-band with the gather form of pyrDown instead of the tile-staged one,
 * 19: multi-band with only the levels >= 2 sharing one k_mb_coarse launch (6 launches per frame)). */
int  sb_compositor_set_fused(sb_compositor *c, int fused);
/* Which frame kernel this calibration was planned for (tests assert that the fast path really is the one that runs):
 * feather / no blending: 2 = tensor-TMA streaming kernel (k_fs2), 1 = round-1 streaming kernel, 0 = gather kernel;
 * multi-band: 2 = RGBX fast path with the streaming warp stage, 1 = RGBX fast path, 0 = CV_16S band kernels. */
int  sb_compositor_kernel_plan(const sb_compositor *c);
/* Debugging aid: with the environment variable SB_FS2_TRACE=<file> set, the feather / no-blend frame kernel records per-CTA
 * pipeline timestamps; this writes them out (scripts/fs2_trace.py reads the file).  Returns the number of launches dumped. */
int  sb_debug_fs2_trace_dump(void);
/* Pipelined form for throughput: up to `depth` frame sets in flight, each on its own stream/slot.
 * enqueue returns a slot id; wait blocks until that slot's pano has landed in the buffers given
 * to enqueue. */
int  sb_compositor_set_depth(sb_compositor *c, int depth);
int  sb_compositor_enqueue(sb_compositor *c, const sb_image *srcs, sb_image *pano, sb_image *pano_mask, int *slot);
int  sb_compositor_wait(sb_compositor *c, int slot);
/* Batches: one lap of a video pipeline's buffer ring (n_frames frame sets; srcs[f * n_cameras + i], panos[f], pano_masks[f]
 * or NULL) run with ONE host call per lap.  Results are identical to n_frames calls of sb_compositor_enqueue (the reference's
 * per-frame loop, LIB/src/stitcher.cpp:221-313, APP64:724-770).  How a lap runs (sb_batch_mode):
 *   1 = one persistent launch of the frame kernel walks all the lap's frame sets - feather / no blending with every source and
 *       every panorama (and mask) a device image: no kernel boundary, ramp-up or tail between frame sets;
 *   0 = the frame sets are enqueued back to back on the slots' streams (frame f on slot f % depth): host buffers, multi-band;
 *   2 = as 0 but recorded as a CUDA graph (environment SB_BATCH_GRAPH=1; slower on B200, kept for comparison).
 * sb_batch_launch is asynchronous (stream order after earlier launches of any batch of the handle); do not mix it with
 * enqueue/wait while a batch is in flight.  panos[f].data == NULL: the panorama stays in the slot's device buffer (lent),
 * as with enqueue (mode 0 / 2). */
typedef struct sb_batch sb_batch;
int  sb_compositor_batch_create(sb_compositor *c, int n_frames, const sb_image *srcs, sb_image *panos, sb_image *pano_masks, sb_batch **batch);
int  sb_batch_launch(sb_batch *b);
int  sb_batch_wait(sb_batch *b);
int  sb_batch_last_gpu_ms(sb_batch *b, float *ms);       /* device time of the last launch (CUDA events around the graph) */
int  sb_batch_frames(const sb_batch *b);
int  sb_batch_mode(const sb_batch *b);
void sb_batch_destroy(sb_batch *b);
/* device-resident timing of the last compose on a slot, ms (CUDA events on the slot's stream) */
int  sb_compositor_last_gpu_ms(sb_compositor *c, int slot, float *ms);
/* Device-side timing of a region spanning every slot: mark(0) before the first enqueue, mark(1)
 * after the last; marked_ms waits for mark 1 and returns the elapsed milliseconds between them. */
int  sb_compositor_mark(sb_compositor *c, int which);
int  sb_compositor_marked_ms(sb_compositor *c, float *ms);
/* Measurement hook for bench.py's roofline: runs one frame on slot 0 with every kernel bracketed by
 * CUDA events on its launching stream and writes a JSON array of
 * {"name", "ms", "bytes" (algorithmic bytes of that launch)} into buf. */
int  sb_compositor_profile_frame(sb_compositor *c, const sb_image *srcs, char *buf, size_t cap);

/* =====================================================================================
 * Latency ("strip") mode — SURVEY.md §8e, BASELINE.json configs[4]: ONE very wide panorama cut into
 * `world` column strips (boundaries are multiples of 2^num_bands of the padded panorama), one rank per
 * strip.  Multi-band path.  The reference has no counterpart (it bounds the work per image with the
 * padded sub-rectangle of blenders.cpp:242-264); results are bit-identical to the unsplit panorama.
 * Three ways for a strip to get the columns next to its boundaries, all bit-identical (measured on 8 B200, 16384-wide
 * panorama: one GPU 0.57 ms): RECOMPUTE them locally (set_strip_halo(1); no communication, 0.17 ms - the default of
 * strips.py and bench.py), PEER writes into the neighbour's memory with flags (strip_peer_*; no host in the loop,
 * 0.41 ms), or the step-by-step EXCHANGE below over any transport (NCCL send/recv in strips.py: 1.3-2.3 ms, slower than
 * one GPU - 11 tiny host-driven exchanges per frame cost latency, not bandwidth).
 * One frame on every rank, in lock step, exchange form:
 *     strip_warp(srcs)
 *     for l = 0 .. num_bands:   exchange(SB_HALO_GAUSS, l);  if l < num_bands: strip_down(l)
 *     for l = num_bands .. 0:   strip_band(l);               if l >= 1: exchange(SB_HALO_RESTORED, l)
 *     strip_result(...)
 * exchange(what, l) = for both sides: strip_pack into a device buffer, send it to that neighbour, receive the
 * neighbour's buffer, strip_unpack.  Everything is asynchronous on the handle's stream (set_stream lets it
 * share the stream the collective library is ordered with); strip_result synchronises.
 * ===================================================================================== */
enum { SB_HALO_GAUSS = 0,      /* 2 columns of every camera's Gaussian level (pyrDown / Laplacian pyrUp taps) */
       SB_HALO_RESTORED = 1 }; /* 1 column of the restored band (collapse pyrUp taps), levels 1..num_bands */
enum { SB_SIDE_LEFT = 0, SB_SIDE_RIGHT = 1 };
int  sb_compositor_num_bands(const sb_compositor *c);      /* effective number of bands (blenders.cpp:205-209) */
int  sb_compositor_set_strip(sb_compositor *c, int rank, int world);
/* recompute != 0: no exchange at all — every rank recomputes the halo columns itself (about 94 level-0 columns per
 * side for 5 bands, SURVEY.md §8e "alternative with zero comms"); strip_compose then runs a whole frame. */
int  sb_compositor_set_strip_halo(sb_compositor *c, int recompute);
int  sb_compositor_strip_compose(sb_compositor *c, const sb_image *srcs);
/* columns [*x0, *x1) of the final panorama that strip `rank` of `world` produces */
int  sb_compositor_strip_range(const sb_compositor *c, int rank, int world, int *x0, int *x1);
/* run slot 0 on the caller's CUDA stream (a cudaStream_t) */
int  sb_compositor_set_stream(sb_compositor *c, void *cuda_stream);
int  sb_compositor_strip_halo_bytes(sb_compositor *c, int what, int level, int side, size_t *send_bytes, size_t *recv_bytes);
int  sb_compositor_strip_pack(sb_compositor *c, int what, int level, int side, void *device_buf);
int  sb_compositor_strip_unpack(sb_compositor *c, int what, int level, int side, const void *device_buf);
int  sb_compositor_strip_warp(sb_compositor *c, const sb_image *srcs);
int  sb_compositor_strip_down(sb_compositor *c, int level);
int  sb_compositor_strip_band(sb_compositor *c, int level);
/* this rank's columns of the panorama (and mask); data == NULL lends a view of the device buffer */
int  sb_compositor_strip_result(sb_compositor *c, sb_image *strip, sb_image *strip_mask);
/* Halo exchange WITHOUT the host in the loop (exchange halo mode): every rank keeps one receive area per side in its own HBM;
 * the neighbour maps it (CUDA IPC handle from another process, the plain pointer inside one process) and its kernels write
 * their edge columns straight into it over NVLink, then store the step's sequence number into the area's flag word; the
 * receiving rank's kernel waits for that number and moves the columns into place.  Setup once per calibration:
 *     peer_export(side) on every rank -> exchange the 64-byte handles -> peer_connect(side, handle of that neighbour's
 *     OPPOSITE side);   then per frame ONE call: strip_frame_peer(srcs) (+ strip_result).
 * The reference has no counterpart. */
int  sb_compositor_strip_peer_export(sb_compositor *c, int side, void *ipc_handle_64, void **local_ptr, size_t *bytes);
int  sb_compositor_strip_peer_connect(sb_compositor *c, int side, const void *ipc_handle_64, void *same_process_ptr);
int  sb_compositor_strip_frame_peer(sb_compositor *c, const sb_image *srcs);
/* the two halves of one exchange step (for a process that plays several ranks on one device: all pushes of a step first) */
int  sb_compositor_strip_peer_push(sb_compositor *c, int what, int level);
int  sb_compositor_strip_peer_pull(sb_compositor *c, int what, int level);

#ifdef __cplusplus
}
#endif
#endif
