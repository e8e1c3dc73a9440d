// sb_fs2.h — the feather / no-blend frame kernel, second generation (kernels_fstream2.cu).
//
// One launch per frame: cv::remap (fixed-point bilinear, BORDER_REFLECT) + ExposureCompensator::apply +
// convertTo(CV_16S) + FeatherBlender::feed x n + blend + convertTo(CV_8U)   (SURVEY.md §8a a3-a7, a16-a18;
// warpers_inl.hpp:88-99, exposure_compensate.cpp:150-153,225-246, blenders.cpp:81-112,115-155,383-424).
//
// What changed against the first streaming kernel (round 1, k_feather_stream):
//   * the sequence-constant table is 4 bytes per (camera, panorama pixel) instead of 8
//         e = fx|fy<<5 (10 bits) << 22  |  word offset of the tap in the staged source box << 5  |  byte shift << 3
//     plus, only in tiles where cameras are actually blended, one byte per pixel (the L1 distance behind the feather
//     weight, saturated where the weight saturates; Blender::NO: the mask byte);
//   * source boxes arrive by 2-D tensor TMA (cp.async.bulk.tensor, one instruction per 8 box rows), table blocks by
//     1-D bulk TMA; ONE producer warp issues them, lanes in parallel, far ahead through a byte-granular ring;
//   * a consumer thread owns FOUR horizontally adjacent panorama pixels: one 16-byte shared-memory load brings its
//     table entries, its 12 output bytes leave as three 32-bit stores (mask: one), the per-tile bookkeeping is
//     amortised over 128 pixels per warp; 8 warps finish a 32x32 tile, FS2_GROUPS such groups work on different tiles;
//   * exact integer tricks: bilinear weights pre-scaled by 64 so the result byte is byte 2 of the dot product
//     (no shifts), float <-> int through the 2^23 magic constant on the FP32 pipe instead of I2F/F2I.
#pragma once
#include <cuda.h>          // CUtensorMap (types only: the encoder entry point is resolved at run time, libcuda is not linked)

#include "sb_fused.h"

namespace sb {

#ifndef SB_CFG_FS2_GROUPS
#define SB_CFG_FS2_GROUPS 3
#endif
#ifndef SB_CFG_FS2_CTAS_PER_SM
#define SB_CFG_FS2_CTAS_PER_SM 1
#endif
#ifndef SB_CFG_FS2_PRODUCERS
#define SB_CFG_FS2_PRODUCERS 4
#endif
#ifndef SB_CFG_FS2_ROWS
#define SB_CFG_FS2_ROWS 48
#endif
#ifndef SB_CFG_FS2_RING_KB
#define SB_CFG_FS2_RING_KB 200
#endif

#ifndef SB_CFG_FS2_TMAP_CAMS
#define SB_CFG_FS2_TMAP_CAMS 12
#endif
constexpr int FS2_TMAP_CAMS = SB_CFG_FS2_TMAP_CAMS;
constexpr int FS2_W = 32, FS2_H = 32;                 // panorama tile
constexpr int FS2_NCLS = 4;                           // source-box width classes (one tensor map per camera and class)
constexpr int FS2_ROWS = SB_CFG_FS2_ROWS;             // box rows per tensor copy
constexpr int FS2_MAX_OPS = 96 / FS2_ROWS;            // <= 96 box rows
static_assert(1 + 3 + 3 * (2 + FS2_MAX_OPS) <= 16, "a tile's copy list must fit its descriptor");
constexpr int FS2_MAX_BOX_W = 256;                    // TMA box dimension limit
constexpr int FS2_ENT_BYTES = FS2_W * FS2_H * 4;      // tap entries of one (tile, camera)
constexpr int FS2_BLOCK_BYTES = FS2_ENT_BYTES + FS2_W * FS2_H;   // + the weight-index plane (fetched only where cameras blend)
constexpr int FS2_GAIN_W = FS2_W + 4;                 // ... fetched from a 16-byte aligned column: up to 3 leading columns are skipped
constexpr int FS2_GAIN_BYTES = FS2_GAIN_W * FS2_H * 4; // a tile of a camera's resized block gain map (SB_COMP_GAIN_BLOCKS)
constexpr int FS2_MAXC = 3;
constexpr int FS2_DESC_RECS = 16;                      // 16-byte records per tile descriptor: 1 + FS2_MAXC + the tile's copy list                           // cameras with weight inside one tile (more -> k_feather_fused_px1)
constexpr int FS2_GROUPS = SB_CFG_FS2_GROUPS;         // consumer groups (8 warps each) per CTA, each on its own tile
constexpr int FS2_GROUP_WARPS = 8;
constexpr int FS2_STAGES = 8 * FS2_GROUPS;            // tile entries (descriptor + barriers) in flight per CTA
constexpr int FS2_CTAS_PER_SM = SB_CFG_FS2_CTAS_PER_SM;
constexpr int FS2_RING_BYTES = SB_CFG_FS2_RING_KB * 1024;
constexpr int FS2_PRODUCERS = SB_CFG_FS2_PRODUCERS;   // producer warps, each on its own tile
constexpr int FS2_THREADS = (FS2_GROUPS * FS2_GROUP_WARPS + FS2_PRODUCERS) * 32;
static_assert(FS2_MAXC * (FS2_BLOCK_BYTES + FS2_MAX_BOX_W * FS2_ROWS * FS2_MAX_OPS + FS2_GAIN_BYTES) <= FS2_RING_BYTES, "one tile must fit the ring");

struct Fs2Cam {
    const unsigned char *blocks;   // tile-major table blocks of this camera (FS2_BLOCK_BYTES each)
    const float *gmap;             // SB_COMP_GAIN_BLOCKS: gain per warped pixel (null: the scalar gain)
    unsigned gmstep;
    float gain;
    int dx, dy;                    // warped corner in panorama coordinates (gain map lookup)
    unsigned char *mb_out;         // output planes (launch_fs2 out_mode 2): this camera's RGBX plane, rows mb_step bytes apart,
    unsigned mb_step;              // ... whose row 0 is row dy of the stacked coordinate system
    int ow, oh;                    // ... and its size
};
// one frame set of a multi-frame launch (global memory, written by the host before the launch)
struct alignas(128) Fs2Frame {
    CUtensorMap tmap[FS2_TMAP_CAMS * FS2_NCLS];
    void *out;
    uint8_t *out_mask;
    const uint8_t *src0;           // camera 0's frame (crop_app_fill)
    unsigned sstep0;
};
struct alignas(64) Fs2Args {
    CUtensorMap tmap[FS2_TMAP_CAMS * FS2_NCLS];     // source image of camera i as a 2-D byte tensor, box = cls_w[c] x FS2_ROWS
    CUtensorMap gtmap[FS2_TMAP_CAMS];               // SB_COMP_GAIN_BLOCKS: camera i's resized gain map as a 2-D float tensor, box = one tile
    int gain_tma;                                   // gain tiles arrive with the tile (else: read from global memory per pixel)
    Fs2Cam cam[SB_MAX_CAMERAS];
    const uint4 *desc;             // per tile in schedule order, CTA-major: FS2_DESC_RECS records (kernels_fstream2.cu)
    unsigned cls_w[FS2_NCLS];
    float sharpness;
    int no_blend;
    void *out;
    unsigned out_step;
    uint8_t *out_mask;
    unsigned mask_step;
    int pw, ph, n_tiles, n;
    int per_cta;                   // descriptors of CTA b start at desc + b * per_cta * FS2_DESC_RECS
    const Fs2Frame *frames;        // multi-frame launch: n_frames frame sets (tmap / out / out_mask above are then unused)
    int n_frames;                  // 0 or 1: the single frame set described by this block
    int steady;                    // Fs2Plan::steady
    int fill_on;                   // Blender::NO with crop_app_fill: an uncovered pixel takes camera 0's warped pixel (0, 0) (APP64:165-172)
    unsigned fill_tex[2];          // ... whose row-major table entry (sb_fused.h: FeatherCam) this is
    const uint8_t *src0;           // ... sampled from camera 0's frame
    unsigned sstep0;
    int debug;                     // tuning experiments (SB_FS2_DEBUG): 1 = supply side only (no pixel work)
    unsigned long long *trace;     // null, or (SB_FS2_TRACE=file) per CTA and tile {issued, wait, landed, done} timestamps
};

// one (camera, tile) of the setup pass: the source bounding box of the weighted entries
struct Fs2Box {
    int mnx, mny, mxx, mxy;        // mxx < mnx: the camera carries no weight in this tile
    int all_one;                   // every pixel of the tile carries weight exactly 1.0f (Blender::NO: mask 255)
    int pad[3];
};
// ... and what the entries pass needs back: where the box starts and its shared-memory pitch
struct Fs2Place {
    int xlo, ylo, pitch, valid;
};
struct Fs2Plan {                    // host-side result of the setup: what the compositor keeps per calibration
    unsigned cls_w[FS2_NCLS];
    int n_tiles = 0, grid = 0, per_cta = 0;
    double table_bytes = 0;         // bytes of table blocks one frame fetches (algorithmic bytes of the table stream)
    bool ok = false;
    bool gain_tma = false;          // the descriptors carry a gain-tile copy per camera slot
    bool steady = false;            // the ring plan of frames >= 1 of a multi-frame launch is cyclic (else the ring is drained between frames)
};

struct Fs2CamSetup {               // one camera as the setup sees it
    const uint2 *table;             // row-major feather table (kernels_fused.cu: k_build_feather_table)
    size_t tstep;
    int ww, wh, dx, dy;             // warped size, warped corner in panorama coordinates
    int tx0, ty0, ntx, nty;         // the panorama tiles its warped rect touches
    unsigned char *blocks;          // out: ntx * nty blocks of FS2_BLOCK_BYTES
};
int fs2_grid(int n_tiles, int sm_count);
// the whole per-calibration setup; plan->ok == false: this calibration needs k_feather_fused_px1 (too many cameras per
// tile, a source box the ring cannot stage, or a sharpness below 1/255)
int fs2_build(const Fs2CamSetup *cams, int n, int pw, int ph, float sharpness, int sm_count, bool gain_maps, DevBuf &desc_out, Fs2Plan *plan, cudaStream_t s,
              bool drop_empty = false);      // drop_empty: tiles without a camera leave the schedule (output planes)
int tmap_encode_u32(const void *base, size_t step, int w, int h, int box_w, int box_h, CUtensorMap *out);   // image of 32-bit pixels, box in pixels
int fs2_encode_gain_tmap(const float *gmap, size_t step, int w, int h, CUtensorMap *out);
// setup, per camera: bounding boxes of the tile blocks covering the camera's warped rect (row-major feather table in)
int launch_fs2_bbox(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                    float sharpness, Fs2Box *boxes, cudaStream_t s);
// setup, per camera: the tile-major blocks
int launch_fs2_entries(const uint2 *table, size_t tstep, int ww, int wh, int dx, int dy, int tx0, int ty0, int ntx, int nty,
                       const Fs2Place *place, unsigned dist_cap, unsigned char *blocks, cudaStream_t s);
// the source image of one camera as FS2_NCLS tensor maps
int fs2_encode_tmaps(const void *src, size_t sstep, int sw, int sh, const unsigned cls_w[FS2_NCLS], CUtensorMap out[FS2_NCLS]);
int launch_fs2(const Fs2Args &a, bool apply_gain, int out_mode, int grid, cudaStream_t s);   // out_mode: 0 CV_16SC3, 1 CV_8UC3, 2 per-camera RGBX planes (k_fs2)
int fs2_trace_dump();           // debugging aid (SB_FS2_TRACE): see kernels_fstream2.cu

}  // namespace sb
