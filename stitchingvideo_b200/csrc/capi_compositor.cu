// capi_compositor.cu — the per-frame compositing loop with calibration fixed.
//
// Host-side shape: Stitcher::composePanorama's full-resolution loop (LIB/src/stitcher.cpp:221-313)
// and the live app's StitchingAll (APP64:724-770):
//   warp(img) -> compensator.apply -> convertTo(CV_16S) -> blender.feed  (x n) -> blender.blend
//   -> convertTo(CV_8U).
// Everything that depends only on the calibration is hoisted to create() and kept resident in HBM,
// exactly as the reference app hoists its maps/LUTs out of the frame loop:
//   separable trig tables per camera (replace the 8 B/px float maps), warped masks, float/int16
//   weight pyramids per camera, per-level weight sums (dst_band_weights_), feather weight maps.
// compose() then launches only the image-dependent kernels.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <string>

#include "sb_fs2.h"
#include "sb_fused.h"
#include "sb_host_projector.h"
#include "sb_mb.h"
#include "sb_kernels.h"

using namespace sb;

namespace {

struct Camera {
    ProjParams proj;
    sb_point tl;                 // corner of the warped image in panorama coordinates
    int ww = 0, wh = 0;          // warped size (br - tl + 1)
    float gain = 1.f;
    double gain_bytes = 0;       // SB_COMP_GAIN_BLOCKS: bytes of gain_full inside the camera's weighted column spans
    DevImage gain_full;          // SB_COMP_GAIN_BLOCKS: block gain map resized to the warped image (sequence-constant)
    DevImage gain_pad;           // ... with one tile of zeros all round (k_fs2 fetches gain tiles by tensor copy; they must lie inside the tensor)
    DevImage gain_rect;          // ... and laid out over the padded feed rect (BORDER_REFLECT, multi-band fast path)
    DevBuf tables;               // col_sin | col_cos | row_a | row_b
    WarpTables wt{};
    DevImage mask;               // warped (and seam-ANDed) mask, 8UC1 ww x wh
    // multi-band: padded feed rect (blenders.cpp:241-269), in dst_roi_ coordinates
    int rx = 0, ry = 0, rw = 0, rh = 0, top = 0, left = 0;
    std::vector<DevImage> w_pyr; // weight pyramid (sequence-constant)
    DevImage feather_w;          // feather weight map (sequence-constant)
    DevBuf feather_table;        // fixed-point map + distance, 8 B per warped pixel (sequence-constant)
    DevBuf mb_table;             // multi-band fast path: resolved bilinear taps per padded-rect pixel (8 B)
    size_t mb_tstep = 0;
    DevBuf mbs_tiles, mbs_rec;   // the same taps tile-major + per-tile source boxes (streaming warp stage, kernels_mb_stream.cu)
    DevBuf mbf_blocks;           // ... and as k_fs2 table blocks (4-byte entries; the warp stage on the frame kernel of the feather path)
    int mbf_dy = 0;              // first row of this camera's plane in the stacked coordinate system of that launch
    int mbs_ntx = 0, mbs_nty = 0;
    size_t feather_tstep = 0;
    DevBuf feather_tiles;        // the same table tile-major (one 8 KB block per 32x32 panorama tile) for the streaming kernel
    DevBuf feather_rec;          // per tile block: source box record
    int ftx0 = 0, fty0 = 0, fntx = 0, fnty = 0;
    DevImage xmap, ymap;         // projectors evaluated on the host (kinds beyond plane / cylindrical / spherical): the float maps, kept for the table builders
    DevImage umap1, umap2;       // fisheye-undistort stage: initUndistortRectifyMap's CV_16SC2 / CV_16UC1 pair (APP64:201-238)
    int f2tx0 = 0, f2ty0 = 0, f2ntx = 0, f2nty = 0;   // k_fs2: the (cropped) output tiles this camera's warped rect touches
    DevBuf fs2_blocks;           // second-generation streaming kernel: tile-major 4-byte tap entries + weight-index plane
    // the source image as tensor maps (one per box width class), keyed on the frame buffer: video pipelines cycle through
    // a few input buffers, so after the first lap every frame is a cache hit
    struct TmapSet { const void *p; size_t step; unsigned long long stamp; CUtensorMap m[FS2_NCLS]; };
    std::vector<TmapSet> tmaps;
    // panorama(-level) column ranges that hold non-zero weights, per level: [s0,s1) U [s2,s3)
    std::vector<std::array<int, 4>> spans;
    // rect-local columns of Gaussian level l that the weighted pixels can see through the pyramid taps (two runs):
    // everything else of the padded rect (most of a seam-straddling camera's panorama-wide rect) is never produced
    std::vector<std::array<int, 4>> g_runs;
};

// pitched device buffer for the pixel formats outside the OpenCV type set (RGBX bytes, short4)
struct RawImage {
    DevBuf buf;
    size_t step = 0;
    int rows = 0, cols = 0, esz = 0;
    int create(int r, int c, int elem)
    {
        rows = r; cols = c; esz = elem;
        step = ((size_t)c * elem + 255) & ~(size_t)255;
        return buf.ensure(step * (size_t)(r > 0 ? r : 1));
    }
    double bytes() const { return (double)rows * cols * esz; }
};

// per-kernel record of one profiled frame (bench.py's roofline): CUDA events on the launching stream
struct ProfRec {
    const char *name;
    double bytes;            // algorithmic (compulsory) bytes of this launch: inputs once + outputs once
    cudaEvent_t e0, e1;
};

struct Slot {
    std::vector<ProfRec> *prof = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool busy = false;
    bool want_mask = true;                       // the caller asked for dst_mask (else it is never materialised)
    std::vector<DevImage> src;                   // staged source frames (host input)
    DevBuf src_all;                              // one staging block for a frame set that is contiguous in host memory
    DevBuf mb_sync;                              // k_mb_coarse: this slot's work / stage counters (only ever grow)
    std::vector<unsigned long long> mb_totals;   // ... and their running totals on the host
    std::vector<DevImage> undist;                // fisheye-undistort stage: the undistorted frames (what the warp reads)
    std::vector<std::vector<DevImage>> gpyr;     // per camera Gaussian pyramid of the padded warped image
    DevImage warped;                             // feather / no-blend: one warped image at a time
    DevImage warped_pad;                         // staged multi-band with block gains: the padded 8UC3 image before convertTo(16S)
    std::vector<std::vector<RawImage>> grgbx;    // fast path: per camera Gaussian pyramid as RGBX bytes
    std::vector<std::vector<CUtensorMap>> g_tmap; // ... and each level as the tensor k_mb_pyr_down_tma reads its tiles from (empty: not available)
    std::vector<RawImage> rband;                 // fast path: restored bands 1..n as short4 pixels
    std::vector<DevImage> acc;                   // dst_pyr_laplace_ (level 0 = dst_)
    DevImage acc_mask;                           // Blender::NO dst_mask_
    DevImage out, out_mask;
};

}  // namespace

struct sb_compositor {
    int device = 0;
    sb_compositor_config cfg{};
    std::vector<Camera> cams;
    sb_rect dst_roi{}, dst_roi_final{};
    int num_bands = 0;
    std::vector<DevImage> wsum;                  // dst_band_weights_ / dst_weight_map_ (sequence-constant)
    std::vector<Slot> slots;
    int next_slot = 0;
    cudaStream_t setup_stream = nullptr;
    cudaEvent_t marks[2] = {nullptr, nullptr};
    bool feather_fast = false;                   // every weighted pixel's taps fit the resolved-tap table
    float stream_sharpness = 0.02f;              // feather: FeatherBlender::sharpness_; Blender::NO: 1/255 (see setup)
    bool feather_tma = false;                    // <= SB_FTT_MAXC cameras per 32x32 tile: persistent table-streaming kernel
    int feather_variant = 1;                     // 1: k_fs2 (else the next that applies), 3: k_feather_stream, 0: k_feather_fused_px1
    DevBuf tma_desc;                             // per 32x32 panorama tile: camera slots + source boxes, schedule order, ring plan
    bool host_maps = false;                      // projector kind whose maps are built on the host (fused fast paths only)
    bool undistort = false;                      // fisheye-undistort stage in front of the warp
    sb_rect out_rect{};                          // what compose hands back, in dst_roi_final coordinates (crop margins; else all of it)
    bool crop = false;
    unsigned fill_tex[2] = {0, 0};               // crop_app_fill: camera 0's table entry of warped pixel (0, 0)
    CUtensorMap fs2_gtmap[SB_MAX_CAMERAS];       // ... the cameras' resized gain maps as tensors (SB_COMP_GAIN_BLOCKS)
    Fs2Plan fs2;                                 // second-generation streaming kernel (kernels_fstream2.cu): usable when fs2.ok
    DevBuf fs2_desc;
    unsigned long long tmap_clock = 0;
    int sm_count = 148;
    DevBuf bilin_lut;                            // 1024 x uint2 bilinear product weights (sb_device.cuh)
    int mb_variant = 1;                          // 1: RGBX fast path, 0: CV_16S band kernels
    bool mb_fast = false;                        // RGBX pyramid + tap-table path available (sources <= 4096 px)
    std::vector<DevBuf> mb_tile_mask;            // per band: per 32x8 tile bitmask of contributing cameras
    DevBuf tile_cams;                            // feather: per panorama tile, bitmask of contributing cameras
    // coarse pyramid levels / bands in one cooperative launch each: -1 = automatic (on when a single frame is in
    // flight: lowest latency; off when frames are pipelined over several slots: the small per-level launches of one
    // frame then overlap with the big kernels of the others, which is worth more than the saved launches), 0 / 1 = forced
    int mb_multilevel = -1;
    bool mbs_tables = false;                     // tile-major taps + source boxes built
    bool mbs_ok = false;                         // streaming warp stage usable for panorama columns [mbs_x0, mbs_x1)
    int mbs_x0 = 0, mbs_x1 = 0;
    bool mbs_enabled = true;                     // tuning hook (set_fused 12 keeps the gather kernel)
    DevBuf mbs_desc;                             // tile descriptors in schedule order with the ring plan
    Fs2Plan mbf;                                 // the warp stage as a k_fs2 launch over per-camera output planes (mb_fs2_setup)
    DevBuf mbf_desc;
    int mbf_pw = 0, mbf_ph = 0;
    bool mbf_enabled = true;                     // tuning hook (set_fused 17 keeps the round-1 streaming kernel)
    bool pyr_tma_enabled = true;                 // tuning hook (set_fused 18 keeps the gather form of pyrDown)
    int mbs_n_tiles = 0;
    bool fused = true;                           // panorama-centric fused kernels (default); false = staged reference-shaped path
    // latency ("strip") mode: this handle produces padded-panorama columns [strip_x0, strip_x1)
    int strip_rank = 0, strip_world = 1, strip_x0 = 0, strip_x1 = 0;
    bool strip_recompute = false;                // halo columns recomputed locally instead of exchanged
    // per level: Gaussian columns [g_lo, g_hi) and band columns [b_lo, b_hi) this rank produces (level coordinates)
    std::vector<int> g_lo, g_hi, b_lo, b_hi;
    bool external_stream = false;
    std::vector<DImage> strip_src;               // device views of the current frame's sources
    // peer-memory halo exchange: this rank's receive areas (one per side), the neighbours' as mapped here, step layout
    DevBuf peer_mine[2], peer_done;
    char *peer_theirs[2] = {nullptr, nullptr};
    bool peer_ipc[2] = {false, false};
    std::vector<unsigned long long> peer_region_mine[2], peer_region_theirs[2];   // per exchange step
    unsigned long long peer_frames = 0;
};

namespace {

int make_slot(sb_compositor *c, Slot &s)
{
    SB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    SB_CUDA(cudaEventCreate(&s.ev_start));
    SB_CUDA(cudaEventCreate(&s.ev_stop));
    const int n = c->cfg.n_cameras;
    s.src.resize(n);
    s.gpyr.resize(n);
    const int levels = c->cfg.blender_kind == SB_BLEND_MULTI_BAND ? c->num_bands : 0;
    s.acc.resize(levels + 1);
    int rows = c->dst_roi.height, cols = c->dst_roi.width;
    for (int l = 0; l <= levels; ++l) {
        SB_TRY(s.acc[l].create(rows, cols, SB_16SC3));
        rows = (rows + 1) / 2; cols = (cols + 1) / 2;
    }
    if (c->cfg.blender_kind == SB_BLEND_NO) SB_TRY(s.acc_mask.create(c->dst_roi.height, c->dst_roi.width, SB_8UC1));
    if (c->cfg.blender_kind == SB_BLEND_MULTI_BAND) {
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            s.gpyr[i].resize(levels + 1);
            int r = cam.rh, w = cam.rw;
            for (int l = 0; l <= levels; ++l) {
                SB_TRY(s.gpyr[i][l].create(r, w, SB_16SC3));
                r = (r + 1) / 2; w = (w + 1) / 2;
            }
        }
    }
    if (c->mb_fast) {
        s.grgbx.resize(n);
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            s.grgbx[i].resize(levels + 1);
            int r = cam.rh, w = cam.rw;
            for (int l = 0; l <= levels; ++l) {
                SB_TRY(s.grgbx[i][l].create(r, w, 4));
                r = (r + 1) / 2; w = (w + 1) / 2;
            }
        }
        s.g_tmap.assign(n, std::vector<CUtensorMap>(levels));     // levels 0 .. n-1 are pyrDown inputs (static buffers: encoded once)
        for (int i = 0; i < n && !s.g_tmap.empty(); ++i)
            for (int l = 0; l < levels; ++l) {
                const RawImage &g = s.grgbx[i][l];
                if (tmap_encode_u32(g.buf.p, g.step, g.cols, g.rows, MB_PT_IW, MB_PT_IH, &s.g_tmap[i][l]) != SB_OK) { s.g_tmap.clear(); break; }
            }
        s.rband.resize(levels + 1);
        for (int l = 1; l <= levels; ++l) SB_TRY(s.rband[l].create(s.acc[l].v.rows, s.acc[l].v.cols, 8));
    }
    // packed rows: the device->host copy of the panorama into a contiguous host image is one linear DMA
    SB_TRY(s.out.create(c->out_rect.height, c->out_rect.width, c->cfg.output_type));
    // (the frame kernels store 4 pixels per thread: 32-bit words of 8UC3, 64-bit words of 16SC3 - keep the rows so aligned)
    if (((size_t)c->out_rect.width * elem_size(c->cfg.output_type)) % (c->cfg.output_type == SB_8UC3 ? 4 : 8) == 0) {
        s.out.v.step = (size_t)c->out_rect.width * elem_size(c->cfg.output_type);
    }
    SB_TRY(s.out_mask.create(c->out_rect.height, c->out_rect.width, SB_8UC1));
    if (c->out_rect.width % 4 == 0) s.out_mask.v.step = (size_t)c->out_rect.width;
    if (c->undistort) {
        s.undist.resize(n);
        for (int i = 0; i < n; ++i) SB_TRY(s.undist[i].create_packed(c->cfg.src_size.height, c->cfg.src_size.width, SB_8UC3));
    }
    return SB_OK;
}

void free_slot(Slot &s, bool external = false)
{
    if (s.stream && !external) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_stop) cudaEventDestroy(s.ev_stop);
    s.stream = nullptr; s.ev_start = s.ev_stop = nullptr;
}

// Column ranges of a weight image that hold any non-zero weight, shifted by `offset` into
// panorama(-level) coordinates.  A seam-straddling camera has two runs (both panorama ends).
int weight_spans(const DImage &w, int offset, DevBuf &scratch, cudaStream_t s, std::array<int, 4> *out)
{
    SB_TRY(scratch.ensure(sizeof(int) * (size_t)w.cols));
    SB_TRY(launch_column_nonzero(w, static_cast<int *>(scratch.p), s));
    std::vector<int> flags(w.cols);
    SB_CUDA(cudaMemcpyAsync(flags.data(), scratch.p, sizeof(int) * (size_t)w.cols, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    std::vector<std::pair<int, int>> runs;
    for (int x = 0; x < w.cols;) {
        if (!flags[x]) { ++x; continue; }
        int e = x;
        while (e < w.cols && flags[e]) ++e;
        runs.push_back({x, e});
        x = e;
    }
    *out = {0, 0, 0, 0};
    if (runs.empty()) return SB_OK;
    if (runs.size() == 1) { *out = {runs[0].first + offset, runs[0].second + offset, 0, 0}; return SB_OK; }
    // more than two runs: keep the widest gap as the hole between the two spans
    size_t gap_at = 0;
    int gap = -1;
    for (size_t i = 0; i + 1 < runs.size(); ++i)
        if (runs[i + 1].first - runs[i].second > gap) { gap = runs[i + 1].first - runs[i].second; gap_at = i; }
    *out = {runs.front().first + offset, runs[gap_at].second + offset, runs[gap_at + 1].first + offset, runs.back().second + offset};
    return SB_OK;
}

// Which columns of each Gaussian level of a camera's padded rect can influence a weighted pixel?  Band l reads level l
// where the weight is non-zero (spans[l]) and level l+1 at (x >> 1) +- 1 (Laplacian pyrUp taps); level l+1 is made
// from level l columns [2a - 2, 2b + 1) (pyrDown taps).  Same recurrences as strip_plan, per run of the weight spans.
void gaussian_runs(Camera &cam, int nb)
{
    cam.g_runs.assign(nb + 1, std::array<int, 4>{0, 0, 0, 0});
    for (int run = 0; run < 2; ++run) {
        std::vector<int> lo(nb + 1, 0), hi(nb + 1, 0);
        for (int l = nb; l >= 0; --l) {
            const int rx = cam.rx >> l, w = cam.w_pyr[l].v.cols;
            int a = cam.spans[l][2 * run] - rx, b = cam.spans[l][2 * run + 1] - rx;      // band l (rect-local)
            bool any = b > a;
            if (l >= 1) {
                const int rxf = cam.rx >> (l - 1);
                const int fa = cam.spans[l - 1][2 * run] - rxf, fb = cam.spans[l - 1][2 * run + 1] - rxf;
                if (fb > fa) {
                    const int ca = (fa >> 1) - 1, cb = ((fb - 1) >> 1) + 2;
                    a = any ? std::min(a, ca) : ca; b = any ? std::max(b, cb) : cb; any = true;
                }
            }
            if (l < nb && hi[l + 1] > lo[l + 1]) {
                const int da = 2 * lo[l + 1] - 2, db = 2 * hi[l + 1] + 1;
                a = any ? std::min(a, da) : da; b = any ? std::max(b, db) : db; any = true;
            }
            if (!any) { lo[l] = hi[l] = 0; continue; }
            lo[l] = std::max(0, a); hi[l] = std::min(w, b);
        }
        for (int l = 0; l <= nb; ++l) { cam.g_runs[l][2 * run] = lo[l]; cam.g_runs[l][2 * run + 1] = hi[l]; }
    }
    for (int l = 0; l <= nb; ++l) {      // merge overlapping runs
        auto &g = cam.g_runs[l];
        if (g[3] > g[2] && g[1] > g[0] && g[2] <= g[1]) { g[1] = std::max(g[1], g[3]); g[0] = std::min(g[0], g[2]); g[2] = g[3] = 0; }
    }
}

// runs of `g` (rect-local) clipped to the panorama-level range [x0, x1) given in panorama coordinates
void clip_runs(const std::array<int, 4> &g, int rx, int x0, int x1, int out[4], int *max_col, double *cols)
{
    *cols = 0;
    for (int r = 0; r < 2; ++r) {
        int a = std::max(g[2 * r], x0 - rx), b = std::min(g[2 * r + 1], x1 - rx);
        if (b <= a) a = b = 0;
        out[2 * r] = a; out[2 * r + 1] = b;
        *max_col = std::max(*max_col, b);
        *cols += b - a;
    }
}

// The multi-band warp stage on the frame kernel of the feather path (k_fs2, out_mode 2).  Every camera's padded feed rect is
// an output plane; the planes are stacked vertically (heights rounded up to whole tiles) into the coordinate system the k_fs2
// setup works in, so each of its tiles sees exactly one camera, "Blender::NO" semantics produce the plane's pixels, and the
// mask byte says which pixels the pyramid needs (the level-0 column runs).  Whole-frame launches only: a latency-mode strip
// keeps the streaming kernel with its per-strip schedule.
int mb_fs2_setup(sb_compositor *c, cudaStream_t s)
{
    c->mbf = Fs2Plan{};
    const int n = c->cfg.n_cameras;
    if (n > FS2_TMAP_CAMS || n > 16 || c->cfg.src_size.width * 3 > 65535) return SB_OK;
    int pw = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
        Camera &cam = c->cams[i];
        cam.mbf_dy = ph;
        ph += div_up(cam.rh, FS2_H) * FS2_H;
        pw = std::max(pw, cam.rw);
    }
    if (pw >= 65536 || ph >= 65536) return SB_OK;
    std::vector<Fs2CamSetup> cs(n);
    std::vector<DevBuf> tabs(n);
    for (int i = 0; i < n; ++i) {
        Camera &cam = c->cams[i];
        int cx[4], cmax = 0;
        double cols = 0;
        clip_runs(cam.g_runs[0], cam.rx, 0, c->wsum[0].v.cols, cx, &cmax, &cols);
        const size_t tstep = ((size_t)cam.rw * sizeof(uint2) + 255) & ~(size_t)255;
        SB_TRY(tabs[i].ensure(tstep * cam.rh));
        SB_TRY(launch_mbs_feather_format(static_cast<const uint2 *>(cam.mb_table.p), cam.mb_tstep, cam.rw, cam.rh, cx, static_cast<uint2 *>(tabs[i].p), tstep, s));
        Fs2CamSetup &f = cs[i];
        f.table = static_cast<const uint2 *>(tabs[i].p); f.tstep = tstep;
        f.ww = cam.rw; f.wh = cam.rh; f.dx = 0; f.dy = cam.mbf_dy;
        f.tx0 = 0; f.ty0 = cam.mbf_dy / FS2_H; f.ntx = div_up(cam.rw, FS2_W); f.nty = div_up(cam.rh, FS2_H);
        SB_TRY(cam.mbf_blocks.ensure((size_t)f.ntx * f.nty * FS2_BLOCK_BYTES));
        f.blocks = static_cast<unsigned char *>(cam.mbf_blocks.p);
    }
    SB_TRY(fs2_build(cs.data(), n, pw, ph, 1.f / 255.f, c->sm_count, false, c->mbf_desc, &c->mbf, s, true));
    c->mbf_pw = pw; c->mbf_ph = ph;
    return SB_OK;
}

// Tile schedule of the streaming warp stage (kernels_mb_stream.cu) for the panorama columns [x0, x1): the tiles of every
// camera's padded rect that intersect the level-0 column runs the pyramid needs there, their descriptors in schedule
// order with the shared-memory ring plan.  Whole frame at setup; a rank's strip (+ halo) in latency mode.
int build_mbs_schedule(sb_compositor *c, int x0, int x1, cudaStream_t s)
{
    c->mbs_ok = false;
    if (!c->mbs_tables) return SB_OK;
    const int n = c->cfg.n_cameras;
    std::vector<unsigned> list;
    MbsSetup ms{};
    for (int i = 0; i < n; ++i) {
        const Camera &cam = c->cams[i];
        ms.rec[i] = static_cast<const uint4 *>(cam.mbs_rec.p); ms.ntx[i] = cam.mbs_ntx;
        int cx[4], cmax = 0;
        double cols = 0;
        clip_runs(cam.g_runs[0], cam.rx, x0, x1, cx, &cmax, &cols);
        for (int ty = 0; ty < cam.mbs_nty; ++ty)
            for (int tx = 0; tx < cam.mbs_ntx; ++tx) {
                const int a0 = tx * SB_FTT_W, a1 = a0 + SB_FTT_W;
                if ((a0 < cx[1] && a1 > cx[0]) || (a0 < cx[3] && a1 > cx[2])) list.push_back((unsigned)i | ((unsigned)(ty * cam.mbs_ntx + tx) << 4));
            }
    }
    if (list.empty()) return SB_OK;
    DevBuf dl;
    SB_TRY(dl.ensure(sizeof(unsigned) * list.size()));
    SB_CUDA(cudaMemcpyAsync(dl.p, list.data(), sizeof(unsigned) * list.size(), cudaMemcpyHostToDevice, s));
    c->mbs_n_tiles = (int)list.size();
    SB_TRY(c->mbs_desc.ensure(sizeof(uint4) * (1 + SB_FTT_MAXC) * list.size()));
    SB_TRY(launch_mbs_descriptors(ms, static_cast<const unsigned *>(dl.p), c->mbs_n_tiles, static_cast<uint4 *>(c->mbs_desc.p),
                                  fts_grid(c->mbs_n_tiles, c->sm_count), s));
    SB_CUDA(cudaStreamSynchronize(s));
    c->mbs_x0 = x0; c->mbs_x1 = x1;
    c->mbs_ok = true;
    return SB_OK;
}

// calibration-time work: geometry, tables, masks, weights
int setup(sb_compositor *c)
{
    const sb_compositor_config &cfg = c->cfg;
    cudaStream_t s = c->setup_stream;
    const int n = cfg.n_cameras;
    c->cams.resize(n);
    c->host_maps = cfg.warper_kind > SB_WARP_SPHERICAL;
    c->undistort = cfg.undistort_map1 != nullptr;
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    SB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    SB_TRY(c->bilin_lut.ensure(1024 * sizeof(uint2)));
    SB_TRY(launch_build_bilin_lut(static_cast<uint2 *>(c->bilin_lut.p), s));
    DevImage ones, xmap, ymap;
    SB_TRY(ones.create(cfg.src_size.height, cfg.src_size.width, SB_8UC1));
    SB_CUDA(cudaMemset2DAsync(ones.v.data, ones.v.step, 255, ones.v.cols, ones.v.rows, s));
    const float zeroT[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        Camera &cam = c->cams[i];
        projector_set(cam.proj, cfg.warper_kind, cfg.warper_scale, cfg.K + 9 * i, cfg.R + 9 * i, zeroT);
        cam.proj.a = cfg.warper_a != 0.f ? cfg.warper_a : 1.f; cam.proj.b = cfg.warper_b != 0.f ? cfg.warper_b : 1.f;
        sb_point br;
        projector_detect_result_roi(cam.proj, cfg.src_size.width, cfg.src_size.height, &cam.tl, &br);
        long long ww = (long long)br.x - cam.tl.x + 1, wh = (long long)br.y - cam.tl.y + 1;
        if (ww <= 0 || wh <= 0 || ww * wh > (1LL << 31)) return fail(SB_ERR_ASSERT, "camera %d: degenerate warped ROI", i);
        cam.ww = (int)ww; cam.wh = (int)wh;
        cam.gain = (cfg.comp_kind == SB_COMP_GAIN && cfg.gains) ? (float)cfg.gains[i] : 1.f;
        if (cfg.comp_kind == SB_COMP_GAIN_BLOCKS) {
            // BlocksGainCompensator::apply: resize(gain_maps_[index], gain_map, image.size(), 0, 0, INTER_LINEAR) (exposure_compensate.cpp:233)
            const sb_image &gm = cfg.gain_maps[i];
            SB_ASSERT(gm.type == SB_32FC1 && gm.data && gm.rows > 0 && gm.cols > 0);
            DevImage st;
            DImage dgm;
            SB_TRY(to_device(gm, st, s, &dgm));
            SB_TRY(cam.gain_full.create(cam.wh, cam.ww, SB_32FC1));
            SB_TRY(launch_resize_linear_32f(dgm, cam.gain_full.v, s));
            SB_CUDA(cudaStreamSynchronize(s));
        }
        tlx = std::min(tlx, cam.tl.x); tly = std::min(tly, cam.tl.y);
        brx = std::max(brx, cam.tl.x + cam.ww); bry = std::max(bry, cam.tl.y + cam.wh);
        // warped mask: w->warp(mask, K, R, INTER_NEAREST, BORDER_CONSTANT) (stitcher.cpp:278-280)
        SB_TRY(cam.mask.create(cam.wh, cam.ww, SB_8UC1));
        if (!c->host_maps) {
            // separable trig tables
            SB_TRY(cam.tables.ensure(sizeof(float) * 2 * (size_t)(cam.ww + cam.wh)));
            float *t = static_cast<float *>(cam.tables.p);
            SB_TRY(launch_build_warp_tables(cam.proj, cam.tl.x, cam.tl.y, cam.ww, cam.wh, t, t + cam.ww, t + 2 * cam.ww, t + 2 * cam.ww + cam.wh, s));
            cam.wt.col_sin = t; cam.wt.col_cos = t + cam.ww; cam.wt.row_a = t + 2 * cam.ww; cam.wt.row_b = t + 2 * cam.ww + cam.wh;
            SB_TRY(xmap.create(cam.wh, cam.ww, SB_32FC1));
            SB_TRY(ymap.create(cam.wh, cam.ww, SB_32FC1));
            SB_TRY(launch_build_maps(cam.proj, cam.tl.x, cam.tl.y, xmap.v, ymap.v, s));
            SB_TRY(launch_remap(ones.v, cam.mask.v, xmap.v, ymap.v, SB_INTER_NEAREST, SB_BORDER_CONSTANT, nullptr, s));
        } else {
            // the libm-heavy projectors (warpers_inl.hpp:302-759): mapBackward on the host once per calibration, exactly as the
            // reference evaluates it; the maps stay on the device for the table builders below, no frame ever sees the host
            std::vector<float> hx((size_t)cam.ww * cam.wh), hy(hx.size());
            projector_build_maps_host(cam.proj, cam.tl, br, hx.data(), hy.data());
            SB_TRY(cam.xmap.create(cam.wh, cam.ww, SB_32FC1));
            SB_TRY(cam.ymap.create(cam.wh, cam.ww, SB_32FC1));
            SB_CUDA(cudaMemcpy2DAsync(cam.xmap.v.data, cam.xmap.v.step, hx.data(), (size_t)cam.ww * 4, (size_t)cam.ww * 4, (size_t)cam.wh, cudaMemcpyHostToDevice, s));
            SB_CUDA(cudaMemcpy2DAsync(cam.ymap.v.data, cam.ymap.v.step, hy.data(), (size_t)cam.ww * 4, (size_t)cam.ww * 4, (size_t)cam.wh, cudaMemcpyHostToDevice, s));
            SB_CUDA(cudaStreamSynchronize(s));
            SB_TRY(launch_remap(ones.v, cam.mask.v, cam.xmap.v, cam.ymap.v, SB_INTER_NEAREST, SB_BORDER_CONSTANT, nullptr, s));
        }
        if (cfg.undistort_map1) {   // the app's fisheye front end (APP64:201-238): fixed-point maps of the frame size
            const sb_image &m1 = cfg.undistort_map1[i], &m2 = cfg.undistort_map2[i];
            SB_ASSERT(m1.type == SB_16SC2 && m2.type == SB_16UC1 && m1.data && m2.data && m1.rows == cfg.src_size.height && m1.cols == cfg.src_size.width &&
                      m2.rows == m1.rows && m2.cols == m1.cols);
            DImage v;
            SB_TRY(to_device(m1, cam.umap1, s, &v));
            SB_TRY(to_device(m2, cam.umap2, s, &v));
            SB_CUDA(cudaStreamSynchronize(s));
        }
        if (cfg.seam_masks) {   // mask_warped = seam_mask & mask_warped (stitcher.cpp:294)
            const sb_image &sm = cfg.seam_masks[i];
            SB_ASSERT(sm.type == SB_8UC1 && sm.rows == cam.wh && sm.cols == cam.ww && sm.data);
            DevImage st;
            DImage dsm;
            SB_TRY(to_device(sm, st, s, &dsm));
            SB_TRY(launch_and_8u(dsm, cam.mask.v, s));
            SB_CUDA(cudaStreamSynchronize(s));
        }
    }
    // Blender::prepare(corners, sizes) -> resultRoi (util.cpp:127-140) -> prepare(Rect)
    sb_rect roi = {tlx, tly, brx - tlx, bry - tly};
    c->dst_roi_final = roi;
    c->out_rect = sb_rect{0, 0, roi.width, roi.height};
    c->crop = cfg.crop_up != 0.f || cfg.crop_down != 0.f || cfg.crop_left != 0 || cfg.crop_right != 0;
    if (c->crop) {
        // UpdateMat / feedSizeRemap (APP64:702, 153): float arithmetic exactly as written there, truncating conversions
        const float keep = 1 - cfg.crop_up - cfg.crop_down;
        const int out_w = (int)((float)roi.width - (float)cfg.crop_left - (float)cfg.crop_right), out_h = (int)((float)roi.height * keep);
        const int yy = (int)((float)out_h / keep * cfg.crop_up);
        if (out_w <= 0 || out_h <= 0 || yy + out_h > roi.height || cfg.crop_left + out_w > roi.width)
            return fail(SB_ERR_BAD_ARG, "crop margins leave no panorama (%d x %d at %d, %d of %d x %d)", out_w, out_h, cfg.crop_left, yy, roi.width, roi.height);
        c->out_rect = sb_rect{cfg.crop_left, yy, out_w, out_h};
    }
    c->num_bands = 0;
    if (cfg.blender_kind == SB_BLEND_MULTI_BAND) {
        double max_len = static_cast<double>(std::max(roi.width, roi.height));
        c->num_bands = std::min(cfg.num_bands, static_cast<int>(std::ceil(std::log(max_len) / std::log(2.0))));
        SB_ASSERT(c->num_bands >= 0 && c->num_bands < 24);
        const int m = 1 << c->num_bands;
        roi.width += (m - roi.width % m) % m;
        roi.height += (m - roi.height % m) % m;
    }
    c->dst_roi = roi;

    if (cfg.blender_kind == SB_BLEND_MULTI_BAND) {
        const int nb = c->num_bands, m = 1 << nb, gap = 3 * m;
        const int wt = cfg.weight_type == SB_32F ? SB_32FC1 : SB_16SC1;
        c->wsum.resize(nb + 1);
        int rows = roi.height, cols = roi.width;
        for (int l = 0; l <= nb; ++l) {
            SB_TRY(c->wsum[l].create_zero(rows, cols, wt, s));
            rows = (rows + 1) / 2; cols = (cols + 1) / 2;
        }
        const int rbr_x = roi.x + roi.width, rbr_y = roi.y + roi.height;
        DevBuf span_scratch;
        for (int i = 0; i < n; ++i) {
            Camera &cam = c->cams[i];
            const FeedRect fr = multiband_feed_rect(roi, cam.tl, cam.ww, cam.wh, nb);      // MultiBandBlender::feed geometry (blenders.cpp:241-269)
            const int width = fr.width, height = fr.height;
            cam.top = fr.top; cam.left = fr.left;
            cam.rx = fr.tl_new.x - roi.x; cam.ry = fr.tl_new.y - roi.y; cam.rw = width; cam.rh = height;
            SB_ASSERT(cam.top >= 0 && cam.left >= 0 && cam.rx >= 0 && cam.ry >= 0);
            // weight pyramid (:282-298) and its contribution to dst_band_weights_ (:324, :349)
            cam.w_pyr.resize(nb + 1);
            SB_TRY(cam.w_pyr[0].create(height, width, wt));
            SB_TRY(launch_mask_to_weight(cam.mask.v, cam.w_pyr[0].v, cam.top, cam.left, s));
            int x_tl = cam.rx, y_tl = cam.ry;
            for (int l = 0; l <= nb; ++l) {
                if (l > 0) {
                    const DImage &p = cam.w_pyr[l - 1].v;
                    SB_TRY(cam.w_pyr[l].create((p.rows + 1) / 2, (p.cols + 1) / 2, wt));
                    SB_TRY(launch_pyr_down(p, cam.w_pyr[l].v, s));
                }
                SB_TRY(launch_weight_accumulate(cam.w_pyr[l].v, c->wsum[l].v, x_tl, y_tl, s));
                cam.spans.emplace_back();
                SB_TRY(weight_spans(cam.w_pyr[l].v, x_tl, span_scratch, s, &cam.spans.back()));
                x_tl /= 2; y_tl /= 2;
            }
        }
        for (int i = 0; i < n; ++i) gaussian_runs(c->cams[i], nb);
        // fast path tables: resolved taps per padded pixel, camera bitmask per band tile
        c->mb_fast = cfg.src_size.width <= 4096 && cfg.src_size.height <= 4096;
        if (c->mb_fast) {
            for (int i = 0; i < n; ++i) {
                Camera &cam = c->cams[i];
                if (cfg.comp_kind == SB_COMP_GAIN_BLOCKS) {      // the gain seen by each padded-rect pixel (copyMakeBorder REFLECT)
                    SB_TRY(cam.gain_rect.create(cam.rh, cam.rw, SB_32FC1));
                    SB_TRY(launch_copy_make_border(cam.gain_full.v, cam.gain_rect.v, cam.top, cam.left, SB_BORDER_REFLECT, s));
                }
                cam.mb_tstep = ((size_t)cam.rw * sizeof(uint2) + 255) & ~(size_t)255;
                SB_TRY(cam.mb_table.ensure(cam.mb_tstep * cam.rh));
                SB_TRY(launch_mb_tap_table(cam.proj, cam.tl.x, cam.tl.y, cam.ww, cam.wh, cam.left, cam.top, cfg.src_size.width,
                                           cfg.src_size.height, static_cast<uint2 *>(cam.mb_table.p), cam.mb_tstep, cam.rw, cam.rh, s,
                                           c->host_maps ? &cam.xmap.v : nullptr, c->host_maps ? &cam.ymap.v : nullptr));
            }
            // streaming warp stage: tile-major taps + source boxes per camera, then the tile schedule for whole-frame launches
            {
                bool ok = cfg.src_size.width * 3 <= 65535 && n <= 16;
                for (int i = 0; i < n && ok; ++i) {
                    Camera &cam = c->cams[i];
                    cam.mbs_ntx = div_up(cam.rw, SB_FTT_W); cam.mbs_nty = div_up(cam.rh, SB_FTT_H);
                    const size_t nt = (size_t)cam.mbs_ntx * cam.mbs_nty;
                    if (nt >= (1u << 27)) { ok = false; break; }
                    SB_TRY(cam.mbs_rec.ensure(sizeof(uint4) * nt));
                    SB_TRY(cam.mbs_tiles.ensure(sizeof(uint2) * SB_FTT_W * SB_FTT_H * nt));
                    SB_TRY(launch_mbs_camera_tiles(static_cast<const uint2 *>(cam.mb_table.p), cam.mb_tstep, cam.rw, cam.rh, cam.mbs_ntx, cam.mbs_nty,
                                                   static_cast<uint4 *>(cam.mbs_rec.p), static_cast<uint2 *>(cam.mbs_tiles.p), s));
                }
                c->mbs_tables = ok;
                SB_TRY(build_mbs_schedule(c, 0, c->wsum[0].v.cols, s));
                if (ok) SB_TRY(mb_fs2_setup(c, s));
            }
            c->mb_tile_mask.resize(nb + 1);
            for (int l = 0; l <= nb; ++l) {
                MbBandGeom g{};
                g.n = n;
                g.lw = c->wsum[l].v.cols; g.lh = c->wsum[l].v.rows;
                for (int i = 0; i < n; ++i) {
                    const Camera &cam = c->cams[i];
                    g.cam[i].weight = cam.w_pyr[l].v.data; g.cam[i].wstep = cam.w_pyr[l].v.step;
                    g.cam[i].rx = cam.rx >> l; g.cam[i].ry = cam.ry >> l;
                    g.cam[i].rw = cam.w_pyr[l].v.cols; g.cam[i].rh = cam.w_pyr[l].v.rows;
                }
                SB_TRY(c->mb_tile_mask[l].ensure(sizeof(uint32_t) * (size_t)div_up(g.lw, 32) * div_up(g.lh, 8)));
                SB_TRY(launch_mb_tile_mask(g, wt == SB_32FC1, g.lw, g.lh, c->wsum[l].v.data, c->wsum[l].v.step, static_cast<uint32_t *>(c->mb_tile_mask[l].p), s));
            }
        }
    } else if (cfg.blender_kind == SB_BLEND_FEATHER || cfg.blender_kind == SB_BLEND_NO) {
        // Blender::NO rides on the feather tables: the "distance" field carries the mask byte (its OR is dst_mask_,
        // the last camera with a non-zero mask supplies the pixel, blenders.cpp:81-102), "sharpness" 1/255 marks
        // the tiles whose mask is 255 throughout
        const bool noblend = cfg.blender_kind == SB_BLEND_NO;
        c->stream_sharpness = noblend ? 1.f / 255.f : cfg.sharpness;
        c->wsum.resize(1);
        SB_TRY(c->wsum[0].create_zero(roi.height, roi.width, SB_32FC1, s));
        DevBuf scratch;
        for (int i = 0; i < n; ++i) {
            Camera &cam = c->cams[i];
            SB_TRY(cam.feather_w.create(cam.wh, cam.ww, SB_32FC1));
            if (noblend) SB_TRY(launch_convert(cam.mask.v, cam.feather_w.v, s));      // the mask byte as an exact float
            else SB_TRY(launch_distance_l1(cam.mask.v, cam.feather_w.v, scratch, s));
            cam.feather_tstep = ((size_t)cam.ww * sizeof(uint2) + 255) & ~(size_t)255;
            SB_TRY(cam.feather_table.ensure(cam.feather_tstep * cam.wh));
            SB_TRY(launch_build_feather_table(cam.proj, cam.tl.x, cam.tl.y, cam.feather_w.v, cfg.src_size.width, cfg.src_size.height,
                                              static_cast<uint2 *>(cam.feather_table.p), cam.feather_tstep, s,
                                              c->host_maps ? &cam.xmap.v : nullptr, c->host_maps ? &cam.ymap.v : nullptr));
            SB_TRY(launch_weight_from_dist(cam.feather_w.v, c->stream_sharpness, s));
            SB_TRY(launch_weight_accumulate(cam.feather_w.v, c->wsum[0].v, cam.tl.x - roi.x, cam.tl.y - roi.y, s));
            cam.spans.emplace_back();
            SB_TRY(weight_spans(cam.feather_w.v, cam.tl.x - roi.x, scratch, s, &cam.spans.back()));
            cam.gain_bytes = 4.0 * cam.wh * ((cam.spans.back()[1] - cam.spans.back()[0]) + (cam.spans.back()[3] - cam.spans.back()[2]));
        }
        // per panorama tile: which cameras carry weight there
        const int tiles_x = div_up(roi.width, SB_FT_W), tiles_y = div_up(roi.height, SB_FT_H);
        SB_TRY(c->tile_cams.ensure(sizeof(uint32_t) * ((size_t)tiles_x * tiles_y + 1)));
        SB_CUDA(cudaMemsetAsync(c->tile_cams.p, 0, sizeof(uint32_t) * ((size_t)tiles_x * tiles_y + 1), s));
        unsigned *bad = static_cast<unsigned *>(c->tile_cams.p) + (size_t)tiles_x * tiles_y;
        for (int i = 0; i < n; ++i) {
            Camera &cam = c->cams[i];
            FeatherCam fc{};
            fc.table = static_cast<const uint2 *>(cam.feather_table.p); fc.tstep = cam.feather_tstep;
            fc.ww = cam.ww; fc.wh = cam.wh; fc.dx = cam.tl.x - roi.x; fc.dy = cam.tl.y - roi.y;
            SB_TRY(launch_feather_tile_cams(fc, i, roi.width, roi.height, static_cast<uint32_t *>(c->tile_cams.p), bad, s));
        }
        unsigned h_bad = 1;
        SB_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof h_bad, cudaMemcpyDeviceToHost, s));
        SB_CUDA(cudaStreamSynchronize(s));
        c->feather_fast = h_bad == 0 && cfg.src_size.width <= 8192 && cfg.src_size.height <= 8192;
        if (c->feather_fast) {   // tile-major tables, source boxes and per-tile descriptors for the streaming kernel
            FtsSetup ta{};
            ta.n = n;
            ta.tiles_x = div_up(roi.width, SB_FTT_W);
            ta.n_tiles = ta.tiles_x * div_up(roi.height, SB_FTT_H);
            for (int i = 0; i < n; ++i) {
                Camera &cam = c->cams[i];
                const int dx = cam.tl.x - roi.x, dy = cam.tl.y - roi.y;
                cam.ftx0 = dx / SB_FTT_W; cam.fty0 = dy / SB_FTT_H;
                cam.fntx = div_up(dx + cam.ww, SB_FTT_W) - cam.ftx0; cam.fnty = div_up(dy + cam.wh, SB_FTT_H) - cam.fty0;
                SB_TRY(cam.feather_tiles.ensure(sizeof(uint2) * SB_FTT_W * SB_FTT_H * (size_t)cam.fntx * cam.fnty));
                SB_TRY(cam.feather_rec.ensure(sizeof(uint4) * (size_t)cam.fntx * cam.fnty));
                SB_TRY(launch_fts_camera_tiles(static_cast<const uint2 *>(cam.feather_table.p), cam.feather_tstep, cam.ww, cam.wh, dx, dy,
                                               cam.ftx0, cam.fty0, cam.fntx, cam.fnty, c->stream_sharpness, static_cast<uint4 *>(cam.feather_rec.p),
                                               static_cast<uint2 *>(cam.feather_tiles.p), s));
                ta.cam[i].rec = static_cast<const uint4 *>(cam.feather_rec.p);
                ta.cam[i].tx0 = cam.ftx0; ta.cam[i].ty0 = cam.fty0; ta.cam[i].ntx = cam.fntx; ta.cam[i].nty = cam.fnty;
            }
            SB_TRY(c->tma_desc.ensure(sizeof(uint4) * (1 + SB_FTT_MAXC) * (size_t)ta.n_tiles + 16));
            int *status = reinterpret_cast<int *>(static_cast<uint4 *>(c->tma_desc.p) + (size_t)(1 + SB_FTT_MAXC) * ta.n_tiles);
            SB_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
            SB_TRY(launch_fts_descriptors(ta, static_cast<uint4 *>(c->tma_desc.p), status, fts_grid(ta.n_tiles, c->sm_count), s));
            int h_status = 1;
            SB_CUDA(cudaMemcpyAsync(&h_status, status, sizeof h_status, cudaMemcpyDeviceToHost, s));
            SB_CUDA(cudaStreamSynchronize(s));
            c->feather_tma = h_status == 0 && n <= 16;
            // ... and for its successor (4-byte entries, tensor-TMA source boxes).  Its tiles cover what compose hands back:
            // with crop margins (APP64:150-177) the cropped-away part of the composite is never scheduled.
            std::vector<Fs2CamSetup> fc(n);
            const int otx = div_up(c->out_rect.width, FS2_W), oty = div_up(c->out_rect.height, FS2_H);
            for (int i = 0; i < n; ++i) {
                Camera &cam = c->cams[i];
                const int dx = cam.tl.x - roi.x - c->out_rect.x, dy = cam.tl.y - roi.y - c->out_rect.y;
                auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
                cam.f2tx0 = std::max(0, fdiv(dx, FS2_W)); cam.f2ty0 = std::max(0, fdiv(dy, FS2_H));
                cam.f2ntx = std::max(0, std::min(otx, div_up(dx + cam.ww, FS2_W)) - cam.f2tx0);
                cam.f2nty = std::max(0, std::min(oty, div_up(dy + cam.wh, FS2_H)) - cam.f2ty0);
                if (cam.f2ntx == 0 || cam.f2nty == 0) cam.f2ntx = cam.f2nty = 0;
                SB_TRY(cam.fs2_blocks.ensure((size_t)FS2_BLOCK_BYTES * std::max(1, cam.f2ntx * cam.f2nty)));
                fc[i] = Fs2CamSetup{static_cast<const uint2 *>(cam.feather_table.p), cam.feather_tstep, cam.ww, cam.wh, dx, dy,
                                    cam.f2tx0, cam.f2ty0, cam.f2ntx, cam.f2nty, static_cast<unsigned char *>(cam.fs2_blocks.p)};
                cam.tmaps.clear();
            }
            SB_TRY(fs2_build(fc.data(), n, c->out_rect.width, c->out_rect.height, c->stream_sharpness, c->sm_count, cfg.comp_kind == SB_COMP_GAIN_BLOCKS,
                             c->fs2_desc, &c->fs2, s));
            if (c->fs2.ok && c->fs2.gain_tma)
                for (int i = 0; i < n; ++i) {      // the gain map with one tile of zero padding all round (see fs2_build): every gain tile lies inside it
                    Camera &cam = c->cams[i];
                    SB_TRY(cam.gain_pad.create_zero(cam.wh + 2 * FS2_H, cam.ww + 2 * FS2_W + 4, SB_32FC1, s));
                    SB_CUDA(cudaMemcpy2DAsync(static_cast<char *>(cam.gain_pad.v.data) + (size_t)FS2_H * cam.gain_pad.v.step + (size_t)FS2_W * 4, cam.gain_pad.v.step,
                                              cam.gain_full.v.data, cam.gain_full.v.step, (size_t)cam.ww * 4, (size_t)cam.wh, cudaMemcpyDeviceToDevice, s));
                    SB_TRY(fs2_encode_gain_tmap(cam.gain_pad.v.ptr<float>(), cam.gain_pad.v.step, cam.ww + 2 * FS2_W + 4, cam.wh + 2 * FS2_H, &c->fs2_gtmap[i]));
                }
            if (cfg.crop_app_fill) {   // camera 0's table entry of warped pixel (0, 0): what an uncovered pixel gathers (APP64:165-172)
                uint2 e;
                SB_CUDA(cudaMemcpyAsync(&e, c->cams[0].feather_table.p, sizeof e, cudaMemcpyDeviceToHost, s));
                SB_CUDA(cudaStreamSynchronize(s));
                c->fill_tex[0] = e.x; c->fill_tex[1] = e.y;
            }
        }
        SB_CUDA(cudaStreamSynchronize(s));
    }
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}

double img_bytes(const DImage &d) { return (double)d.rows * d.cols * elem_size(d.type); }

// PROF(name, bytes, call): run `call`; when the slot is in profiling mode bracket it with events
#define PROF(NAME, BYTES, CALL)                                                          \
    do {                                                                                 \
        ProfRec r_{NAME, (double)(BYTES), nullptr, nullptr};                             \
        if (s.prof) {                                                                    \
            SB_CUDA(cudaEventCreate(&r_.e0)); SB_CUDA(cudaEventCreate(&r_.e1));          \
            SB_CUDA(cudaEventRecord(r_.e0, st));                                         \
        }                                                                                \
        SB_TRY(CALL);                                                                    \
        if (s.prof) { SB_CUDA(cudaEventRecord(r_.e1, st)); s.prof->push_back(r_); }      \
    } while (0)

// ---- the multi-band fast path, stage by stage.  [x0, x1) is the range of padded-panorama columns (level-0
// coordinates, multiples of 2^num_bands) this call produces: the whole width for a frame on one GPU, one
// column strip in latency mode, where the host exchanges halo columns between the stages.
int fs2_tmaps(sb_compositor *c, int i, const DImage &src, CUtensorMap *out);

int mb_warp_stage(sb_compositor *c, Slot &s, const std::vector<DImage> &src, int x0, int x1)
{
    // K1: remap + gain + convertTo(16S) + copyMakeBorder for every camera, one launch
    const int n = c->cfg.n_cameras;
    cudaStream_t st = s.stream;
    MbWarpArgs a{};
    a.n = n;
    a.bilin_lut = static_cast<const uint2 *>(c->bilin_lut.p);
    double bytes = 0;
    int mw = 0, mh = 0;
    for (int i = 0; i < n; ++i) {
        const Camera &cam = c->cams[i];
        MbWarpCam &wc = a.cam[i];
        wc.src = src[i].ptr<uint8_t>(); wc.sstep = src[i].step;
        wc.table = static_cast<const uint2 *>(cam.mb_table.p); wc.tstep = cam.mb_tstep;
        wc.g0 = static_cast<uint32_t *>(s.grgbx[i][0].buf.p); wc.gstep = s.grgbx[i][0].step;
        wc.rw = cam.rw; wc.rh = cam.rh; wc.gain = cam.gain;
        if (c->cfg.comp_kind == SB_COMP_GAIN_BLOCKS) { wc.gmap = cam.gain_rect.v.ptr<float>(); wc.gmstep = cam.gain_rect.v.step; }
        double cols = 0;
        int cmax = 0;
        clip_runs(cam.g_runs[0], cam.rx, x0, x1, wc.cx, &cmax, &cols);
        if (cols == 0) continue;
        mw = std::max(mw, cmax); mh = std::max(mh, cam.rh);
        // the source pixels behind the produced columns are gathered ~once; table entry in, RGBX pixel out
        bytes += img_bytes(src[i]) * std::min(1.0, cols / (double)std::min(cam.ww, src[i].cols)) + cols * cam.rh * (8 + 4);
    }
    if (mw == 0) return SB_OK;
    bool aligned = true;                                     // cp.async source boxes: 16-byte aligned rows
    for (int i = 0; i < n; ++i) aligned = aligned && (reinterpret_cast<uintptr_t>(src[i].data) & 15) == 0 && (src[i].step & 15) == 0;
    if (c->mbf.ok && c->mbf_enabled && c->mbs_enabled && aligned && c->strip_world == 1 && x0 <= 0 && x1 >= c->wsum[0].v.cols) {
        Fs2Args fa{};
        fa.n = n;
        double fbytes = c->mbf.table_bytes;
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            Fs2Cam &fc = fa.cam[i];
            fc.blocks = static_cast<const unsigned char *>(cam.mbf_blocks.p);
            fc.gain = cam.gain; fc.dx = 0; fc.dy = cam.mbf_dy;
            fc.gmap = a.cam[i].gmap; fc.gmstep = (unsigned)a.cam[i].gmstep;
            fc.mb_out = reinterpret_cast<unsigned char *>(a.cam[i].g0); fc.mb_step = (unsigned)a.cam[i].gstep;
            fc.ow = cam.rw; fc.oh = cam.rh;
            SB_TRY(fs2_tmaps(c, i, src[i], &fa.tmap[i * FS2_NCLS]));
            const double cols = (a.cam[i].cx[1] - a.cam[i].cx[0]) + (a.cam[i].cx[3] - a.cam[i].cx[2]);
            fbytes += img_bytes(src[i]) * std::min(1.0, cols / (double)std::min(cam.ww, src[i].cols)) + cols * cam.rh * 4;
        }
        fa.desc = static_cast<const uint4 *>(c->mbf_desc.p);
        for (int k = 0; k < FS2_NCLS; ++k) fa.cls_w[k] = c->mbf.cls_w[k];
        fa.sharpness = 1.f / 255.f; fa.no_blend = 1;
        fa.pw = c->mbf_pw; fa.ph = c->mbf_ph;
        fa.n_tiles = c->mbf.n_tiles; fa.per_cta = c->mbf.per_cta;
        PROF("mb_warp", fbytes, launch_fs2(fa, c->cfg.comp_kind != SB_COMP_NO, 2, c->mbf.grid, st));
        return SB_OK;
    }
    if (c->mbs_ok && c->mbs_enabled && aligned && c->mbs_x0 <= std::max(x0, 0) && std::min(x1, c->wsum[0].v.cols) <= c->mbs_x1) {
        MbStreamArgs sa{};
        sa.n = n;
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            MbStreamCam &wc = sa.cam[i];
            wc.src = src[i].ptr<uint8_t>(); wc.sstep = (unsigned)src[i].step;
            wc.tiles = static_cast<const uint2 *>(cam.mbs_tiles.p);
            wc.g0 = a.cam[i].g0; wc.gstep = (unsigned)a.cam[i].gstep;
            wc.rw = cam.rw; wc.rh = cam.rh; wc.gain = cam.gain;
            wc.gmap = a.cam[i].gmap; wc.gmstep = (unsigned)a.cam[i].gmstep;
        }
        sa.desc = static_cast<const uint4 *>(c->mbs_desc.p);
        sa.bilin_lut = a.bilin_lut;
        sa.n_tiles = c->mbs_n_tiles;
        PROF("mb_warp", bytes, launch_mb_warp_stream(sa, c->cfg.comp_kind != SB_COMP_NO, c->sm_count, st));
        return SB_OK;
    }
    PROF("mb_warp", bytes, launch_mb_warp(a, c->cfg.comp_kind != SB_COMP_NO, mw, mh, st));
    return SB_OK;
}

// [ox0, ox1): columns of level l+1 (panorama level coordinates) to produce.  Returns false when there is nothing to do.
bool fill_down(sb_compositor *c, Slot &s, int l, int ox0, int ox1, MbPyrArgs &a, int tiles_x[SB_MAX_CAMERAS], int tiles[SB_MAX_CAMERAS],
               int &mw, int &mh, double &bytes)
{
    const int n = c->cfg.n_cameras;
    a = MbPyrArgs{};
    a.n = n;
    mw = mh = 0;
    for (int i = 0; i < n; ++i) {
        const Camera &cam = c->cams[i];
        const RawImage &in = s.grgbx[i][l], &out = s.grgbx[i][l + 1];
        a.cam[i].src = static_cast<const uint32_t *>(in.buf.p); a.cam[i].sstep = in.step; a.cam[i].sw = in.cols; a.cam[i].sh = in.rows;
        a.cam[i].dst = static_cast<uint32_t *>(out.buf.p); a.cam[i].dstep = out.step;
        const int rx = cam.rx >> (l + 1);
        double cols = 0;
        int cmax = 0;
        clip_runs(cam.g_runs[l + 1], rx, ox0, ox1, a.cam[i].ox, &cmax, &cols);
        tiles_x[i] = tiles[i] = 0;
        if (cols == 0) continue;
        tiles_x[i] = div_up(div_up(cmax, 2), 32);
        tiles[i] = tiles_x[i] * div_up(div_up(out.rows, 2), 8);
        mw = std::max(mw, cmax); mh = std::max(mh, out.rows);
        bytes += (in.bytes() + out.bytes()) * cols / out.cols;
    }
    return mw > 0;
}

int mb_down_stage(sb_compositor *c, Slot &s, int l, int ox0, int ox1)
{
    // K2: Gaussian level l -> l+1 for every camera, one launch
    cudaStream_t st = s.stream;
    MbPyrListArgs a;
    int tx[SB_MAX_CAMERAS], t[SB_MAX_CAMERAS], mw, mh;
    double bytes = 0;
    if (!fill_down(c, s, l, ox0, ox1, a.p, tx, t, mw, mh, bytes)) return SB_OK;
    static const char *const names[] = {"mb_pyr_down_L0", "mb_pyr_down_L1", "mb_pyr_down_L2", "mb_pyr_down_L3", "mb_pyr_down_L4", "mb_pyr_down_L5+"};
    if (c->pyr_tma_enabled && !s.g_tmap.empty()) {          // tile-staged form (kernels_mb_pyr.cu); set_fused(18) keeps the gather form
        MbPyrTmaArgs ta;
        ta.list = a;
        for (int i = 0; i < c->cfg.n_cameras; ++i) ta.map[i] = s.g_tmap[i][l];
        PROF(names[std::min(l, 5)], bytes, launch_mb_pyr_down_tma(ta, st));
        return SB_OK;
    }
    PROF(names[std::min(l, 5)], bytes, launch_mb_pyr_down_list(a, st));
    return SB_OK;
}

// Gaussian levels l0 -> l0+1 -> ... -> l1 in ONE cooperative launch; lo[l] / hi[l]: columns of level l to produce
int mb_down_tail(sb_compositor *c, Slot &s, int l0, int l1, const int *lo, const int *hi)
{
    cudaStream_t st = s.stream;
    MbPyrTailArgs a{};
    double bytes = 0;
    for (int l = l0; l < l1; ++l) {
        const int k = a.n_levels;
        int tx[SB_MAX_CAMERAS], t[SB_MAX_CAMERAS], mw, mh;
        if (!fill_down(c, s, l, lo[l + 1], hi[l + 1], a.level[k], tx, t, mw, mh, bytes)) continue;
        int first = 0;
        for (int i = 0; i < c->cfg.n_cameras; ++i) { a.first_item[k][i] = first; a.tiles_x[k][i] = std::max(1, tx[i]); first += t[i]; }
        a.items[k] = first;
        ++a.n_levels;
    }
    if (a.n_levels == 0) return SB_OK;
    PROF("mb_pyr_tail", bytes, launch_mb_pyr_tail(a, c->sm_count, st));
    return SB_OK;
}

// [bx0, bx1): columns of band l (panorama level coordinates) to produce.  Returns false when there is nothing to do.
bool fill_band(sb_compositor *c, Slot &s, int l, int bx0, int bx1, MbBandArgs &a, double &bytes)
{
    const sb_compositor_config &cfg = c->cfg;
    const int n = cfg.n_cameras, nb = c->num_bands;
    a = MbBandArgs{};
    a.g.n = n;
    const DImage &ws = c->wsum[l].v;
    a.x_begin = std::max(0, bx0); a.x_end = std::min(ws.cols, bx1);
    if (a.x_end <= a.x_begin) return false;
    const double part = (double)(a.x_end - a.x_begin) / ws.cols;
    for (int i = 0; i < n; ++i) {
        const Camera &cam = c->cams[i];
        MbBandCam &bc = a.g.cam[i];
        const RawImage &f = s.grgbx[i][l];
        bc.fine = static_cast<const uint32_t *>(f.buf.p); bc.fstep = f.step;
        if (l < nb) { bc.coarse = static_cast<const uint32_t *>(s.grgbx[i][l + 1].buf.p); bc.cstep = s.grgbx[i][l + 1].step; }
        bc.weight = cam.w_pyr[l].v.data; bc.wstep = cam.w_pyr[l].v.step;
        bc.rx = cam.rx >> l; bc.ry = cam.ry >> l; bc.rw = f.cols; bc.rh = f.rows;
        const auto &sp = cam.spans[l];
        const double frac = std::min(1.0, (double)((sp[1] - sp[0]) + (sp[3] - sp[2])) / std::max(1, f.cols));
        bytes += part * frac * (f.bytes() * (l < nb ? 1.25 : 1.0) + img_bytes(cam.w_pyr[l].v));
    }
    a.g.lw = ws.cols; a.g.lh = ws.rows;
    a.tile_mask = static_cast<const uint32_t *>(c->mb_tile_mask[l].p); a.tiles_x = div_up(ws.cols, 32);
    a.wsum = ws.data; a.wsum_step = ws.step;
    if (l < nb) { a.coarse_r = static_cast<const short4 *>(s.rband[l + 1].buf.p); a.coarse_r_step = s.rband[l + 1].step; bytes += part * s.rband[l + 1].bytes(); }
    if (l == 0) {
        a.out = s.out.v.data; a.out_step = s.out.v.step;
        a.out_mask = s.want_mask ? s.out_mask.v.ptr<uint8_t>() : nullptr; a.mask_step = s.out_mask.v.step;
        a.out_w = s.out.v.cols; a.out_h = s.out.v.rows;
        a.x_end = std::min(a.x_end, a.out_w);
        bytes += part * (img_bytes(s.out.v) + (s.want_mask ? img_bytes(s.out_mask.v) : 0) + (double)a.out_w * a.out_h * elem_size(ws.type));
    } else {
        a.out = s.rband[l].buf.p; a.out_step = s.rband[l].step;
        bytes += part * (s.rband[l].bytes() + img_bytes(ws));
    }
    return true;
}

int mb_band_stage(sb_compositor *c, Slot &s, int l, int bx0, int bx1)
{
    // K3: band l for every camera, restored coarse -> fine
    cudaStream_t st = s.stream;
    MbBandArgs a;
    double bytes = 0;
    if (!fill_band(c, s, l, bx0, bx1, a, bytes)) return SB_OK;
    const bool fin = l == 0;
    static const char *const names[] = {"mb_band_final", "mb_band_L1", "mb_band_L2", "mb_band_L3", "mb_band_L4", "mb_band_L5+"};
    PROF(names[std::min(l, 5)], bytes,
         launch_mb_band(a, c->cfg.weight_type == SB_32F, l < c->num_bands, fin, s.out.v.type == SB_8UC3, st));
    return SB_OK;
}

// bands l_hi, l_hi - 1, ..., l_lo (l_lo >= 1) in ONE cooperative launch; lo[l] / hi[l]: columns of band l to produce
int mb_band_head(sb_compositor *c, Slot &s, int l_hi, int l_lo, const int *lo, const int *hi)
{
    cudaStream_t st = s.stream;
    MbBandHeadArgs a{};
    a.top_is_top = l_hi == c->num_bands;
    double bytes = 0;
    for (int l = l_hi; l >= l_lo; --l) {
        const int k = a.n_levels;
        if (!fill_band(c, s, l, lo[l], hi[l], a.level[k], bytes)) {
            if (k == 0) a.top_is_top = 0;      // (cannot happen for a non-empty strip; keep the flag honest)
            continue;
        }
        a.tiles_x[k] = div_up(a.level[k].g.lw, 64);
        a.items[k] = a.tiles_x[k] * div_up(a.level[k].g.lh, 16);
        ++a.n_levels;
    }
    if (a.n_levels == 0) return SB_OK;
    PROF("mb_band_head", bytes, launch_mb_band_head(a, c->cfg.weight_type == SB_32F, c->sm_count, st));
    return SB_OK;
}

// Gaussian levels l0 -> ... -> nb and bands nb ... l_lo in ONE launch (k_mb_coarse): with the warp stage, the level-0
// pyrDown and the final band a multi-band frame is 4 launches instead of 12
int mb_coarse(sb_compositor *c, Slot &s, int l0, int l_lo, const int *g_lo, const int *g_hi, const int *b_lo, const int *b_hi)
{
    cudaStream_t st = s.stream;
    const int nb = c->num_bands;
    MbCoarseArgs a{};
    double bytes = 0;
    for (int l = l0; l < nb; ++l) {
        const int k = a.down.n_levels;
        int tx[SB_MAX_CAMERAS], t[SB_MAX_CAMERAS], mw, mh;
        if (!fill_down(c, s, l, g_lo[l + 1], g_hi[l + 1], a.down.level[k], tx, t, mw, mh, bytes)) continue;
        int first = 0;
        for (int i = 0; i < c->cfg.n_cameras; ++i) { a.down.first_item[k][i] = first; a.down.tiles_x[k][i] = std::max(1, tx[i]); first += t[i]; }
        if (first == 0) continue;
        a.down.items[k] = first;
        ++a.down.n_levels;
    }
    a.band.top_is_top = 1;
    for (int l = nb; l >= l_lo; --l) {
        const int k = a.band.n_levels;
        if (!fill_band(c, s, l, b_lo[l], b_hi[l], a.band.level[k], bytes)) {
            if (k == 0) a.band.top_is_top = 0;
            continue;
        }
        a.band.tiles_x[k] = div_up(a.band.level[k].g.lw, 64);
        a.band.items[k] = a.band.tiles_x[k] * div_up(a.band.level[k].g.lh, 16);
        ++a.band.n_levels;
    }
    if (a.down.n_levels + a.band.n_levels == 0) return SB_OK;
    if (!s.mb_sync.p) {
        SB_TRY(s.mb_sync.ensure(sizeof(unsigned long long) * (2 + 2 * SB_MB_MAX_FUSED_LEVELS)));
        SB_CUDA(cudaMemsetAsync(s.mb_sync.p, 0, sizeof(unsigned long long) * (2 + 2 * SB_MB_MAX_FUSED_LEVELS), st));
        s.mb_totals.assign(2 + 2 * SB_MB_MAX_FUSED_LEVELS, 0ull);
    }
    a.sync = static_cast<unsigned long long *>(s.mb_sync.p);
    PROF("mb_coarse", bytes, launch_mb_coarse(a, s.mb_totals.data(), c->cfg.weight_type == SB_32F, c->sm_count, c->slots.size() == 1, st));
    return SB_OK;
}

// the whole multi-band fast path for the column ranges g_lo/g_hi (Gaussian levels) and b_lo/b_hi (bands)
int mb_frame(sb_compositor *c, Slot &s, const std::vector<DImage> &src, const int *g_lo, const int *g_hi, const int *b_lo, const int *b_hi)
{
    const int nb = c->num_bands;
    SB_TRY(mb_warp_stage(c, s, src, g_lo[0], g_hi[0]));
    // Default: one launch per level (warp, pyrDown x n, band x n+1).  set_fused(16): warp, pyrDown 0 -> 1, every coarser level
    // and every band but the last in ONE launch (k_mb_coarse), the final band - 4 launches.  Fewer launches did not mean less
    // time: the fused coarse launch is a chain of 2n-1 dependent stages whose CTAs wait on each other's counters, while
    // separate launches of the tile-staged kernels cost ~9 us each in-stream and, with several frames in flight, hide
    // behind the other frames' heavy kernels (measured on B200, C3: 8 in flight 105 vs 126 us per frame; one in flight
    // 169 vs ~200 us).
    const bool coarse = c->mb_multilevel == 2;
    if (coarse && nb >= 2 && nb <= SB_MB_MAX_FUSED_LEVELS && c->strip_world == 1) {
        SB_TRY(mb_down_stage(c, s, 0, g_lo[1], g_hi[1]));
        SB_TRY(mb_coarse(c, s, 1, 1, g_lo, g_hi, b_lo, b_hi));
        SB_TRY(mb_band_stage(c, s, 0, b_lo[0], b_hi[0]));
        return SB_OK;
    }
    if (c->mb_multilevel == 3 && nb >= 4 && nb <= SB_MB_MAX_FUSED_LEVELS && c->strip_world == 1) {
        // hybrid (set_fused(19)): the levels that carry bytes keep their own tile-staged launches, the tiny ones (level >= 2:
        // < 15 MB in total, each launch latency-bound) share one k_mb_coarse launch - 6 launches per frame
        SB_TRY(mb_down_stage(c, s, 0, g_lo[1], g_hi[1]));
        SB_TRY(mb_down_stage(c, s, 1, g_lo[2], g_hi[2]));
        SB_TRY(mb_coarse(c, s, 2, 2, g_lo, g_hi, b_lo, b_hi));
        SB_TRY(mb_band_stage(c, s, 1, b_lo[1], b_hi[1]));
        SB_TRY(mb_band_stage(c, s, 0, b_lo[0], b_hi[0]));
        return SB_OK;
    }
    const bool multilevel = c->mb_multilevel == 1;
    if (multilevel && nb >= 3 && nb <= SB_MB_MAX_FUSED_LEVELS) {
        SB_TRY(mb_down_stage(c, s, 0, g_lo[1], g_hi[1]));
        SB_TRY(mb_down_tail(c, s, 1, nb, g_lo, g_hi));
        SB_TRY(mb_band_head(c, s, nb, 1, b_lo, b_hi));
        SB_TRY(mb_band_stage(c, s, 0, b_lo[0], b_hi[0]));
    } else {
        for (int l = 0; l < nb; ++l) SB_TRY(mb_down_stage(c, s, l, g_lo[l + 1], g_hi[l + 1]));
        for (int l = nb; l >= 0; --l) SB_TRY(mb_band_stage(c, s, l, b_lo[l], b_hi[l]));
    }
    return SB_OK;
}

// k_fs2 arguments that depend on the calibration only
void fs2_static_args(sb_compositor *c, Fs2Args &a)
{
    const int n = c->cfg.n_cameras;
    a.n = n;
    for (int i = 0; i < n; ++i) {
        Camera &cam = c->cams[i];
        Fs2Cam &fc = a.cam[i];
        fc.blocks = static_cast<const unsigned char *>(cam.fs2_blocks.p);
        fc.gain = cam.gain;
        fc.dx = cam.tl.x - c->dst_roi.x - c->out_rect.x; fc.dy = cam.tl.y - c->dst_roi.y - c->out_rect.y;      // (in the coordinates of what compose hands back)
        if (c->cfg.comp_kind == SB_COMP_GAIN_BLOCKS) { fc.gmap = cam.gain_full.v.ptr<float>(); fc.gmstep = (unsigned)cam.gain_full.v.step; }
    }
    a.desc = static_cast<const uint4 *>(c->fs2_desc.p);
    for (int k = 0; k < FS2_NCLS; ++k) a.cls_w[k] = c->fs2.cls_w[k];
    a.sharpness = c->stream_sharpness;
    a.no_blend = c->cfg.blender_kind == SB_BLEND_NO;
    a.n_tiles = c->fs2.n_tiles; a.per_cta = c->fs2.per_cta; a.steady = c->fs2.steady ? 1 : 0;
    a.gain_tma = c->fs2.gain_tma ? 1 : 0;
    if (c->fs2.gain_tma) std::memcpy(a.gtmap, c->fs2_gtmap, sizeof(CUtensorMap) * (size_t)n);
    a.fill_on = c->cfg.crop_app_fill ? 1 : 0; a.fill_tex[0] = c->fill_tex[0]; a.fill_tex[1] = c->fill_tex[1];
}

// the tensor maps of one camera's frame buffer (encoded on first sight, then cached)
int fs2_tmaps(sb_compositor *c, int i, const DImage &src, CUtensorMap *out)
{
    Camera &cam = c->cams[i];
    Camera::TmapSet *hit = nullptr, *lru = nullptr;
    for (auto &t : cam.tmaps) {
        if (t.p == src.data && t.step == src.step) { hit = &t; break; }
        if (!lru || t.stamp < lru->stamp) lru = &t;
    }
    if (!hit) {
        if (cam.tmaps.size() < 32) { cam.tmaps.emplace_back(); hit = &cam.tmaps.back(); }
        else hit = lru;
        hit->p = src.data; hit->step = src.step;
        SB_TRY(fs2_encode_tmaps(src.data, src.step, src.cols, src.rows, c->fs2.ok ? c->fs2.cls_w : c->mbf.cls_w, hit->m));
    }
    hit->stamp = ++c->tmap_clock;
    std::memcpy(out, hit->m, sizeof hit->m);
    return SB_OK;
}

// Staged (reference-shaped) warp of camera i with BlocksGainCompensator::apply (exposure_compensate.cpp:225-246): warp to 8UC3,
// multiply by the resized gain map with uchar saturation, then - multi-band - copyMakeBorder(BORDER_REFLECT) and
// convertTo(CV_16S) into the padded feed rect `dst16` (blenders.cpp:272-274).  dst16 == nullptr: the 8UC3 result stays in s.warped.
int staged_warp_blocks(sb_compositor *c, Slot &s, int i, const DImage &src, const DImage *dst16)
{
    const Camera &cam = c->cams[i];
    cudaStream_t st = s.stream;
    SB_TRY(s.warped.create(cam.wh, cam.ww, SB_8UC3));
    PROF("warp_fused", img_bytes(src) + img_bytes(s.warped.v),
         launch_warp_fused(cam.proj, cam.wt, src, cam.ww, cam.wh, 0, 0, 1.f, false, s.warped.v, st));
    PROF("gain_blocks", img_bytes(s.warped.v) * 2 + img_bytes(cam.gain_full.v), launch_mul_map_8u(s.warped.v, cam.gain_full.v, st));
    if (!dst16) return SB_OK;
    SB_TRY(s.warped_pad.create(dst16->rows, dst16->cols, SB_8UC3));
    PROF("copy_make_border", img_bytes(s.warped.v) + img_bytes(s.warped_pad.v),
         launch_copy_make_border(s.warped.v, s.warped_pad.v, cam.top, cam.left, SB_BORDER_REFLECT, st));
    PROF("convert_16s", img_bytes(s.warped_pad.v) * 3, launch_convert(s.warped_pad.v, *dst16, st));
    return SB_OK;
}

int run_frame(sb_compositor *c, Slot &s, const std::vector<DImage> &src)
{
    const sb_compositor_config &cfg = c->cfg;
    const int n = cfg.n_cameras;
    const bool gain_on = cfg.comp_kind != SB_COMP_NO, blocks = cfg.comp_kind == SB_COMP_GAIN_BLOCKS;

    cudaStream_t st = s.stream;
    DImage none;
    static_assert(FS2_W == SB_FTT_W && FS2_H == SB_FTT_H, "both streaming kernels use the same panorama tile grid");
    bool aligned_src = true;                      // bulk / tensor copies need 16-byte aligned rows
    for (int i = 0; i < n && aligned_src; ++i)
        aligned_src = src[i].step % 16 == 0 && src[i].step < (1ull << 24) && reinterpret_cast<uintptr_t>(src[i].data) % 16 == 0;
    const bool fs2_ok = c->fs2.ok && c->feather_variant == 1 && aligned_src;                       // the streaming frame kernel (feather / no blending)
    const bool stream_ok = fs2_ok || (c->feather_tma && (c->feather_variant == 1 || c->feather_variant == 3) && aligned_src);   // ... or its round-1 predecessor
    if ((c->crop || cfg.crop_app_fill) && !fs2_ok)
        return fail(SB_ERR_NOT_IMPL, "crop margins are applied by the streaming frame kernel only (16-byte aligned sources, default kernel variant)");
    if (c->host_maps && !((cfg.blender_kind == SB_BLEND_MULTI_BAND && c->fused && c->mb_fast && c->mb_variant == 1) ||
                          (cfg.blender_kind != SB_BLEND_MULTI_BAND && c->fused && c->feather_fast)))
        return fail(SB_ERR_NOT_IMPL, "projectors beyond plane / cylindrical / spherical run on the fused fast paths only");
    // Blender::prepare zeroes the accumulators (blenders.cpp:71-78, 227-232): only the unfused
    // (camera-by-camera) path has accumulators in HBM; the weight sums are resident either way
    if (cfg.blender_kind == SB_BLEND_MULTI_BAND && c->fused && c->mb_fast && c->mb_variant == 1) {
        const int nb = c->num_bands, W = c->dst_roi.width;
        int zero[32] = {0}, full[32];
        for (int l = 0; l <= nb; ++l) full[l] = W >> l;
        SB_TRY(mb_frame(c, s, src, zero, full, zero, full));
    } else if (cfg.blender_kind == SB_BLEND_MULTI_BAND && c->fused) {
        const int nb = c->num_bands;
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            auto &g = s.gpyr[i];
            if (blocks) SB_TRY(staged_warp_blocks(c, s, i, src[i], &g[0].v));
            else
            PROF("warp_fused", img_bytes(src[i]) + img_bytes(g[0].v),
                 launch_warp_fused(cam.proj, cam.wt, src[i], cam.ww, cam.wh, cam.left, cam.top, cam.gain, gain_on, g[0].v, st));
            for (int l = 0; l < nb; ++l)
                PROF("pyr_down", img_bytes(g[l].v) + img_bytes(g[l + 1].v), launch_pyr_down(g[l].v, g[l + 1].v, st));
        }
        // bands coarse -> fine: Laplacian on the fly, weighted sum over cameras, normalise, collapse add
        for (int l = nb; l >= 0; --l) {
            BandFusedArgs a{};
            a.n = n;
            double bytes = 0;
            for (int i = 0; i < n; ++i) {
                const Camera &cam = c->cams[i];
                BandCam &bc = a.cam[i];
                const DImage &f = s.gpyr[i][l].v;
                bc.fine = f.ptr<short>(); bc.fstep = f.step;
                if (l < nb) { bc.coarse = s.gpyr[i][l + 1].v.ptr<short>(); bc.cstep = s.gpyr[i][l + 1].v.step; }
                bc.weight = cam.w_pyr[l].v.data; bc.wstep = cam.w_pyr[l].v.step;
                bc.rx = cam.rx >> l; bc.ry = cam.ry >> l; bc.rw = f.cols; bc.rh = f.rows;
                for (int k = 0; k < 4; ++k) bc.span[k] = cam.spans[l][k];
                // only the columns inside the spans are touched: fine + coarse/4 + weights
                const double frac = std::min(1.0, (double)((bc.span[1] - bc.span[0]) + (bc.span[3] - bc.span[2])) / std::max(1, f.cols));
                bytes += frac * (img_bytes(f) * (l < nb ? 1.25 : 1.0) + img_bytes(cam.w_pyr[l].v));
            }
            const DImage &ws = c->wsum[l].v;
            a.wsum = ws.data; a.wsum_step = ws.step;
            if (l < nb) { a.coarse_r = s.acc[l + 1].v.ptr<short>(); a.coarse_r_step = s.acc[l + 1].v.step; bytes += img_bytes(s.acc[l + 1].v); }
            a.lw = s.acc[l].v.cols; a.lh = s.acc[l].v.rows;
            const bool fin = l == 0;
            if (fin) {
                a.out = s.out.v.data; a.out_step = s.out.v.step;
                a.out_mask = s.want_mask ? s.out_mask.v.ptr<uint8_t>() : nullptr; a.mask_step = s.out_mask.v.step;
                a.out_w = s.out.v.cols; a.out_h = s.out.v.rows;
                bytes += img_bytes(s.out.v) + (s.want_mask ? img_bytes(s.out_mask.v) : 0) + (double)a.out_w * a.out_h * elem_size(ws.type);
            } else {
                a.out = s.acc[l].v.data; a.out_step = s.acc[l].v.step;
                bytes += img_bytes(s.acc[l].v) + img_bytes(ws);
            }
            PROF(fin ? "band_fused_final" : "band_fused", bytes,
                 launch_band_fused(a, cfg.weight_type == SB_32F, l < nb, fin, s.out.v.type == SB_8UC3, st));
        }
    } else if ((cfg.blender_kind == SB_BLEND_FEATHER || cfg.blender_kind == SB_BLEND_NO) && c->fused && c->feather_fast &&
               c->stream_sharpness > 0.f && (cfg.blender_kind == SB_BLEND_FEATHER || stream_ok)) {
        double bytes = 0;
        for (int i = 0; i < n; ++i) {
            const auto &sp = c->cams[i].spans[0];
            // source gathered ~once; table entries (8 B) read inside the non-zero-weight column spans
            bytes += img_bytes(src[i]) + 8.0 * c->cams[i].wh * ((sp[1] - sp[0]) + (sp[3] - sp[2]));
        }
        bytes += img_bytes(s.out.v) + (s.want_mask ? img_bytes(s.out_mask.v) : 0);
        if (fs2_ok) {
            Fs2Args a{};
            fs2_static_args(c, a);
            double fs2_bytes = c->fs2.table_bytes + img_bytes(s.out.v) + (s.want_mask ? img_bytes(s.out_mask.v) : 0);
            if (blocks) for (int i = 0; i < n; ++i) fs2_bytes += c->cams[i].gain_bytes;      // the resized gain maps are read once per weighted pixel
            for (int i = 0; i < n; ++i) {
                fs2_bytes += img_bytes(src[i]);
                SB_TRY(fs2_tmaps(c, i, src[i], &a.tmap[i * FS2_NCLS]));
            }
            a.out = s.out.v.data; a.out_step = (unsigned)s.out.v.step;
            a.out_mask = s.want_mask ? s.out_mask.v.ptr<uint8_t>() : nullptr; a.mask_step = (unsigned)s.out_mask.v.step;
            a.pw = s.out.v.cols; a.ph = s.out.v.rows;
            a.src0 = src[0].ptr<uint8_t>(); a.sstep0 = (unsigned)src[0].step;
            PROF(a.no_blend ? "noblend_stream" : "feather_stream", fs2_bytes, launch_fs2(a, gain_on, s.out.v.type == SB_8UC3, c->fs2.grid, st));
        } else if (stream_ok) {
            FeatherTmaArgs a{};
            a.n = n;
            for (int i = 0; i < n; ++i) {
                const Camera &cam = c->cams[i];
                FeatherTmaCam &fc = a.cam[i];
                fc.src = src[i].ptr<uint8_t>(); fc.sstep = (unsigned)src[i].step; fc.gain = cam.gain;
                fc.dx = cam.tl.x - c->dst_roi.x; fc.dy = cam.tl.y - c->dst_roi.y;
                if (blocks) { fc.gmap = cam.gain_full.v.ptr<float>(); fc.gmstep = (unsigned)cam.gain_full.v.step; }
                fc.tiles = static_cast<const uint2 *>(cam.feather_tiles.p);
            }
            a.desc = static_cast<const uint4 *>(c->tma_desc.p);
            a.bilin_lut = static_cast<const uint2 *>(c->bilin_lut.p);
            a.sharpness = c->stream_sharpness;
            a.no_blend = cfg.blender_kind == SB_BLEND_NO;
            a.out = s.out.v.data; a.out_step = s.out.v.step;
            a.out_mask = s.want_mask ? s.out_mask.v.ptr<uint8_t>() : nullptr; a.mask_step = s.out_mask.v.step;
            a.pw = s.out.v.cols; a.ph = s.out.v.rows;
            a.tiles_x = div_up(a.pw, SB_FTT_W); a.n_tiles = a.tiles_x * div_up(a.ph, SB_FTT_H);
            PROF(a.no_blend ? "noblend_stream" : "feather_stream", bytes, launch_feather_stream(a, gain_on, s.out.v.type == SB_8UC3, c->sm_count, st));
        } else {
            FeatherFusedArgs a{};
            a.n = n;
            for (int i = 0; i < n; ++i) {
                const Camera &cam = c->cams[i];
                FeatherCam &fc = a.cam[i];
                fc.src = src[i].ptr<uint8_t>(); fc.sstep = src[i].step;
                fc.table = static_cast<const uint2 *>(cam.feather_table.p); fc.tstep = cam.feather_tstep;
                fc.ww = cam.ww; fc.wh = cam.wh; fc.dx = cam.tl.x - c->dst_roi.x; fc.dy = cam.tl.y - c->dst_roi.y;
                fc.gain = cam.gain;
                if (blocks) { fc.gmap = cam.gain_full.v.ptr<float>(); fc.gmstep = cam.gain_full.v.step; }
            }
            a.tiles_x = div_up(s.out.v.cols, SB_FT_W);
            a.sharpness = cfg.sharpness;
            a.bilin_lut = static_cast<const uint2 *>(c->bilin_lut.p);
            a.tile_cams = static_cast<const uint32_t *>(c->tile_cams.p);
            a.out = s.out.v.data; a.out_step = s.out.v.step;
            a.out_mask = s.want_mask ? s.out_mask.v.ptr<uint8_t>() : nullptr; a.mask_step = s.out_mask.v.step;
            a.pw = s.out.v.cols; a.ph = s.out.v.rows;
            PROF("feather_fused", bytes, launch_feather_fused(a, gain_on, s.out.v.type == SB_8UC3, st));
        }
    } else if (cfg.blender_kind == SB_BLEND_MULTI_BAND) {
        for (auto &acc : s.acc) PROF("zero_fill", img_bytes(acc.v), launch_set_zero(acc.v, st));
        const int nb = c->num_bands;
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            auto &g = s.gpyr[i];
            // warp + gain + convertTo(16S) + copyMakeBorder(REFLECT) -> Gaussian level 0
            // source is gathered (each pixel read ~once), output written once
            if (blocks) SB_TRY(staged_warp_blocks(c, s, i, src[i], &g[0].v));
            else
            PROF("warp_fused", img_bytes(src[i]) + img_bytes(g[0].v),
                 launch_warp_fused(cam.proj, cam.wt, src[i], cam.ww, cam.wh, cam.left, cam.top, cam.gain, gain_on, g[0].v, st));
            for (int l = 0; l < nb; ++l)
                PROF("pyr_down", img_bytes(g[l].v) + img_bytes(g[l + 1].v), launch_pyr_down(g[l].v, g[l + 1].v, st));
            int x_tl = cam.rx, y_tl = cam.ry;
            for (int l = 0; l <= nb; ++l) {
                // reads fine + coarse + weights, read-modify-write of the accumulator window
                PROF("lap_accumulate", img_bytes(g[l].v) * 3 + (l < nb ? img_bytes(g[l + 1].v) : 0) + img_bytes(cam.w_pyr[l].v),
                     launch_lap_accumulate(g[l].v, l < nb ? g[l + 1].v : none, cam.w_pyr[l].v, s.acc[l].v, none, x_tl, y_tl, st));
                x_tl /= 2; y_tl /= 2;
            }
        }
        PROF("normalize", img_bytes(s.acc[nb].v) * 2 + img_bytes(c->wsum[nb].v), launch_normalize(c->wsum[nb].v, s.acc[nb].v, st));
        for (int l = nb - 1; l >= 0; --l)
            PROF("normalize_collapse", img_bytes(s.acc[l].v) * 2 + img_bytes(s.acc[l + 1].v) + img_bytes(c->wsum[l].v),
                 launch_normalize_collapse(s.acc[l + 1].v, c->wsum[l].v, s.acc[l].v, st));
        PROF("finalize", img_bytes(s.out.v) * (1 + 6.0 / elem_size(s.out.v.type)) + img_bytes(s.out_mask.v) * (1 + elem_size(c->wsum[0].v.type)),
             launch_finalize(s.acc[0].v, c->wsum[0].v, nullptr, s.out.v, s.out_mask.v, st));
    } else {
        PROF("zero_fill", img_bytes(s.acc[0].v), launch_set_zero(s.acc[0].v, st));
        if (cfg.blender_kind == SB_BLEND_NO) PROF("zero_fill", img_bytes(s.acc_mask.v), launch_set_zero(s.acc_mask.v, st));
        for (int i = 0; i < n; ++i) {
            const Camera &cam = c->cams[i];
            if (blocks) SB_TRY(staged_warp_blocks(c, s, i, src[i], nullptr));
            else {
            SB_TRY(s.warped.create(cam.wh, cam.ww, SB_8UC3));
            PROF("warp_fused", img_bytes(src[i]) + img_bytes(s.warped.v),
                 launch_warp_fused(cam.proj, cam.wt, src[i], cam.ww, cam.wh, 0, 0, cam.gain, gain_on, s.warped.v, st));
            }
            const int dx = cam.tl.x - c->dst_roi.x, dy = cam.tl.y - c->dst_roi.y;
            if (cfg.blender_kind == SB_BLEND_FEATHER)
                PROF("feather_accumulate", img_bytes(s.warped.v) * (1 + 4.0 / 3 + 2 * 2), 
                     launch_feather_accumulate(s.warped.v, cam.feather_w.v, s.acc[0].v, none, dx, dy, st));
            else
                PROF("masked_copy", img_bytes(s.warped.v) * (1 + 2) + img_bytes(cam.mask.v) * 3,
                     launch_masked_copy(s.warped.v, cam.mask.v, s.acc[0].v, s.acc_mask.v, dx, dy, st));
        }
        if (cfg.blender_kind == SB_BLEND_FEATHER) {
            PROF("normalize", img_bytes(s.acc[0].v) * 2 + img_bytes(c->wsum[0].v), launch_normalize(c->wsum[0].v, s.acc[0].v, st));
            PROF("finalize", img_bytes(s.out.v) * (1 + 6.0 / elem_size(s.out.v.type)) + img_bytes(s.out_mask.v) * 5,
                 launch_finalize(s.acc[0].v, c->wsum[0].v, nullptr, s.out.v, s.out_mask.v, st));
        } else
            PROF("finalize", img_bytes(s.out.v) * (1 + 6.0 / elem_size(s.out.v.type)) + img_bytes(s.out_mask.v) * 2,
                 launch_finalize(s.acc[0].v, none, &s.acc_mask.v, s.out.v, s.out_mask.v, st));
    }
    return SB_OK;
}

}  // namespace

extern "C" {

int sb_compositor_create(const sb_compositor_config *cfg, int device, sb_compositor **out)
{
    if (!out) return fail(SB_ERR_ASSERT, "out is null");
    *out = nullptr;
    SB_ASSERT(cfg && cfg->n_cameras > 0 && cfg->K && cfg->R);
    SB_ASSERT(cfg->src_size.width > 0 && cfg->src_size.height > 0);
    if (cfg->warper_kind < SB_WARP_PLANE || cfg->warper_kind > SB_WARP_PLANE_PORTRAIT)
        return fail(SB_ERR_BAD_ARG, "unsupported warper kind %d", cfg->warper_kind);
    if ((cfg->undistort_map1 != nullptr) != (cfg->undistort_map2 != nullptr))
        return fail(SB_ERR_BAD_ARG, "the undistort stage needs both maps of initUndistortRectifyMap (CV_16SC2 + CV_16UC1)");
    const bool crop = cfg->crop_up != 0.f || cfg->crop_down != 0.f || cfg->crop_left != 0 || cfg->crop_right != 0;
    if (crop || cfg->crop_app_fill) {
        SB_ASSERT(cfg->crop_up >= 0.f && cfg->crop_down >= 0.f && cfg->crop_up + cfg->crop_down < 1.f && cfg->crop_left >= 0 && cfg->crop_right >= 0);
        if (cfg->blender_kind == SB_BLEND_MULTI_BAND) return fail(SB_ERR_NOT_IMPL, "crop margins apply to the Blender::NO / FeatherBlender composite");
        if (cfg->crop_app_fill && cfg->blender_kind != SB_BLEND_NO) return fail(SB_ERR_BAD_ARG, "crop_app_fill is the Blender::NO look-up composite's behaviour");
    }
    if (cfg->blender_kind != SB_BLEND_NO && cfg->blender_kind != SB_BLEND_FEATHER && cfg->blender_kind != SB_BLEND_MULTI_BAND)
        return fail(SB_ERR_BAD_ARG, "unsupported blending method");
    if (cfg->comp_kind != SB_COMP_NO && cfg->comp_kind != SB_COMP_GAIN && cfg->comp_kind != SB_COMP_GAIN_BLOCKS)
        return fail(SB_ERR_BAD_ARG, "unsupported exposure compensation method");
    SB_ASSERT(cfg->comp_kind != SB_COMP_GAIN_BLOCKS || cfg->gain_maps);
    if (cfg->blender_kind == SB_BLEND_MULTI_BAND) SB_ASSERT(cfg->weight_type == SB_32F || cfg->weight_type == SB_16S);
    SB_ASSERT(cfg->output_type == SB_8UC3 || cfg->output_type == SB_16SC3);
    SB_ASSERT(cfg->comp_kind != SB_COMP_GAIN || cfg->gains);
    SB_ASSERT(cfg->n_cameras <= SB_MAX_CAMERAS);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_compositor *c = new sb_compositor;
    c->device = device;
    c->cfg = *cfg;
    cudaError_t e = cudaStreamCreateWithFlags(&c->setup_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(SB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    int rc = setup(c);
    if (rc == SB_OK) {
        c->slots.resize(1);
        rc = make_slot(c, c->slots[0]);
    }
    // the config's pointers are not retained
    c->cfg.K = c->cfg.R = nullptr; c->cfg.gains = nullptr; c->cfg.seam_masks = nullptr; c->cfg.gain_maps = nullptr;
    c->cfg.undistort_map1 = c->cfg.undistort_map2 = nullptr;
    if (rc != SB_OK) { sb_compositor_destroy(c); return rc; }
    *out = c;
    return SB_OK;
}

void sb_compositor_destroy(sb_compositor *c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    for (size_t i = 0; i < c->slots.size(); ++i) free_slot(c->slots[i], i == 0 && c->external_stream);
    if (c->setup_stream) cudaStreamDestroy(c->setup_stream);
    delete c;
}

int sb_compositor_pano_size(const sb_compositor *c, sb_size *size)
{
    SB_ASSERT(c && size);
    size->width = c->out_rect.width; size->height = c->out_rect.height;       // (the composite's ROI, less the crop margins)
    return SB_OK;
}

int sb_compositor_camera_roi(const sb_compositor *c, int index, sb_rect *roi)
{
    SB_ASSERT(c && roi && index >= 0 && index < (int)c->cams.size());
    const Camera &cam = c->cams[index];
    roi->x = cam.tl.x; roi->y = cam.tl.y; roi->width = cam.ww; roi->height = cam.wh;
    return SB_OK;
}

int sb_compositor_set_fused(sb_compositor *c, int fused)
{
    SB_ASSERT(c);
    c->fused = fused != 0;
    if (fused >= 10) {   // 10: CV_16S band kernels / px1 feather; 11: fast paths (default); 12 / 13: fast paths with one launch per
        // pyramid level / with the multi-level launches forced; 14: fast paths with the gather warp stage
        c->feather_variant = fused == 10 ? 0 : fused == 15 ? 3 : 1; c->mb_variant = fused == 10 ? 0 : 1;   // 15: the round-1 feather streaming kernel
        c->mb_multilevel = fused == 12 ? 0 : fused == 13 ? 1 : fused == 16 ? 2 : fused == 19 ? 3 : -1;   // 16: k_mb_coarse whatever the number of frames in flight
        c->mbs_enabled = fused != 14;                        // 14: default fast paths with the gather form of the multi-band warp stage
        c->pyr_tma_enabled = fused != 18;                    // 18: ... with the gather form of pyrDown (k_mb_pyr_down_list)
        c->mbf_enabled = fused != 17;                        // 17: ... with the round-1 streaming kernel (k_mb_warp_stream) as the warp stage
    }   // test/tuning hook: 10 / 11 select the kernel variant
    return SB_OK;
}

int sb_debug_fs2_trace_dump(void) { return fs2_trace_dump(); }

int sb_compositor_kernel_plan(const sb_compositor *c)
{
    if (!c) return SB_ERR_ASSERT;
    if (c->cfg.blender_kind == SB_BLEND_MULTI_BAND) return c->mb_fast ? (c->mbs_ok ? 2 : 1) : 0;
    return c->feather_fast ? (c->fs2.ok ? 2 : c->feather_tma ? 1 : 0) : 0;
}

int sb_compositor_set_depth(sb_compositor *c, int depth)
{
    SB_ASSERT(c && depth >= 1 && depth <= 16);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    for (auto &s : c->slots) if (s.busy) return fail(SB_ERR_ASSERT, "set_depth with frames in flight");
    while ((int)c->slots.size() > depth) { free_slot(c->slots.back()); c->slots.pop_back(); }
    while ((int)c->slots.size() < depth) {
        c->slots.emplace_back();
        SB_TRY(make_slot(c, c->slots.back()));
    }
    c->next_slot = 0;
    return SB_OK;
}

// The live app's first remap (APP64:736-745): remap(frame, img3, mapEye1, mapEye2, INTER_LINEAR) - fixed-point maps, constant
// zero border - per camera; the warp then reads the 8-bit result, as in the reference (two roundings).
static int undistort_stage(sb_compositor *c, Slot &s, DImage *src)
{
    if (!c->undistort) return SB_OK;
    cudaStream_t st = s.stream;
    for (int i = 0; i < c->cfg.n_cameras; ++i) {
        const Camera &cam = c->cams[i];
        PROF("undistort", img_bytes(src[i]) + img_bytes(s.undist[i].v) + img_bytes(cam.umap1.v) + img_bytes(cam.umap2.v),
             launch_remap(src[i], s.undist[i].v, cam.umap1.v, cam.umap2.v, SB_INTER_LINEAR, SB_BORDER_CONSTANT, nullptr, st));
        src[i] = s.undist[i].v;
    }
    return SB_OK;
}

// One frame set on slot `s`: host->device staging of host sources, the frame kernels, the copy of the panorama (and mask)
// into the caller's buffers - everything asynchronous on the slot's stream.
static int enqueue_on_slot(sb_compositor *c, Slot &s, const sb_image *srcs, sb_image *pano, sb_image *pano_mask, bool timing)
{
    const int n = c->cfg.n_cameras;
    if (timing) SB_CUDA(cudaEventRecord(s.ev_start, s.stream));
    DImage src_local[SB_MAX_CAMERAS];
    std::vector<DImage> src_heap;
    DImage *srcp = src_local;
    if (n > SB_MAX_CAMERAS) { src_heap.resize(n); srcp = src_heap.data(); }
    // A frame set delivered as ONE host block (camera i+1 starts where camera i ends, rows packed) crosses PCIe as a
    // single DMA: n separate 6 MB copies cost ~20 % of the link's throughput in per-copy overhead.
    bool one_block = n > 1;
    const size_t row_bytes = (size_t)c->cfg.src_size.width * 3, img_bytes_ = row_bytes * c->cfg.src_size.height;
    for (int i = 0; i < n; ++i) {
        SB_ASSERT(srcs[i].type == SB_8UC3 && srcs[i].rows == c->cfg.src_size.height && srcs[i].cols == c->cfg.src_size.width);
        one_block = one_block && srcs[i].device < 0 && srcs[i].data && srcs[i].step == row_bytes && row_bytes % 16 == 0 &&
                    (i == 0 || static_cast<const char *>(srcs[i].data) == static_cast<const char *>(srcs[i - 1].data) + img_bytes_);
    }
    if (one_block) {
        SB_TRY(s.src_all.ensure(img_bytes_ * n + 16));
        SB_CUDA(cudaMemcpyAsync(s.src_all.p, srcs[0].data, img_bytes_ * n, cudaMemcpyHostToDevice, s.stream));
        for (int i = 0; i < n; ++i) {
            srcp[i].data = static_cast<char *>(s.src_all.p) + img_bytes_ * i;
            srcp[i].rows = srcs[i].rows; srcp[i].cols = srcs[i].cols; srcp[i].type = SB_8UC3; srcp[i].step = row_bytes;
        }
    } else {
        for (int i = 0; i < n; ++i) SB_TRY(to_device(srcs[i], s.src[i], s.stream, &srcp[i]));
    }
    s.want_mask = pano_mask != nullptr;
    SB_TRY(undistort_stage(c, s, srcp));
    const std::vector<DImage> src(srcp, srcp + n);
    // A caller's panorama that already lives on this device is written in place by the frame's last kernel (no 2 x 20 MB
    // device-to-device copy behind it) when its pitch and alignment suit that kernel's vector stores.
    auto in_place = [&](const sb_image *u, const DImage &own, size_t align) {
        return u && u->data && u->device == c->device && c->strip_world == 1 && u->rows == own.rows && u->cols == own.cols && u->type == own.type &&
               u->step % align == 0 && reinterpret_cast<uintptr_t>(u->data) % 16 == 0;
    };
    const DImage own_out = s.out.v, own_mask = s.out_mask.v;
    const bool direct = in_place(pano, own_out, 8), direct_mask = direct && in_place(pano_mask, own_mask, 4);
    if (direct) { s.out.v.data = pano->data; s.out.v.step = pano->step; }
    if (direct_mask) { s.out_mask.v.data = pano_mask->data; s.out_mask.v.step = pano_mask->step; }
    const int rc = run_frame(c, s, src);
    s.out.v = own_out; s.out_mask.v = own_mask;
    SB_TRY(rc);
    if (!pano->data) lend(s.out.v, c->device, pano);
    else if (!direct) SB_TRY(from_device(s.out.v, pano, s.stream));
    if (pano_mask) {
        if (!pano_mask->data) lend(s.out_mask.v, c->device, pano_mask);
        else if (!direct_mask) SB_TRY(from_device(s.out_mask.v, pano_mask, s.stream));
    }
    if (timing) SB_CUDA(cudaEventRecord(s.ev_stop, s.stream));
    return SB_OK;
}

int sb_compositor_enqueue(sb_compositor *c, const sb_image *srcs, sb_image *pano, sb_image *pano_mask, int *slot)
{
    SB_ASSERT(c && srcs && pano && slot);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    const int si = c->next_slot;
    Slot &s = c->slots[si];
    if (s.busy) return fail(SB_ERR_ASSERT, "slot %d still in flight: call sb_compositor_wait first", si);
    SB_TRY(enqueue_on_slot(c, s, srcs, pano, pano_mask, true));
    s.busy = true;
    *slot = si;
    c->next_slot = (si + 1) % (int)c->slots.size();
    return SB_OK;
}

// ---- batches: one lap of a video pipeline's buffer ring as ONE host call -------------------------------------------------
// A video pipeline cycles through a ring of input and output buffers.  At 20-25 us of device time per frame set the per-frame
// host work of enqueue / wait (two C calls, event records, a launch) and, worse, the launch / ramp-up / tail of one kernel per
// frame set are what limit a GPU (round-1 SCALE: 0.87 efficiency at N=8 from host launch issue alone).  A batch runs a lap as
//   PERSISTENT  one launch of the frame kernel that walks all the lap's frame sets (feather / no blending, sources and
//               panoramas resident on the device): no kernel boundary between frame sets, the tile pipeline never drains;
//   STREAMS     the frame sets enqueued back to back on the slots' streams by one C call (host buffers, multi-band);
//   GRAPH       the same recorded as a CUDA graph (SB_BATCH_GRAPH=1; measured slower on B200: a dependent kernel node
//               starts ~14 us after its predecessor ends).
struct sb_batch {
    sb_compositor *c = nullptr;
    enum Mode { STREAMS, PERSISTENT, GRAPH } mode = STREAMS;
    int n_frames = 0;
    std::vector<sb_image> srcs, panos, masks;     // STREAMS: the lap's argument blocks
    DevBuf frames;                                // PERSISTENT: Fs2Frame[n_frames]
    Fs2Args args{};
    bool out8 = true;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int kernel_nodes = 0;        // kernel launches one replay performs (sb_kernel_launch_count)
    bool launched = false;
};

// PERSISTENT applies when the frame kernel can read every source and write every panorama in place
static bool batch_can_persist(sb_compositor *c, int n_frames, const sb_image *srcs, const sb_image *panos, const sb_image *masks)
{
    const sb_compositor_config &cfg = c->cfg;
    if (!(cfg.blender_kind == SB_BLEND_FEATHER || cfg.blender_kind == SB_BLEND_NO) || !c->fused || !c->feather_fast || !c->fs2.ok ||
        c->feather_variant != 1 || !(c->stream_sharpness > 0.f) || cfg.n_cameras > FS2_TMAP_CAMS || c->undistort)
        return false;
    const int n = cfg.n_cameras;
    const size_t px = (size_t)elem_size(cfg.output_type);
    for (int f = 0; f < n_frames; ++f) {
        for (int i = 0; i < n; ++i) {
            const sb_image &im = srcs[(size_t)f * n + i];
            if (im.device != c->device || !im.data || im.step % 16 || im.step >= (1ull << 24) || reinterpret_cast<uintptr_t>(im.data) % 16) return false;
        }
        const sb_image &p = panos[f];
        if (p.device != c->device || !p.data || p.step != panos[0].step || p.step % (cfg.output_type == SB_8UC3 ? 4 : 8) ||
            reinterpret_cast<uintptr_t>(p.data) % 8 || p.rows != c->out_rect.height || p.cols != c->out_rect.width ||
            p.type != cfg.output_type || (unsigned long long)p.rows * p.step >= (1ull << 32) || p.step < p.cols * px)
            return false;
        if (masks) {
            const sb_image &m = masks[f];
            if (m.device != c->device || !m.data || m.step != masks[0].step || m.step % 4 || reinterpret_cast<uintptr_t>(m.data) % 4 ||
                m.rows != p.rows || m.cols != p.cols || m.type != SB_8UC1)
                return false;
        }
    }
    return true;
}

int sb_compositor_batch_create(sb_compositor *c, int n_frames, const sb_image *srcs, sb_image *panos, sb_image *pano_masks, sb_batch **out)
{
    SB_ASSERT(c && srcs && panos && out && n_frames > 0);
    *out = nullptr;
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    if (c->external_stream) return fail(SB_ERR_NOT_IMPL, "batches run on the handle's own streams");
    for (auto &s : c->slots) if (s.busy) return fail(SB_ERR_ASSERT, "batch_create with frames in flight");
    const int n = c->cfg.n_cameras, depth = (int)c->slots.size();
    for (int f = 0; f < n_frames; ++f)
        for (int i = 0; i < n; ++i) {
            const sb_image &im = srcs[(size_t)f * n + i];
            SB_ASSERT(im.type == SB_8UC3 && im.rows == c->cfg.src_size.height && im.cols == c->cfg.src_size.width);
        }
    sb_batch *b = new sb_batch();
    b->c = c; b->n_frames = n_frames;
    int rc = SB_OK;
    auto cuda_ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == SB_OK) rc = fail(SB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    // (a recorded graph would replay k_mb_coarse with the counter values of the recording: multi-band laps are never graphs)
    static const bool graph_env = getenv("SB_BATCH_GRAPH") != nullptr && atoi(getenv("SB_BATCH_GRAPH")) != 0;
    const bool want_graph = graph_env && c->cfg.blender_kind != SB_BLEND_MULTI_BAND;
    if (!want_graph && batch_can_persist(c, n_frames, srcs, panos, pano_masks)) {
        // ---- one launch for the whole lap
        b->mode = sb_batch::PERSISTENT;
        std::vector<Fs2Frame> h(n_frames);
        for (int f = 0; f < n_frames && rc == SB_OK; ++f) {
            std::memset(&h[f], 0, sizeof(Fs2Frame));
            for (int i = 0; i < n && rc == SB_OK; ++i) {
                const sb_image &im = srcs[(size_t)f * n + i];
                DImage d; d.data = im.data; d.rows = im.rows; d.cols = im.cols; d.type = im.type; d.step = im.step;
                rc = fs2_tmaps(c, i, d, &h[f].tmap[i * FS2_NCLS]);
            }
            h[f].out = panos[f].data;
            h[f].out_mask = pano_masks ? static_cast<uint8_t *>(pano_masks[f].data) : nullptr;
            h[f].src0 = static_cast<const uint8_t *>(srcs[(size_t)f * n].data); h[f].sstep0 = (unsigned)srcs[(size_t)f * n].step;
        }
        if (rc == SB_OK) rc = b->frames.ensure(sizeof(Fs2Frame) * (size_t)n_frames);
        if (rc == SB_OK) cuda_ok(cudaMemcpy(b->frames.p, h.data(), sizeof(Fs2Frame) * (size_t)n_frames, cudaMemcpyHostToDevice), "cudaMemcpy");
        fs2_static_args(c, b->args);
        b->args.frames = static_cast<const Fs2Frame *>(b->frames.p);
        b->args.n_frames = n_frames;
        b->args.out_step = (unsigned)panos[0].step;
        b->args.mask_step = pano_masks ? (unsigned)pano_masks[0].step : 0u;
        b->args.out_mask = pano_masks ? static_cast<uint8_t *>(pano_masks[0].data) : nullptr;       // (non-null = masks wanted)
        b->args.out = panos[0].data;
        b->args.pw = c->out_rect.width; b->args.ph = c->out_rect.height;
        b->out8 = c->cfg.output_type == SB_8UC3;
        b->kernel_nodes = 1;
    } else {
        b->mode = want_graph ? sb_batch::GRAPH : sb_batch::STREAMS;
        b->srcs.assign(srcs, srcs + (size_t)n_frames * n);
        b->panos.assign(panos, panos + n_frames);
        if (pano_masks) b->masks.assign(pano_masks, pano_masks + n_frames);
        // eager pass first: every buffer the lap needs is allocated, every tensor map encoded, every kernel configured;
        // it also fills in the lent panoramas (data == NULL) of the caller's array
        const uint64_t l0 = g_launches.load();
        for (int f = 0; f < n_frames && rc == SB_OK; ++f)
            rc = enqueue_on_slot(c, c->slots[f % depth], srcs + (size_t)f * n, &panos[f], pano_masks ? &pano_masks[f] : nullptr, false);
        for (auto &s : c->slots) cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
        b->kernel_nodes = (int)(g_launches.load() - l0);
        if (b->mode == sb_batch::GRAPH && rc == SB_OK) {
            cudaStream_t origin = c->setup_stream;
            cudaEvent_t fork = nullptr;
            std::vector<cudaEvent_t> joins(depth, nullptr);
            cuda_ok(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "cudaEventCreate");
            for (auto &j : joins) cuda_ok(cudaEventCreateWithFlags(&j, cudaEventDisableTiming), "cudaEventCreate");
            if (rc == SB_OK && cuda_ok(cudaStreamBeginCapture(origin, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture")) {
                cuda_ok(cudaEventRecord(fork, origin), "fork");
                for (auto &s : c->slots) cuda_ok(cudaStreamWaitEvent(s.stream, fork, 0), "fork wait");
                for (int f = 0; f < n_frames && rc == SB_OK; ++f) {
                    sb_image p = b->panos[f], m = pano_masks ? b->masks[f] : sb_image{};
                    rc = enqueue_on_slot(c, c->slots[f % depth], b->srcs.data() + (size_t)f * n, &p, pano_masks ? &m : nullptr, false);
                }
                for (int k = 0; k < depth; ++k) {
                    cuda_ok(cudaEventRecord(joins[k], c->slots[k].stream), "join");
                    cuda_ok(cudaStreamWaitEvent(origin, joins[k], 0), "join wait");
                }
                cuda_ok(cudaStreamEndCapture(origin, &b->graph), "cudaStreamEndCapture");      // (always ends the capture, also after an error)
                if (rc == SB_OK) cuda_ok(cudaGraphInstantiate(&b->exec, b->graph, 0), "cudaGraphInstantiate");
            }
            if (fork) cudaEventDestroy(fork);
            for (auto j : joins) if (j) cudaEventDestroy(j);
        }
    }
    if (rc == SB_OK) { cuda_ok(cudaEventCreate(&b->e0), "cudaEventCreate"); cuda_ok(cudaEventCreate(&b->e1), "cudaEventCreate"); }
    if (rc != SB_OK) { sb_batch_destroy(b); return rc; }
    *out = b;
    return SB_OK;
}

int sb_batch_launch(sb_batch *b)
{
    SB_ASSERT(b);
    sb_compositor *c = b->c;
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    cudaStream_t origin = c->setup_stream;
    const int n = c->cfg.n_cameras, depth = (int)c->slots.size();
    if (b->mode == sb_batch::STREAMS) {
        // the lap on the slots' streams, bracketed on the origin stream: fork (slots wait for e0), frames, join (e1 waits for slots)
        SB_CUDA(cudaEventRecord(b->e0, origin));
        for (auto &s : c->slots) SB_CUDA(cudaStreamWaitEvent(s.stream, b->e0, 0));
        for (int f = 0; f < b->n_frames; ++f) {
            sb_image p = b->panos[f], m = b->masks.empty() ? sb_image{} : b->masks[f];
            SB_TRY(enqueue_on_slot(c, c->slots[f % depth], b->srcs.data() + (size_t)f * n, &p, b->masks.empty() ? nullptr : &m, false));
        }
        for (auto &s : c->slots) {
            SB_CUDA(cudaEventRecord(s.ev_stop, s.stream));
            SB_CUDA(cudaStreamWaitEvent(origin, s.ev_stop, 0));
        }
        SB_CUDA(cudaEventRecord(b->e1, origin));
    } else if (b->mode == sb_batch::PERSISTENT) {
        SB_CUDA(cudaEventRecord(b->e0, origin));
        SB_TRY(launch_fs2(b->args, c->cfg.comp_kind != SB_COMP_NO, b->out8, c->fs2.grid, origin));
        SB_CUDA(cudaEventRecord(b->e1, origin));
    } else {
        SB_ASSERT(b->exec);
        SB_CUDA(cudaEventRecord(b->e0, origin));
        SB_CUDA(cudaGraphLaunch(b->exec, origin));
        g_launches.fetch_add((uint64_t)b->kernel_nodes, std::memory_order_relaxed);
        SB_CUDA(cudaEventRecord(b->e1, origin));
    }
    b->launched = true;
    return SB_OK;
}

int sb_batch_mode(const sb_batch *b) { return b ? (int)b->mode : -1; }

int sb_batch_wait(sb_batch *b)
{
    SB_ASSERT(b);
    if (!b->launched) return SB_OK;
    DeviceGuard g(b->c->device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_CUDA(cudaEventSynchronize(b->e1));
    return SB_OK;
}

int sb_batch_last_gpu_ms(sb_batch *b, float *ms)
{
    SB_ASSERT(b && ms && b->launched);
    DeviceGuard g(b->c->device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_CUDA(cudaEventSynchronize(b->e1));
    SB_CUDA(cudaEventElapsedTime(ms, b->e0, b->e1));
    return SB_OK;
}

int sb_batch_frames(const sb_batch *b) { return b ? b->n_frames : 0; }

void sb_batch_destroy(sb_batch *b)
{
    if (!b) return;
    DeviceGuard g(b->c->device);
    if (b->launched && b->e1) cudaEventSynchronize(b->e1);
    if (b->exec) cudaGraphExecDestroy(b->exec);
    if (b->graph) cudaGraphDestroy(b->graph);
    if (b->e0) cudaEventDestroy(b->e0);
    if (b->e1) cudaEventDestroy(b->e1);
    delete b;
}

int sb_compositor_wait(sb_compositor *c, int slot)
{
    SB_ASSERT(c && slot >= 0 && slot < (int)c->slots.size());
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[slot];
    if (!s.busy) return SB_OK;
    SB_CUDA(cudaEventSynchronize(s.ev_stop));
    s.busy = false;
    return SB_OK;
}

int sb_compositor_compose(sb_compositor *c, const sb_image *srcs, sb_image *pano, sb_image *pano_mask)
{
    int slot = -1;
    SB_TRY(sb_compositor_enqueue(c, srcs, pano, pano_mask, &slot));
    return sb_compositor_wait(c, slot);
}

// Device-side timing of a whole region spanning all slots: mark(0) / mark(1) record an event on the
// setup stream after it has been made to wait for every slot's last enqueued work.
int sb_compositor_mark(sb_compositor *c, int which)
{
    SB_ASSERT(c && (which == 0 || which == 1));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    if (!c->marks[which]) SB_CUDA(cudaEventCreate(&c->marks[which]));
    for (auto &s : c->slots) {          // the mark follows everything enqueued so far on every slot
        cudaEvent_t tail;
        SB_CUDA(cudaEventCreateWithFlags(&tail, cudaEventDisableTiming));
        SB_CUDA(cudaEventRecord(tail, s.stream));
        SB_CUDA(cudaStreamWaitEvent(c->setup_stream, tail, 0));
        SB_CUDA(cudaEventDestroy(tail));
    }
    SB_CUDA(cudaEventRecord(c->marks[which], c->setup_stream));
    if (which == 0)                     // and nothing enqueued later may start before the start mark
        for (auto &s : c->slots) SB_CUDA(cudaStreamWaitEvent(s.stream, c->marks[0], 0));
    return SB_OK;
}

int sb_compositor_marked_ms(sb_compositor *c, float *ms)
{
    SB_ASSERT(c && ms && c->marks[0] && c->marks[1]);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_CUDA(cudaEventSynchronize(c->marks[1]));
    SB_CUDA(cudaEventElapsedTime(ms, c->marks[0], c->marks[1]));
    return SB_OK;
}

// One frame on slot 0 with every kernel bracketed by CUDA events on the launching stream.
// Writes a JSON array [{"name":..., "ms":..., "bytes":...}, ...] (one entry per launch) to buf.
int sb_compositor_profile_frame(sb_compositor *c, const sb_image *srcs, char *buf, size_t cap)
{
    SB_ASSERT(c && srcs && buf && cap > 2);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    if (s.busy) return fail(SB_ERR_ASSERT, "profile_frame with slot 0 in flight");
    const int n = c->cfg.n_cameras;
    std::vector<DImage> src(n);
    for (int i = 0; i < n; ++i) {
        SB_ASSERT(srcs[i].type == SB_8UC3 && srcs[i].rows == c->cfg.src_size.height && srcs[i].cols == c->cfg.src_size.width);
        SB_TRY(to_device(srcs[i], s.src[i], s.stream, &src[i]));
    }
    std::vector<ProfRec> recs;
    s.prof = &recs;
    s.want_mask = false;
    int rc = undistort_stage(c, s, src.data());
    if (rc == SB_OK) rc = run_frame(c, s, src);
    s.prof = nullptr;
    cudaError_t e = cudaStreamSynchronize(s.stream);
    std::string out = "[";
    for (size_t i = 0; i < recs.size(); ++i) {
        float ms = 0.f;
        if (rc == SB_OK && e == cudaSuccess) cudaEventElapsedTime(&ms, recs[i].e0, recs[i].e1);
        cudaEventDestroy(recs[i].e0); cudaEventDestroy(recs[i].e1);
        char item[160];
        snprintf(item, sizeof item, "%s{\"name\":\"%s\",\"ms\":%.6f,\"bytes\":%.0f}", i ? "," : "", recs[i].name, ms, recs[i].bytes);
        out += item;
    }
    out += "]";
    SB_TRY(rc);
    if (e != cudaSuccess) return fail(SB_ERR_CUDA, "profile_frame: %s", cudaGetErrorString(e));
    if (out.size() + 1 > cap) return fail(SB_ERR_ASSERT, "profile buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return SB_OK;
}

int sb_compositor_last_gpu_ms(sb_compositor *c, int slot, float *ms)
{
    SB_ASSERT(c && ms && slot >= 0 && slot < (int)c->slots.size());
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[slot];
    SB_CUDA(cudaEventSynchronize(s.ev_stop));
    SB_CUDA(cudaEventElapsedTime(ms, s.ev_start, s.ev_stop));
    return SB_OK;
}


// ------------------------------------------------------------------------------------ latency ("strip") mode
// one very wide panorama cut into column strips, one per rank.  Every stage of the multi-band
// fast path runs on this rank's columns only; between stages the host exchanges the pyramid halo columns with
// the two neighbouring ranks (NCCL send/recv): 2 columns of every camera's Gaussian level l (the pyrDown taps
// of level l+1 and the pyrUp taps of the Laplacian of level l-1) and 1 column of every restored band (the
// pyrUp taps of the collapse).  Buffers are full-size on every rank, so halo columns live at their true
// coordinates and the kernels are the single-GPU kernels restricted to a column range.

namespace {

struct HaloSeg {
    char *base;          // first byte of the buffer
    size_t step;
    int esz, rows;
    int send_col, recv_col, ncols_send, ncols_recv;
};

void strip_bounds(const sb_compositor *c, int rank, int world, int *x0, int *x1)
{
    const int m = 1 << c->num_bands, units = c->dst_roi.width / m;
    *x0 = (int)((long long)units * rank / world) * m;
    *x1 = (int)((long long)units * (rank + 1) / world) * m;
}

// segments of one halo message, in camera order (both neighbours derive the same list from the calibration)
int halo_segments(sb_compositor *c, int what, int level, int side, std::vector<HaloSeg> *out)
{
    out->clear();
    SB_ASSERT(c->strip_world > 1);
    SB_ASSERT(what == SB_HALO_GAUSS || what == SB_HALO_RESTORED);
    SB_ASSERT(side == SB_SIDE_LEFT || side == SB_SIDE_RIGHT);
    const int nb = c->num_bands;
    const int B = side == SB_SIDE_LEFT ? c->strip_x0 : c->strip_x1;
    if (B <= 0 || B >= c->dst_roi.width) return SB_OK;      // panorama edge: no neighbour on this side
    const int Bl = B >> level;
    Slot &s = c->slots[0];
    if (what == SB_HALO_GAUSS) {
        SB_ASSERT(level >= 0 && level <= nb);
        for (int i = 0; i < c->cfg.n_cameras; ++i) {
            RawImage &g = s.grgbx[i][level];
            const int rx = c->cams[i].rx >> level, rw = g.cols;
            if (!(rx < Bl && Bl < rx + rw)) continue;       // the feed rect does not straddle the boundary
            const int lo = std::max(Bl - 2, rx) - rx, mid = Bl - rx, hi = std::min(Bl + 2, rx + rw) - rx;
            HaloSeg h{static_cast<char *>(g.buf.p), g.step, 4, g.rows, 0, 0, 0, 0};
            if (side == SB_SIDE_RIGHT) { h.send_col = lo; h.ncols_send = mid - lo; h.recv_col = mid; h.ncols_recv = hi - mid; }
            else                       { h.send_col = mid; h.ncols_send = hi - mid; h.recv_col = lo; h.ncols_recv = mid - lo; }
            out->push_back(h);
        }
    } else {
        SB_ASSERT(level >= 1 && level <= nb);
        RawImage &r = s.rband[level];
        HaloSeg h{static_cast<char *>(r.buf.p), r.step, 8, r.rows, 0, 0, 1, 1};
        if (side == SB_SIDE_RIGHT) { h.send_col = Bl - 1; h.recv_col = Bl; }
        else                       { h.send_col = Bl; h.recv_col = Bl - 1; }
        out->push_back(h);
    }
    return SB_OK;
}

// Column ranges per level.  Exchange mode: exactly this rank's columns (the neighbours supply the halos).
// Recompute mode: the ranges grow towards the coarse levels' taps so that no halo is ever needed —
//   band l is produced on own +- e_l with e_0 = 0, e_l = ceil(e_{l-1} / 2) + 1 (collapse pyrUp taps),
//   Gaussian level l on the hull of band l, the Laplacian pyrUp taps of band l-1 and the pyrDown taps of level l+1
//   (about 2 * (2^num_bands) + ... = 94 level-0 columns per side for 5 bands).
void strip_plan(sb_compositor *c)
{
    const int nb = c->num_bands;
    c->g_lo.assign(nb + 1, 0); c->g_hi.assign(nb + 1, 0); c->b_lo.assign(nb + 1, 0); c->b_hi.assign(nb + 1, 0);
    int e = 0;
    for (int l = 0; l <= nb; ++l) {
        const int lw = c->dst_roi.width >> l;
        if (l > 0) e = (e + 1) / 2 + 1;
        const int m = c->strip_recompute ? e : 0;
        c->b_lo[l] = std::max(0, (c->strip_x0 >> l) - m);
        c->b_hi[l] = std::min(lw, (c->strip_x1 >> l) + m);
    }
    for (int l = nb; l >= 0; --l) {
        const int lw = c->dst_roi.width >> l;
        int lo = c->b_lo[l], hi = c->b_hi[l];
        if (c->strip_recompute) {
            if (l >= 1) {      // Laplacian of band l-1: pyrUp taps of this level
                lo = std::min(lo, (c->b_lo[l - 1] >> 1) - 1);
                hi = std::max(hi, ((c->b_hi[l - 1] - 1) >> 1) + 2);
            }
            if (l < nb) {      // pyrDown taps of level l+1
                lo = std::min(lo, 2 * c->g_lo[l + 1] - 2);
                hi = std::max(hi, 2 * c->g_hi[l + 1] + 1);
            }
        }
        c->g_lo[l] = std::max(0, lo); c->g_hi[l] = std::min(lw, hi);
    }
}

// ---- halo exchange without the host in the loop -----------------------------------------------------------------------
// Each rank owns, per side, one receive area in its own HBM (a 128-byte flag block + one region per exchange step) that the
// neighbour on that side maps (CUDA IPC between processes, the plain pointer inside one process).  An exchange step is two
// small kernels on the rank's stream: k_halo_push copies this rank's edge columns STRAIGHT into both neighbours' receive areas
// over NVLink and, when its last block is done, stores the step's sequence number into their flag words (release at system
// scope); k_halo_pull waits for the sequence number in its own flag words (acquire) and moves the received columns into place.
// No packing buffers, no collective library call, no host synchronisation: the whole frame is enqueued at once.
struct HaloCopy {                     // one column run of one buffer
    char *base; unsigned long long step;
    int esz, rows, col, ncols;
    unsigned long long off;           // byte offset inside the step's region of the receive area
    int side;
};
constexpr int SB_HALO_MAX_COPIES = 2 * (SB_MAX_CAMERAS + 1);
struct HaloXferArgs {
    int n;
    HaloCopy c[SB_HALO_MAX_COPIES];
    char *area[2];                    // push: the neighbours' receive areas (peer memory); pull: this rank's own
    unsigned long long region[2];     // byte offset of this step's region inside the area of each side
    unsigned long long seq;           // sequence number of this step (monotonic over steps and frames)
    unsigned *done;                   // push: block counter for "last block signals"
};
__global__ void __launch_bounds__(128) k_halo_push(const __grid_constant__ HaloXferArgs a)
{
    for (int k = 0; k < a.n; ++k) {
        const HaloCopy &h = a.c[k];
        char *dst = a.area[h.side] + 128 + a.region[h.side] + h.off;
        const int row_bytes = h.ncols * h.esz, words = row_bytes / 4;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.rows * words; i += gridDim.x * blockDim.x) {
            const int r = i / words, w = i - r * words;
            reinterpret_cast<uint32_t *>(dst + (size_t)r * row_bytes)[w] =
                reinterpret_cast<const uint32_t *>(h.base + (size_t)r * h.step + (size_t)h.col * h.esz)[w];
        }
    }
    __threadfence_system();                                 // this block's columns before its count
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(a.done, 1u) == gridDim.x - 1) {
        *a.done = 0u;                                       // (the next push on this stream starts from zero)
        __threadfence_system();
        for (int side = 0; side < 2; ++side)
            if (a.area[side]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.area[side]), "l"(a.seq) : "memory");
    }
}
__global__ void __launch_bounds__(128) k_halo_pull(const __grid_constant__ HaloXferArgs a)
{
    if (threadIdx.x == 0)
        for (int side = 0; side < 2; ++side) {
            if (!a.area[side]) continue;
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.area[side]) : "memory");
                if (v < a.seq) __nanosleep(200);
            } while (v < a.seq);
        }
    __syncthreads();
    for (int k = 0; k < a.n; ++k) {
        const HaloCopy &h = a.c[k];
        const char *src = a.area[h.side] + 128 + a.region[h.side] + h.off;
        const int row_bytes = h.ncols * h.esz, words = row_bytes / 4;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.rows * words; i += gridDim.x * blockDim.x) {
            const int r = i / words, w = i - r * words;
            reinterpret_cast<uint32_t *>(h.base + (size_t)r * h.step + (size_t)h.col * h.esz)[w] =
                *reinterpret_cast<const volatile uint32_t *>(src + (size_t)r * row_bytes + (size_t)w * 4);
        }
    }
}

int strip_ready(sb_compositor *c)
{
    SB_ASSERT(c);
    if (c->strip_world <= 1) return fail(SB_ERR_ASSERT, "strip mode is not enabled: call sb_compositor_set_strip first");
    return SB_OK;
}

}  // namespace

int sb_compositor_set_strip(sb_compositor *c, int rank, int world)
{
    SB_ASSERT(c && world >= 1 && rank >= 0 && rank < world);
    if (c->cfg.blender_kind != SB_BLEND_MULTI_BAND || !c->mb_fast)
        return fail(SB_ERR_NOT_IMPL, "strip mode is implemented for the multi-band fast path (MultiBandBlender, sources <= 4096 px)");
    if (c->undistort) return fail(SB_ERR_NOT_IMPL, "strip mode does not run the undistort stage");
    const int m = 1 << c->num_bands;
    if (world > 1 && c->dst_roi.width / m / world < 2)
        return fail(SB_ERR_ASSERT, "panorama too narrow for %d strips: every strip needs at least 2 * 2^num_bands columns", world);
    for (auto &s : c->slots) if (s.busy) return fail(SB_ERR_ASSERT, "set_strip with frames in flight");
    c->strip_rank = rank; c->strip_world = world;
    strip_bounds(c, rank, world, &c->strip_x0, &c->strip_x1);
    c->fused = true; c->mb_variant = 1;
    strip_plan(c);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return build_mbs_schedule(c, c->g_lo[0], c->g_hi[0], c->setup_stream);
}

int sb_compositor_set_strip_halo(sb_compositor *c, int recompute)
{
    SB_ASSERT(c);
    for (auto &s : c->slots) if (s.busy) return fail(SB_ERR_ASSERT, "set_strip_halo with frames in flight");
    c->strip_recompute = recompute != 0;
    strip_plan(c);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return build_mbs_schedule(c, c->g_lo[0], c->g_hi[0], c->setup_stream);
}

int sb_compositor_strip_compose(sb_compositor *c, const sb_image *srcs)
{
    SB_TRY(strip_ready(c));
    if (!c->strip_recompute) return fail(SB_ERR_ASSERT, "strip_compose needs the recompute halo mode (exchange mode is driven stage by stage)");
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    const int n = c->cfg.n_cameras;
    c->strip_src.resize(n);
    for (int i = 0; i < n; ++i) {
        SB_ASSERT(srcs[i].type == SB_8UC3 && srcs[i].rows == c->cfg.src_size.height && srcs[i].cols == c->cfg.src_size.width);
        SB_TRY(to_device(srcs[i], s.src[i], s.stream, &c->strip_src[i]));
    }
    s.want_mask = true;
    return mb_frame(c, s, c->strip_src, c->g_lo.data(), c->g_hi.data(), c->b_lo.data(), c->b_hi.data());
}

int sb_compositor_strip_range(const sb_compositor *c, int rank, int world, int *x0, int *x1)
{
    SB_ASSERT(c && x0 && x1 && world >= 1 && rank >= 0 && rank < world);
    strip_bounds(c, rank, world, x0, x1);
    *x0 = std::min(*x0, c->dst_roi_final.width);
    *x1 = std::min(*x1, c->dst_roi_final.width);
    return SB_OK;
}

int sb_compositor_set_stream(sb_compositor *c, void *cuda_stream)
{
    SB_ASSERT(c && !c->slots.empty());
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    if (s.busy) return fail(SB_ERR_ASSERT, "set_stream with a frame in flight");
    if (s.stream && !c->external_stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    s.stream = static_cast<cudaStream_t>(cuda_stream);
    c->external_stream = true;
    return SB_OK;
}

int sb_compositor_strip_halo_bytes(sb_compositor *c, int what, int level, int side, size_t *send_bytes, size_t *recv_bytes)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(send_bytes && recv_bytes);
    std::vector<HaloSeg> segs;
    SB_TRY(halo_segments(c, what, level, side, &segs));
    *send_bytes = *recv_bytes = 0;
    for (const HaloSeg &h : segs) {
        *send_bytes += (size_t)h.rows * h.ncols_send * h.esz;
        *recv_bytes += (size_t)h.rows * h.ncols_recv * h.esz;
    }
    return SB_OK;
}

int sb_compositor_strip_pack(sb_compositor *c, int what, int level, int side, void *device_buf)
{
    SB_TRY(strip_ready(c));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<HaloSeg> segs;
    SB_TRY(halo_segments(c, what, level, side, &segs));
    char *dst = static_cast<char *>(device_buf);
    for (const HaloSeg &h : segs) {
        const size_t w = (size_t)h.ncols_send * h.esz;
        if (w == 0) continue;
        SB_ASSERT(device_buf);
        SB_CUDA(cudaMemcpy2DAsync(dst, w, h.base + (size_t)h.send_col * h.esz, h.step, w, h.rows, cudaMemcpyDeviceToDevice, c->slots[0].stream));
        dst += w * h.rows;
    }
    return SB_OK;
}

int sb_compositor_strip_unpack(sb_compositor *c, int what, int level, int side, const void *device_buf)
{
    SB_TRY(strip_ready(c));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<HaloSeg> segs;
    SB_TRY(halo_segments(c, what, level, side, &segs));
    const char *src = static_cast<const char *>(device_buf);
    for (const HaloSeg &h : segs) {
        const size_t w = (size_t)h.ncols_recv * h.esz;
        if (w == 0) continue;
        SB_ASSERT(device_buf);
        SB_CUDA(cudaMemcpy2DAsync(h.base + (size_t)h.recv_col * h.esz, h.step, src, w, w, h.rows, cudaMemcpyDeviceToDevice, c->slots[0].stream));
        src += w * h.rows;
    }
    return SB_OK;
}

int sb_compositor_strip_warp(sb_compositor *c, const sb_image *srcs)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(srcs);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    const int n = c->cfg.n_cameras;
    c->strip_src.resize(n);
    for (int i = 0; i < n; ++i) {
        SB_ASSERT(srcs[i].type == SB_8UC3 && srcs[i].rows == c->cfg.src_size.height && srcs[i].cols == c->cfg.src_size.width);
        SB_TRY(to_device(srcs[i], s.src[i], s.stream, &c->strip_src[i]));
    }
    s.want_mask = true;
    return mb_warp_stage(c, s, c->strip_src, c->g_lo[0], c->g_hi[0]);
}

int sb_compositor_strip_down(sb_compositor *c, int level)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(level >= 0 && level < c->num_bands);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return mb_down_stage(c, c->slots[0], level, c->g_lo[level + 1], c->g_hi[level + 1]);
}

int sb_compositor_strip_band(sb_compositor *c, int level)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(level >= 0 && level <= c->num_bands);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return mb_band_stage(c, c->slots[0], level, c->b_lo[level], c->b_hi[level]);
}

// The exchange steps of a frame, in schedule order (strips.py: schedule): Gaussian level l before the pyrDown that reads it
// (l = 0 .. num_bands), restored band l right after its band (l = num_bands .. 1).
static void peer_steps(const sb_compositor *c, std::vector<std::pair<int, int>> *steps)
{
    steps->clear();
    for (int l = 0; l <= c->num_bands; ++l) steps->emplace_back(SB_HALO_GAUSS, l);
    for (int l = c->num_bands; l >= 1; --l) steps->emplace_back(SB_HALO_RESTORED, l);
}

// Receive area of `side`: 128 bytes of flags, then one region per exchange step, each the size of what arrives from that side.
int sb_compositor_strip_peer_export(sb_compositor *c, int side, void *ipc_handle_64, void **local_ptr, size_t *bytes)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(side == SB_SIDE_LEFT || side == SB_SIDE_RIGHT);
    if (c->strip_recompute) return fail(SB_ERR_ASSERT, "the peer exchange belongs to the exchange halo mode");
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<std::pair<int, int>> steps;
    peer_steps(c, &steps);
    size_t total = 0;
    c->peer_region_mine[side].clear();
    for (auto &st : steps) {
        size_t sb = 0, rb = 0;
        SB_TRY(sb_compositor_strip_halo_bytes(c, st.first, st.second, side, &sb, &rb));
        c->peer_region_mine[side].push_back(total);
        total += (rb + 127) & ~(size_t)127;
    }
    // what this rank SENDS to that side lands in the neighbour's area of the opposite side, laid out by the same rule
    size_t ttotal = 0;
    c->peer_region_theirs[side].clear();
    for (auto &st : steps) {
        size_t sb = 0, rb = 0;
        SB_TRY(sb_compositor_strip_halo_bytes(c, st.first, st.second, side, &sb, &rb));
        c->peer_region_theirs[side].push_back(ttotal);
        ttotal += (sb + 127) & ~(size_t)127;
    }
    SB_TRY(c->peer_mine[side].ensure(128 + total + 128));
    SB_CUDA(cudaMemset(c->peer_mine[side].p, 0, 128 + total + 128));
    if (!c->peer_done.p) { SB_TRY(c->peer_done.ensure(64)); SB_CUDA(cudaMemset(c->peer_done.p, 0, 64)); }
    c->peer_frames = 0;
    if (ipc_handle_64) {
        cudaIpcMemHandle_t h;
        SB_CUDA(cudaIpcGetMemHandle(&h, c->peer_mine[side].p));
        static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(ipc_handle_64, &h, 64);
    }
    if (local_ptr) *local_ptr = c->peer_mine[side].p;
    if (bytes) *bytes = 128 + total + 128;
    return SB_OK;
}

// Map the neighbour's receive area: its IPC handle (another process) or its plain device pointer (same process).
int sb_compositor_strip_peer_connect(sb_compositor *c, int side, const void *ipc_handle_64, void *same_process_ptr)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(side == SB_SIDE_LEFT || side == SB_SIDE_RIGHT);
    SB_ASSERT((ipc_handle_64 != nullptr) != (same_process_ptr != nullptr));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    if (c->peer_theirs[side] && c->peer_ipc[side]) cudaIpcCloseMemHandle(c->peer_theirs[side]);
    c->peer_theirs[side] = nullptr; c->peer_ipc[side] = false;
    if (ipc_handle_64) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, ipc_handle_64, 64);
        void *p = nullptr;
        SB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_theirs[side] = static_cast<char *>(p); c->peer_ipc[side] = true;
    } else {
        c->peer_theirs[side] = static_cast<char *>(same_process_ptr);
    }
    return SB_OK;
}

// which = 1: push this rank's edge columns of (what, level) into both neighbours' receive areas and raise their flags;
// which = 2: wait for this rank's flags and move what arrived into place;  3: both (push first)
static int peer_exchange(sb_compositor *c, int what, int level, int which)
{
    cudaStream_t st = c->slots[0].stream;
    const int nb = c->num_bands;
    const int step = what == SB_HALO_GAUSS ? level : nb + 1 + (nb - level);
    if (step == 0 && (which & 1)) ++c->peer_frames;         // a frame's first push opens its block of sequence numbers
    SB_ASSERT(c->peer_frames > 0);
    const unsigned long long seq = (c->peer_frames - 1) * (2ull * nb + 1) + step + 1;
    HaloXferArgs push{}, pull{};
    push.seq = pull.seq = seq;
    push.done = static_cast<unsigned *>(c->peer_done.p);
    for (int side = 0; side < 2; ++side) {
        std::vector<HaloSeg> segs;
        SB_TRY(halo_segments(c, what, level, side, &segs));
        if (segs.empty()) continue;
        if (!c->peer_theirs[side] || !c->peer_mine[side].p) return fail(SB_ERR_ASSERT, "peer exchange: side %d is not connected (peer_export / peer_connect)", side);
        push.area[side] = c->peer_theirs[side]; push.region[side] = c->peer_region_theirs[side][step];
        pull.area[side] = static_cast<char *>(c->peer_mine[side].p); pull.region[side] = c->peer_region_mine[side][step];
        unsigned long long so = 0, ro = 0;
        for (const HaloSeg &h : segs) {
            if (h.ncols_send) { SB_ASSERT(push.n < SB_HALO_MAX_COPIES); push.c[push.n++] = HaloCopy{h.base, h.step, h.esz, h.rows, h.send_col, h.ncols_send, so, side}; }
            if (h.ncols_recv) { SB_ASSERT(pull.n < SB_HALO_MAX_COPIES); pull.c[pull.n++] = HaloCopy{h.base, h.step, h.esz, h.rows, h.recv_col, h.ncols_recv, ro, side}; }
            so += (unsigned long long)h.rows * h.ncols_send * h.esz;
            ro += (unsigned long long)h.rows * h.ncols_recv * h.esz;
        }
    }
    if (!push.area[0] && !push.area[1]) return SB_OK;       // a single strip: nothing to exchange
    if (which & 1) { k_halo_push<<<8, 128, 0, st>>>(push); SB_LAUNCHED(); }
    if (which & 2) { k_halo_pull<<<8, 128, 0, st>>>(pull); SB_LAUNCHED(); }
    return SB_OK;
}

// The two halves of one exchange step on their own (a single process that plays several ranks on ONE device must enqueue every
// rank's push of a step before any rank's pull: streams of one process can share a hardware queue, and a pull that waits for a
// push queued behind it would wait for ever).  One rank per process / GPU: sb_compositor_strip_frame_peer.
int sb_compositor_strip_peer_push(sb_compositor *c, int what, int level)
{
    SB_TRY(strip_ready(c));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return peer_exchange(c, what, level, 1);
}
int sb_compositor_strip_peer_pull(sb_compositor *c, int what, int level)
{
    SB_TRY(strip_ready(c));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    return peer_exchange(c, what, level, 2);
}

// One frame of this rank's strip, every stage and every halo exchange enqueued by this ONE call (exchange halo mode).
int sb_compositor_strip_frame_peer(sb_compositor *c, const sb_image *srcs)
{
    SB_TRY(strip_ready(c));
    if (c->strip_recompute) return fail(SB_ERR_ASSERT, "strip_frame_peer runs the exchange halo mode");
    SB_TRY(sb_compositor_strip_warp(c, srcs));
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    const int nb = c->num_bands;
    for (int l = 0; l <= nb; ++l) {
        SB_TRY(peer_exchange(c, SB_HALO_GAUSS, l, 3));
        if (l < nb) SB_TRY(mb_down_stage(c, s, l, c->g_lo[l + 1], c->g_hi[l + 1]));
    }
    for (int l = nb; l >= 0; --l) {
        SB_TRY(mb_band_stage(c, s, l, c->b_lo[l], c->b_hi[l]));
        if (l >= 1) SB_TRY(peer_exchange(c, SB_HALO_RESTORED, l, 3));
    }
    return SB_OK;
}

int sb_compositor_strip_result(sb_compositor *c, sb_image *strip, sb_image *strip_mask)
{
    SB_TRY(strip_ready(c));
    SB_ASSERT(strip);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    Slot &s = c->slots[0];
    const int x0 = std::min(c->strip_x0, s.out.v.cols), x1 = std::min(c->strip_x1, s.out.v.cols);
    DImage v = s.out.v, m = s.out_mask.v;
    v.data = static_cast<char *>(v.data) + (size_t)x0 * elem_size(v.type); v.cols = x1 - x0;
    m.data = static_cast<char *>(m.data) + x0; m.cols = x1 - x0;
    if (!strip->data) lend(v, c->device, strip);
    else SB_TRY(from_device(v, strip, s.stream));
    if (strip_mask) {
        if (!strip_mask->data) lend(m, c->device, strip_mask);
        else SB_TRY(from_device(m, strip_mask, s.stream));
    }
    SB_CUDA(cudaStreamSynchronize(s.stream));
    return SB_OK;
}

int sb_compositor_num_bands(const sb_compositor *c)
{
    return c ? c->num_bands : -1;
}

}  // extern "C"
