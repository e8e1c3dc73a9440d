// capi_calib.cu — calibration-table serialization.  The expensive part of a (re)calibration is
// host work outside this path (features, matching, bundle adjustment, seams, exposure: 1-2 s per APP64:696-722); what it
// hands to the per-frame path is small: K, R, warper kind/scale, blender settings, gains or block gain maps, seam masks
// (sb_compositor_config).  Saving exactly that lets a video process resume — or another GPU/rank start — without
// recalibrating: the device tables (maps, tap tables, weight pyramids, tile descriptors) are rebuilt from it in
// milliseconds by sb_compositor_create and are bit-identical by construction.  Host-only code, no CUDA calls.
#include <cstdio>
#include <memory>

#include "sb_internal.h"

using namespace sb;

struct sb_calibration {
    sb_compositor_config cfg{};
    std::vector<float> K, R;
    std::vector<double> gains;
    std::vector<sb_image> seam_masks, gain_maps, umap1, umap2;
    std::vector<std::vector<uint8_t>> blobs;        // pixel storage of the images above
};

namespace {

const char kMagic[8] = {'S', 'B', 'C', 'A', 'L', '0', '0', '2'};       // 002: + projector a/b, undistort maps, crop margins
const char kMagic1[8] = {'S', 'B', 'C', 'A', 'L', '0', '0', '1'};      // (round-1 files still load)

struct Writer {
    std::vector<uint8_t> buf;
    void put(const void *p, size_t n) { const uint8_t *b = static_cast<const uint8_t *>(p); buf.insert(buf.end(), b, b + n); }
    template <typename T> void pod(const T &v) { put(&v, sizeof v); }
    void image(const sb_image &im)
    {
        const int32_t hdr[3] = {im.rows, im.cols, im.type};
        put(hdr, sizeof hdr);
        const size_t row = (size_t)im.cols * elem_size(im.type);
        for (int y = 0; y < im.rows; ++y) put(static_cast<const uint8_t *>(im.data) + (size_t)y * im.step, row);
    }
};
struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    bool get(void *dst, size_t n) { if (!ok || (size_t)(end - p) < n) { ok = false; return false; } std::memcpy(dst, p, n); p += n; return true; }
    template <typename T> T pod() { T v{}; get(&v, sizeof v); return v; }
};
uint64_t fnv1a(const uint8_t *p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
bool read_image(Reader &r, int want_type, sb_calibration &c, sb_image *out)
{
    int32_t hdr[3];
    if (!r.get(hdr, sizeof hdr) || hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] != want_type) return r.ok = false;
    const size_t row = (size_t)hdr[1] * elem_size(hdr[2]), bytes = row * hdr[0];
    if ((size_t)(r.end - r.p) < bytes) return r.ok = false;
    c.blobs.emplace_back(bytes);
    r.get(c.blobs.back().data(), bytes);
    *out = sb_image{c.blobs.back().data(), hdr[0], hdr[1], hdr[2], row, -1};
    return true;
}

}  // namespace

extern "C" {

int sb_calibration_save(const sb_compositor_config *cfg, const char *path)
{
    SB_ASSERT(cfg && path && cfg->n_cameras > 0 && cfg->n_cameras <= SB_MAX_COMPOSITOR_CAMERAS && cfg->K && cfg->R);
    const int n = cfg->n_cameras;
    Writer w;
    w.put(kMagic, 8);
    const int32_t head[10] = {n, cfg->src_size.width, cfg->src_size.height, cfg->warper_kind, cfg->blender_kind, cfg->num_bands,
                              cfg->weight_type, cfg->comp_kind, cfg->output_type, 0};
    w.put(head, sizeof head);
    w.pod(cfg->warper_scale); w.pod(cfg->sharpness);
    w.put(cfg->K, sizeof(float) * 9 * n); w.put(cfg->R, sizeof(float) * 9 * n);
    const int32_t has[3] = {cfg->gains ? 1 : 0, cfg->seam_masks ? 1 : 0, cfg->gain_maps ? 1 : 0};
    w.put(has, sizeof has);
    if (cfg->gains) w.put(cfg->gains, sizeof(double) * n);
    for (int i = 0; cfg->seam_masks && i < n; ++i) {
        SB_ASSERT(cfg->seam_masks[i].data && cfg->seam_masks[i].device < 0 && cfg->seam_masks[i].type == SB_8UC1);
        w.image(cfg->seam_masks[i]);
    }
    for (int i = 0; cfg->gain_maps && i < n; ++i) {
        SB_ASSERT(cfg->gain_maps[i].data && cfg->gain_maps[i].device < 0 && cfg->gain_maps[i].type == SB_32FC1);
        w.image(cfg->gain_maps[i]);
    }
    // format 002
    w.pod(cfg->warper_a); w.pod(cfg->warper_b);
    w.pod(cfg->crop_up); w.pod(cfg->crop_down);
    const int32_t crop[3] = {cfg->crop_left, cfg->crop_right, cfg->crop_app_fill};
    w.put(crop, sizeof crop);
    SB_ASSERT((cfg->undistort_map1 != nullptr) == (cfg->undistort_map2 != nullptr));
    const int32_t has_u = cfg->undistort_map1 ? 1 : 0;
    w.pod(has_u);
    for (int i = 0; has_u && i < n; ++i) {
        SB_ASSERT(cfg->undistort_map1[i].data && cfg->undistort_map1[i].device < 0 && cfg->undistort_map1[i].type == SB_16SC2);
        SB_ASSERT(cfg->undistort_map2[i].data && cfg->undistort_map2[i].device < 0 && cfg->undistort_map2[i].type == SB_16UC1);
        w.image(cfg->undistort_map1[i]);
        w.image(cfg->undistort_map2[i]);
    }
    w.pod(fnv1a(w.buf.data(), w.buf.size()));
    const std::string tmp = std::string(path) + ".tmp";      // write-then-rename: a crash never leaves a torn file behind
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) return fail(SB_ERR_BAD_ARG, "cannot open %s for writing", tmp.c_str());
    const bool ok = std::fwrite(w.buf.data(), 1, w.buf.size(), f) == w.buf.size();
    if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return fail(SB_ERR_BAD_ARG, "writing %s failed", path); }
    return SB_OK;
}

int sb_calibration_load(const char *path, sb_calibration **out)
{
    SB_ASSERT(path && out);
    *out = nullptr;
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(SB_ERR_BAD_ARG, "cannot open %s", path);
    std::vector<uint8_t> buf;
    uint8_t chunk[65536];
    for (size_t k; (k = std::fread(chunk, 1, sizeof chunk, f)) > 0;) buf.insert(buf.end(), chunk, chunk + k);
    std::fclose(f);
    const bool v2 = buf.size() >= 8 && std::memcmp(buf.data(), kMagic, 8) == 0, v1 = buf.size() >= 8 && std::memcmp(buf.data(), kMagic1, 8) == 0;
    if (buf.size() < 8 + 40 + 8 + 8 || !(v1 || v2)) return fail(SB_ERR_BAD_ARG, "%s is not a calibration file", path);
    uint64_t sum;
    std::memcpy(&sum, buf.data() + buf.size() - 8, 8);
    if (sum != fnv1a(buf.data(), buf.size() - 8)) return fail(SB_ERR_BAD_ARG, "%s: checksum mismatch (truncated or corrupt)", path);
    std::unique_ptr<sb_calibration> c(new sb_calibration);
    Reader r{buf.data() + 8, buf.data() + buf.size() - 8};
    int32_t head[10];
    r.get(head, sizeof head);
    const int n = head[0];
    if (!r.ok || n <= 0 || n > SB_MAX_COMPOSITOR_CAMERAS) return fail(SB_ERR_BAD_ARG, "%s: bad header", path);
    sb_compositor_config &g = c->cfg;
    g.n_cameras = n; g.src_size = sb_size{head[1], head[2]}; g.warper_kind = head[3]; g.blender_kind = head[4]; g.num_bands = head[5];
    g.weight_type = head[6]; g.comp_kind = head[7]; g.output_type = head[8];
    g.warper_scale = r.pod<float>(); g.sharpness = r.pod<float>();
    c->K.resize(9 * (size_t)n); c->R.resize(9 * (size_t)n);
    r.get(c->K.data(), sizeof(float) * 9 * n); r.get(c->R.data(), sizeof(float) * 9 * n);
    int32_t has[3] = {0, 0, 0};
    r.get(has, sizeof has);
    c->blobs.reserve(4 * (size_t)n);                          // (sb_image::data points into the blobs: no reallocation later)
    if (has[0]) { c->gains.resize(n); r.get(c->gains.data(), sizeof(double) * n); }
    if (has[1]) { c->seam_masks.resize(n); for (int i = 0; i < n && r.ok; ++i) read_image(r, SB_8UC1, *c, &c->seam_masks[i]); }
    if (has[2]) { c->gain_maps.resize(n); for (int i = 0; i < n && r.ok; ++i) read_image(r, SB_32FC1, *c, &c->gain_maps[i]); }
    int32_t has_u = 0;
    if (v2) {
        g.warper_a = r.pod<float>(); g.warper_b = r.pod<float>();
        g.crop_up = r.pod<float>(); g.crop_down = r.pod<float>();
        int32_t crop[3] = {0, 0, 0};
        r.get(crop, sizeof crop);
        g.crop_left = crop[0]; g.crop_right = crop[1]; g.crop_app_fill = crop[2];
        has_u = r.pod<int32_t>();
        if (has_u) {
            c->umap1.resize(n); c->umap2.resize(n);
            for (int i = 0; i < n && r.ok; ++i) { read_image(r, SB_16SC2, *c, &c->umap1[i]); read_image(r, SB_16UC1, *c, &c->umap2[i]); }
        }
    }
    if (!r.ok || r.p != r.end) return fail(SB_ERR_BAD_ARG, "%s: malformed body", path);
    g.undistort_map1 = has_u ? c->umap1.data() : nullptr;
    g.undistort_map2 = has_u ? c->umap2.data() : nullptr;
    g.K = c->K.data(); g.R = c->R.data();
    g.gains = has[0] ? c->gains.data() : nullptr;
    g.seam_masks = has[1] ? c->seam_masks.data() : nullptr;
    g.gain_maps = has[2] ? c->gain_maps.data() : nullptr;
    *out = c.release();
    return SB_OK;
}

const sb_compositor_config *sb_calibration_config(const sb_calibration *c) { return c ? &c->cfg : nullptr; }

void sb_calibration_free(sb_calibration *c) { delete c; }

}  // extern "C"
