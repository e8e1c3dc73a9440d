// capi_blender.cu — Blender / FeatherBlender / MultiBandBlender C ABI (blenders.hpp:53-117) and the
// blenders.hpp:122-133 auxiliary functions, over the pyramid and blend kernels.
//
// Host logic restated from blenders.cpp: prepare (:65-78,115-120,203-233), the padded-rect
// geometry of MultiBandBlender::feed (:241-269), level bookkeeping (:300-356), blend (:359-377).
#include <algorithm>
#include <cmath>

#include "sb_kernels.h"

using namespace sb;

struct sb_blender {
    int device = 0;
    int kind = 0;
    int actual_num_bands = 5, num_bands = 5;
    int weight_type = SB_32F;
    float sharpness = 0.02f;
    cudaStream_t stream = nullptr;
    bool prepared = false;
    sb_rect dst_roi = {0, 0, 0, 0}, dst_roi_final = {0, 0, 0, 0};
    // accumulators
    DevImage dst_mask, dst_weight;                       // NO / FEATHER
    std::vector<DevImage> pyr_laplace, band_weights;     // [0] is dst_ for every kind
    // per-feed scratch (grow-only, reused across feeds and frames)
    DevImage img_stage, mask_stage, weight_map;
    std::vector<DevImage> src_pyr, w_pyr;
    DevBuf dist_scratch;
    DevImage out, out_mask;
};

namespace {

int ceil_log2_len(int w, int h)
{
    double max_len = static_cast<double>(std::max(w, h));
    return static_cast<int>(std::ceil(std::log(max_len) / std::log(2.0)));   // blenders.cpp:208-209
}

int prepare_rect(sb_blender *b, sb_rect roi)
{
    SB_ASSERT(roi.width > 0 && roi.height > 0);
    b->dst_roi_final = roi;
    int levels = 0;
    if (b->kind == SB_BLEND_MULTI_BAND) {
        b->num_bands = std::min(b->actual_num_bands, ceil_log2_len(roi.width, roi.height));
        SB_ASSERT(b->num_bands >= 0 && b->num_bands < 24);
        const int m = 1 << b->num_bands;
        roi.width += (m - roi.width % m) % m;
        roi.height += (m - roi.height % m) % m;
        levels = b->num_bands;
    }
    b->dst_roi = roi;
    b->pyr_laplace.resize(levels + 1);
    SB_TRY(b->pyr_laplace[0].create_zero(roi.height, roi.width, SB_16SC3, b->stream));   // Blender::prepare(Rect)
    if (b->kind == SB_BLEND_NO) SB_TRY(b->dst_mask.create_zero(roi.height, roi.width, SB_8UC1, b->stream));
    if (b->kind == SB_BLEND_FEATHER) SB_TRY(b->dst_weight.create_zero(roi.height, roi.width, SB_32FC1, b->stream));
    if (b->kind == SB_BLEND_MULTI_BAND) {
        const int wt = b->weight_type == SB_32F ? SB_32FC1 : SB_16SC1;
        b->band_weights.resize(levels + 1);
        SB_TRY(b->band_weights[0].create_zero(roi.height, roi.width, wt, b->stream));
        for (int i = 1; i <= levels; ++i) {
            const DImage &prev = b->pyr_laplace[i - 1].v;
            SB_TRY(b->pyr_laplace[i].create_zero((prev.rows + 1) / 2, (prev.cols + 1) / 2, SB_16SC3, b->stream));
            SB_TRY(b->band_weights[i].create_zero((prev.rows + 1) / 2, (prev.cols + 1) / 2, wt, b->stream));
        }
    }
    b->prepared = true;
    return SB_OK;
}

int feed_multiband(sb_blender *b, const DImage &img, const DImage &mask, sb_point tl)
{
    const int nb = b->num_bands;
    const sb_rect &r = b->dst_roi;
    const FeedRect fr = multiband_feed_rect(r, tl, img.cols, img.rows, nb);       // blenders.cpp:241-269
    const sb_point tl_new = fr.tl_new;
    const int width = fr.width, height = fr.height, top = fr.top, left = fr.left, bottom = fr.bottom, right = fr.right;
    if (top < 0 || left < 0 || bottom < 0 || right < 0 || tl_new.x < r.x || tl_new.y < r.y)
        return fail(SB_ERR_ASSERT, "feed: image at (%d,%d) %dx%d does not fit the prepared ROI", tl.x, tl.y, img.cols, img.rows);

    // Gaussian pyramid of the bordered image; the Laplacian is formed on the fly while accumulating
    b->src_pyr.resize(std::max<size_t>(b->src_pyr.size(), nb + 1));
    b->w_pyr.resize(std::max<size_t>(b->w_pyr.size(), nb + 1));
    SB_TRY(b->src_pyr[0].create(height, width, img.type));
    SB_TRY(launch_copy_make_border(img, b->src_pyr[0].v, top, left, SB_BORDER_REFLECT, b->stream));
    for (int i = 0; i < nb; ++i) {
        const DImage &p = b->src_pyr[i].v;
        SB_TRY(b->src_pyr[i + 1].create((p.rows + 1) / 2, (p.cols + 1) / 2, img.type));
        SB_TRY(launch_pyr_down(p, b->src_pyr[i + 1].v, b->stream));
    }
    // weight map Gaussian pyramid (:282-298)
    const int wt = b->weight_type == SB_32F ? SB_32FC1 : SB_16SC1;
    SB_TRY(b->w_pyr[0].create(height, width, wt));
    SB_TRY(launch_mask_to_weight(mask, b->w_pyr[0].v, top, left, b->stream));
    for (int i = 0; i < nb; ++i) {
        const DImage &p = b->w_pyr[i].v;
        SB_TRY(b->w_pyr[i + 1].create((p.rows + 1) / 2, (p.cols + 1) / 2, wt));
        SB_TRY(launch_pyr_down(p, b->w_pyr[i + 1].v, b->stream));
    }
    // :300-356
    int x_tl = tl_new.x - r.x, y_tl = tl_new.y - r.y;
    for (int i = 0; i <= nb; ++i) {
        DImage coarse;
        if (i < nb) coarse = b->src_pyr[i + 1].v;
        SB_TRY(launch_lap_accumulate(b->src_pyr[i].v, coarse, b->w_pyr[i].v, b->pyr_laplace[i].v, b->band_weights[i].v, x_tl, y_tl, b->stream));
        x_tl /= 2; y_tl /= 2;
    }
    return SB_OK;
}

int deliver(sb_blender *b, const DImage &src, sb_image *dst)
{
    if (!dst) return SB_OK;
    if (!dst->data) { lend(src, b->device, dst); return SB_OK; }
    return from_device(src, dst, b->stream);
}

}  // namespace

extern "C" {

int sb_blender_create(int kind, int num_bands, int weight_type, float sharpness, int device, sb_blender **out)
{
    if (!out) return fail(SB_ERR_ASSERT, "out is null");
    *out = nullptr;
    if (kind != SB_BLEND_NO && kind != SB_BLEND_FEATHER && kind != SB_BLEND_MULTI_BAND)
        return fail(SB_ERR_BAD_ARG, "unsupported blending method");                          // blenders.cpp:60
    if (kind == SB_BLEND_MULTI_BAND && weight_type != SB_32F && weight_type != SB_16S)
        return fail(SB_ERR_ASSERT, "weight_type == CV_32F || weight_type == CV_16S");        // blenders.cpp:198
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_blender *b = new sb_blender;
    b->device = device; b->kind = kind; b->actual_num_bands = num_bands; b->num_bands = num_bands;
    b->weight_type = weight_type; b->sharpness = sharpness;
    cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete b; return fail(SB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = b;
    return SB_OK;
}

void sb_blender_destroy(sb_blender *b)
{
    if (!b) return;
    DeviceGuard g(b->device);
    if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
    delete b;
}

int sb_blender_num_bands(const sb_blender *b) { return b ? b->actual_num_bands : 0; }
int sb_blender_set_num_bands(sb_blender *b, int n) { SB_ASSERT(b); b->actual_num_bands = n; return SB_OK; }
float sb_blender_sharpness(const sb_blender *b) { return b ? b->sharpness : 0.f; }
int sb_blender_set_sharpness(sb_blender *b, float s) { SB_ASSERT(b); b->sharpness = s; return SB_OK; }

int sb_blender_prepare_rect(sb_blender *b, sb_rect dst_roi)
{
    SB_ASSERT(b);
    DeviceGuard g(b->device);
    if (!g.ok) return SB_ERR_CUDA;
    return prepare_rect(b, dst_roi);
}

int sb_blender_prepare(sb_blender *b, const sb_point *corners, const sb_size *sizes, int n)
{
    SB_ASSERT(b && corners && sizes && n > 0);
    // resultRoi (util.cpp:127-140)
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners[i].x); tly = std::min(tly, corners[i].y);
        brx = std::max(brx, corners[i].x + sizes[i].width); bry = std::max(bry, corners[i].y + sizes[i].height);
    }
    sb_rect roi = {tlx, tly, brx - tlx, bry - tly};
    return sb_blender_prepare_rect(b, roi);
}

int sb_blender_result_size(const sb_blender *b, sb_size *size)
{
    SB_ASSERT(b && size);
    if (!b->prepared) return fail(SB_ERR_ASSERT, "blender is not prepared");
    size->width = b->dst_roi_final.width;
    size->height = b->dst_roi_final.height;
    return SB_OK;
}

int sb_blender_feed(sb_blender *b, const sb_image *img, const sb_image *mask, sb_point tl)
{
    SB_ASSERT(b && img && mask);
    if (!b->prepared) return fail(SB_ERR_ASSERT, "feed before prepare");
    if (b->kind == SB_BLEND_MULTI_BAND) SB_ASSERT(img->type == SB_16SC3 || img->type == SB_8UC3);   // blenders.cpp:238
    else SB_ASSERT(img->type == SB_16SC3);                                                            // :83, :125
    SB_ASSERT(mask->type == SB_8UC1);                                                                 // :84, :126, :239
    SB_ASSERT(mask->rows == img->rows && mask->cols == img->cols);
    DeviceGuard g(b->device);
    if (!g.ok) return SB_ERR_CUDA;
    DImage dimg, dmask;
    SB_TRY(to_device(*img, b->img_stage, b->stream, &dimg));
    SB_TRY(to_device(*mask, b->mask_stage, b->stream, &dmask));
    const int dx = tl.x - b->dst_roi.x, dy = tl.y - b->dst_roi.y;
    int rc;
    if (b->kind == SB_BLEND_MULTI_BAND) rc = feed_multiband(b, dimg, dmask, tl);
    else if (b->kind == SB_BLEND_FEATHER) {
        rc = b->weight_map.create(dimg.rows, dimg.cols, SB_32FC1);                                    // createWeightMap
        if (rc == SB_OK) rc = launch_distance_l1(dmask, b->weight_map.v, b->dist_scratch, b->stream);
        if (rc == SB_OK) rc = launch_weight_from_dist(b->weight_map.v, b->sharpness, b->stream);
        if (rc == SB_OK) rc = launch_feather_accumulate(dimg, b->weight_map.v, b->pyr_laplace[0].v, b->dst_weight.v, dx, dy, b->stream);
    } else
        rc = launch_masked_copy(dimg, dmask, b->pyr_laplace[0].v, b->dst_mask.v, dx, dy, b->stream);
    SB_TRY(rc);
    // host inputs were staged from pageable memory: make the staging buffers reusable
    if (img->device < 0 || mask->device < 0) SB_CUDA(cudaStreamSynchronize(b->stream));
    return SB_OK;
}

int sb_blender_blend(sb_blender *b, sb_image *dst, sb_image *dst_mask)
{
    SB_ASSERT(b);
    if (!b->prepared) return fail(SB_ERR_ASSERT, "blend before prepare");
    DeviceGuard g(b->device);
    if (!g.ok) return SB_ERR_CUDA;
    const int w = b->dst_roi_final.width, h = b->dst_roi_final.height;
    SB_TRY(b->out.create(h, w, SB_16SC3));
    SB_TRY(b->out_mask.create(h, w, SB_8UC1));
    DImage none;
    if (b->kind == SB_BLEND_MULTI_BAND) {
        const int nb = b->num_bands;
        SB_TRY(launch_normalize(b->band_weights[nb].v, b->pyr_laplace[nb].v, b->stream));
        for (int i = nb - 1; i >= 0; --i)
            SB_TRY(launch_normalize_collapse(b->pyr_laplace[i + 1].v, b->band_weights[i].v, b->pyr_laplace[i].v, b->stream));
        SB_TRY(launch_finalize(b->pyr_laplace[0].v, b->band_weights[0].v, nullptr, b->out.v, b->out_mask.v, b->stream));
    } else if (b->kind == SB_BLEND_FEATHER) {
        SB_TRY(launch_normalize(b->dst_weight.v, b->pyr_laplace[0].v, b->stream));
        SB_TRY(launch_finalize(b->pyr_laplace[0].v, b->dst_weight.v, nullptr, b->out.v, b->out_mask.v, b->stream));
    } else
        SB_TRY(launch_finalize(b->pyr_laplace[0].v, none, &b->dst_mask.v, b->out.v, b->out_mask.v, b->stream));
    SB_TRY(deliver(b, b->out.v, dst));
    SB_TRY(deliver(b, b->out_mask.v, dst_mask));
    SB_CUDA(cudaStreamSynchronize(b->stream));
    b->prepared = false;   // the reference releases its buffers here (blenders.cpp:108-111, 373-374)
    return SB_OK;
}

// ---- blenders.hpp:122-133 auxiliaries (one-shot helpers on the default stream) ----
int sb_normalize_using_weight_map(const sb_image *weight, sb_image *src, int device)
{
    SB_ASSERT(weight && src);
    SB_ASSERT(src->type == SB_16SC3);                                   // blenders.cpp:389
    SB_ASSERT(weight->type == SB_32FC1 || weight->type == SB_16SC1);    // :408
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    DevImage sw, ss;
    DImage dw, ds;
    SB_TRY(to_device(*weight, sw, nullptr, &dw));
    SB_TRY(to_device(*src, ss, nullptr, &ds));
    SB_TRY(launch_normalize(dw, ds, nullptr));
    if (src->device < 0) SB_TRY(from_device(ds, src, nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

int sb_create_weight_map(const sb_image *mask, float sharpness, sb_image *weight, int device)
{
    SB_ASSERT(mask && weight && weight->data);
    SB_ASSERT(mask->type == SB_8UC1);                                   // blenders.cpp:429
    SB_ASSERT(weight->type == SB_32FC1 && weight->rows == mask->rows && weight->cols == mask->cols);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    DevImage sm, swt;
    DevBuf scratch;
    DImage dm, dwt;
    SB_TRY(to_device(*mask, sm, nullptr, &dm));
    if (weight->device >= 0) { dwt.data = weight->data; dwt.rows = weight->rows; dwt.cols = weight->cols; dwt.type = weight->type; dwt.step = weight->step; }
    else { SB_TRY(swt.create(mask->rows, mask->cols, SB_32FC1)); dwt = swt.v; }
    SB_TRY(launch_distance_l1(dm, dwt, scratch, nullptr));
    SB_TRY(launch_weight_from_dist(dwt, sharpness, nullptr));
    if (weight->device < 0) SB_TRY(from_device(dwt, weight, nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

// FeatherBlender::createWeightMaps (blenders.hpp:80-81, blenders.cpp:158-186)
int sb_blender_create_weight_maps(sb_blender *b, const sb_image *masks, const sb_point *corners, int n, sb_image *weight_maps, sb_rect *dst_roi)
{
    SB_ASSERT(b && masks && corners && weight_maps && dst_roi && n > 0);
    SB_ASSERT(b->kind == SB_BLEND_FEATHER);
    DeviceGuard g(b->device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<sb_size> sizes(n);
    for (int i = 0; i < n; ++i) {
        SB_TRY(check_image(&masks[i], "mask"));
        SB_ASSERT(masks[i].type == SB_8UC1);                            // blenders.cpp:429
        SB_ASSERT(weight_maps[i].data && weight_maps[i].type == SB_32FC1 && weight_maps[i].rows == masks[i].rows && weight_maps[i].cols == masks[i].cols);
        sizes[i] = sb_size{masks[i].cols, masks[i].rows};
    }
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;       // resultRoi(corners, masks), util.cpp:118-140
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners[i].x); tly = std::min(tly, corners[i].y);
        brx = std::max(brx, corners[i].x + sizes[i].width); bry = std::max(bry, corners[i].y + sizes[i].height);
    }
    const sb_rect roi = {tlx, tly, brx - tlx, bry - tly};
    std::vector<DevImage> stage(n), wown(n);
    std::vector<DImage> w(n);
    DevImage sum;
    SB_TRY(sum.create_zero(roi.height, roi.width, SB_32FC1, b->stream));
    for (int i = 0; i < n; ++i) {
        DImage dm;
        SB_TRY(to_device(masks[i], stage[i], b->stream, &dm));
        if (weight_maps[i].device >= 0) { w[i].data = weight_maps[i].data; w[i].rows = dm.rows; w[i].cols = dm.cols; w[i].type = SB_32FC1; w[i].step = weight_maps[i].step; }
        else { SB_TRY(wown[i].create(dm.rows, dm.cols, SB_32FC1)); w[i] = wown[i].v; }
        SB_TRY(launch_distance_l1(dm, w[i], b->dist_scratch, b->stream));              // createWeightMap (blenders.cpp:427-432)
        SB_TRY(launch_weight_from_dist(w[i], b->sharpness, b->stream));
        SB_TRY(launch_weight_accumulate(w[i], sum.v, corners[i].x - roi.x, corners[i].y - roi.y, b->stream));   // weights_sum(roi) += weight_maps[i]
    }
    for (int i = 0; i < n; ++i) {
        SB_TRY(launch_weight_normalize(w[i], sum.v, corners[i].x - roi.x, corners[i].y - roi.y, b->stream));
        if (weight_maps[i].device < 0) SB_TRY(from_device(w[i], &weight_maps[i], b->stream));
    }
    SB_CUDA(cudaStreamSynchronize(b->stream));
    *dst_roi = roi;
    return SB_OK;
}

int sb_create_laplace_pyr(const sb_image *img, int num_levels, sb_image *pyr, int device)
{
    SB_ASSERT(img && pyr && num_levels >= 0);
    SB_ASSERT(img->type == SB_16SC3 || img->type == SB_8UC3);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<DevImage> gauss(num_levels + 1), lap(num_levels + 1);
    DImage g0;
    SB_TRY(to_device(*img, gauss[0], nullptr, &g0));
    std::vector<DImage> gv(num_levels + 1);
    gv[0] = g0;
    for (int i = 0; i < num_levels; ++i) {
        SB_TRY(gauss[i + 1].create((gv[i].rows + 1) / 2, (gv[i].cols + 1) / 2, img->type));
        gv[i + 1] = gauss[i + 1].v;
        SB_TRY(launch_pyr_down(gv[i], gv[i + 1], nullptr));
    }
    for (int i = 0; i <= num_levels; ++i) {
        SB_ASSERT(pyr[i].data && pyr[i].type == SB_16SC3 && pyr[i].rows == gv[i].rows && pyr[i].cols == gv[i].cols);
        SB_TRY(lap[i].create(gv[i].rows, gv[i].cols, SB_16SC3));
        if (i < num_levels) {
            SB_ASSERT(gv[i + 1].cols * 2 == gv[i].cols && gv[i + 1].rows * 2 == gv[i].rows);
            SB_TRY(launch_laplace_level(gv[i], gv[i + 1], lap[i].v, nullptr));
        } else
            SB_TRY(launch_convert(gv[i], lap[i].v, nullptr));
        SB_TRY(from_device(lap[i].v, &pyr[i], nullptr));
    }
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

int sb_restore_image_from_laplace_pyr(sb_image *pyr, int num_images, int device)
{
    SB_ASSERT(pyr && num_images >= 0);
    if (num_images == 0) return SB_OK;                                  // blenders.cpp:522-523
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<DevImage> st(num_images);
    std::vector<DImage> v(num_images);
    for (int i = 0; i < num_images; ++i) {
        SB_ASSERT(pyr[i].type == SB_16SC3);
        SB_TRY(to_device(pyr[i], st[i], nullptr, &v[i]));
    }
    for (int i = num_images - 1; i > 0; --i) SB_TRY(launch_collapse_level(v[i], v[i - 1], nullptr));
    if (pyr[0].device < 0) SB_TRY(from_device(v[0], &pyr[0], nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

}  // extern "C"
