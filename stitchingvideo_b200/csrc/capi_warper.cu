// capi_warper.cu — RotationWarper C ABI (include/stitchb200.h) over the warp kernels.
//
// Host side of the warper: ProjectorBase::setCameraParams (warpers.cpp:50-78), mapForward
// (warpers_inl.hpp:206-218,237-247,271-280), detectResultRoi (warpers.cpp:139-212,
// warpers_inl.hpp:169-203).  These are O(1) / O(perimeter) per calibration and stay on the host
// (SURVEY.md §8a a1, a2); per-pixel work (a3, a4) is CUDA.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "sb_kernels.h"
#include "sb_host_projector.h"

using namespace sb;

struct sb_warper {
    int device = 0;
    int kind = 0;
    float scale = 1.f;
    float t[3] = {0, 0, 0};
    float a = 1.f, b = 1.f;      // CompressedRectilinear / Panini parameters
    cudaStream_t stream = nullptr;
    // maps cached by the last build_maps (the app's xmap1/ymap1, APP64:188-198)
    DevImage xmap, ymap;
    bool have_maps = false;
    // what those maps were built for: a repeated warp() / buildMaps with the same (size, K, R, scale, T, a, b) reuses them
    // (warpers_inl.hpp:88-99 rebuilds every call; for the projectors evaluated on the host that is a per-pixel libm loop)
    struct Key { int kind, w, h; float scale, t[3], a, b, K[9], R[9]; } key{};
    sb_point key_tl{}, key_br{};
    DevImage bxmap, bymap;       // warpBackward's forward maps, cached the same way
    Key bkey{};
    bool have_bmaps = false;
    // staging / outputs
    DevImage src_stage, dst_out;
};

namespace {
int kind_ok(int kind) { return kind >= SB_WARP_PLANE && kind <= SB_WARP_PLANE_PORTRAIT; }
bool device_maps(int kind) { return kind == SB_WARP_PLANE || kind == SB_WARP_CYLINDRICAL || kind == SB_WARP_SPHERICAL; }
void set_params(const sb_warper *w, ProjParams &p, const float K[9], const float R[9])
{
    projector_set(p, w->kind, w->scale, K, R, w->t);
    p.a = w->a; p.b = w->b;
}

sb_warper::Key make_key(const sb_warper *w, sb_size size, const float K[9], const float R[9])
{
    sb_warper::Key k;
    std::memset(&k, 0, sizeof k);
    k.kind = w->kind; k.w = size.width; k.h = size.height; k.scale = w->scale; k.a = w->a; k.b = w->b;
    std::memcpy(k.t, w->t, sizeof k.t); std::memcpy(k.K, K, sizeof k.K); std::memcpy(k.R, R, sizeof k.R);
    return k;
}

int build_maps_dev(sb_warper *w, sb_size src_size, const float K[9], const float R[9], sb_point *tl, sb_point *br)
{
    SB_ASSERT(K && R);
    SB_ASSERT(src_size.width > 0 && src_size.height > 0);
    const sb_warper::Key key = make_key(w, src_size, K, R);
    if (w->have_maps && std::memcmp(&key, &w->key, sizeof key) == 0) { *tl = w->key_tl; *br = w->key_br; return SB_OK; }
    w->have_maps = false;
    ProjParams p;
    set_params(w, p, K, R);
    projector_detect_result_roi(p, src_size.width, src_size.height, tl, br);
    long long mw = (long long)br->x - tl->x + 1, mh = (long long)br->y - tl->y + 1;
    if (mw <= 0 || mh <= 0 || mw * mh > (1LL << 31))
        return fail(SB_ERR_ASSERT, "degenerate warped ROI %lld x %lld (check K, R, scale)", mw, mh);
    SB_TRY(w->xmap.create((int)mh, (int)mw, SB_32FC1));
    SB_TRY(w->ymap.create((int)mh, (int)mw, SB_32FC1));
    if (device_maps(w->kind)) {
        SB_TRY(launch_build_maps(p, tl->x, tl->y, w->xmap.v, w->ymap.v, w->stream));
    } else {      // libm-heavy projectors: maps built on the host once per calibration (as the reference does), uploaded
        std::vector<float> hx((size_t)mw * mh), hy(hx.size());
        projector_build_maps_host(p, *tl, *br, hx.data(), hy.data());
        SB_CUDA(cudaMemcpy2DAsync(w->xmap.v.data, w->xmap.v.step, hx.data(), (size_t)mw * 4, (size_t)mw * 4, (size_t)mh, cudaMemcpyHostToDevice, w->stream));
        SB_CUDA(cudaMemcpy2DAsync(w->ymap.v.data, w->ymap.v.step, hy.data(), (size_t)mw * 4, (size_t)mw * 4, (size_t)mh, cudaMemcpyHostToDevice, w->stream));
        SB_CUDA(cudaStreamSynchronize(w->stream));
    }
    w->have_maps = true;
    w->key = key; w->key_tl = *tl; w->key_br = *br;
    return SB_OK;
}

int remap_out(sb_warper *w, const sb_image *src, const DImage &xm, const DImage &ym, int interp, int border, sb_image *dst)
{
    SB_TRY(check_image(src, "src"));
    SB_ASSERT(dst != nullptr);
    SB_ASSERT(src->type == SB_8UC1 || src->type == SB_8UC3);
    DImage dsrc;
    SB_TRY(to_device(*src, w->src_stage, w->stream, &dsrc));
    const bool direct = dst->data && dst->device >= 0;
    DImage out;
    if (direct) {
        SB_ASSERT(dst->rows == xm.rows && dst->cols == xm.cols && dst->type == src->type);
        out.data = dst->data; out.rows = dst->rows; out.cols = dst->cols; out.type = dst->type; out.step = dst->step;
    } else {
        SB_TRY(w->dst_out.create(xm.rows, xm.cols, src->type));
        out = w->dst_out.v;
    }
    SB_TRY(launch_remap(dsrc, out, xm, ym, interp, border, nullptr, w->stream));
    if (!dst->data) lend(out, w->device, dst);
    else if (!direct) SB_TRY(from_device(out, dst, w->stream));
    SB_CUDA(cudaStreamSynchronize(w->stream));
    return SB_OK;
}
}  // namespace

extern "C" {

int sb_warper_create(int kind, float scale, int device, sb_warper **out)
{
    if (!out) return fail(SB_ERR_ASSERT, "out is null");
    *out = nullptr;
    if (!kind_ok(kind)) return fail(SB_ERR_BAD_ARG, "unsupported warper kind %d", kind);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_warper *w = new sb_warper;
    w->device = device; w->kind = kind; w->scale = scale;
    cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete w; return fail(SB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = w;
    return SB_OK;
}

void sb_warper_destroy(sb_warper *w)
{
    if (!w) return;
    DeviceGuard g(w->device);
    if (w->stream) { cudaStreamSynchronize(w->stream); cudaStreamDestroy(w->stream); }
    delete w;
}

float sb_warper_get_scale(const sb_warper *w) { return w ? w->scale : 0.f; }
int sb_warper_set_scale(sb_warper *w, float scale) { SB_ASSERT(w); w->scale = scale; return SB_OK; }
int sb_warper_set_translation(sb_warper *w, const float T[3])
{
    SB_ASSERT(w && T);
    w->t[0] = T[0]; w->t[1] = T[1]; w->t[2] = T[2];
    return SB_OK;
}

int sb_warper_set_ab(sb_warper *w, float a, float b)
{
    SB_ASSERT(w);
    w->a = a; w->b = b;
    return SB_OK;
}

int sb_warper_warp_point(sb_warper *w, const float pt[2], const float K[9], const float R[9], float uv[2])
{
    SB_ASSERT(w && pt && K && R && uv);
    ProjParams p;
    set_params(w, p, K, R);
    projector_map_forward(p, pt[0], pt[1], &uv[0], &uv[1]);
    return SB_OK;
}

int sb_warper_warp_roi(sb_warper *w, sb_size src_size, const float K[9], const float R[9], sb_rect *roi)
{
    SB_ASSERT(w && K && R && roi);
    ProjParams p;
    set_params(w, p, K, R);
    sb_point tl, br;
    projector_detect_result_roi(p, src_size.width, src_size.height, &tl, &br);
    roi->x = tl.x; roi->y = tl.y; roi->width = br.x + 1 - tl.x; roi->height = br.y + 1 - tl.y;
    return SB_OK;
}

int sb_warper_build_maps(sb_warper *w, sb_size src_size, const float K[9], const float R[9], sb_image *xmap, sb_image *ymap, sb_rect *roi)
{
    SB_ASSERT(w && roi);
    DeviceGuard g(w->device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_point tl, br;
    SB_TRY(build_maps_dev(w, src_size, K, R, &tl, &br));
    roi->x = tl.x; roi->y = tl.y; roi->width = br.x - tl.x; roi->height = br.y - tl.y;   // Rect(dst_tl, dst_br)
    sb_image *outs[2] = {xmap, ymap};
    const DImage *maps[2] = {&w->xmap.v, &w->ymap.v};
    for (int i = 0; i < 2; ++i) {
        if (!outs[i]) continue;
        if (!outs[i]->data) lend(*maps[i], w->device, outs[i]);
        else SB_TRY(from_device(*maps[i], outs[i], w->stream));
    }
    SB_CUDA(cudaStreamSynchronize(w->stream));
    return SB_OK;
}

int sb_warper_warp(sb_warper *w, const sb_image *src, const float K[9], const float R[9], int interp_mode, int border_mode, sb_image *dst, sb_point *tl)
{
    SB_ASSERT(w && src && dst);
    DeviceGuard g(w->device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_point tl_, br_;
    sb_size ss = {src->cols, src->rows};
    SB_TRY(build_maps_dev(w, ss, K, R, &tl_, &br_));
    if (tl) *tl = tl_;
    return remap_out(w, src, w->xmap.v, w->ymap.v, interp_mode, border_mode, dst);
}

int sb_warper_remap(sb_warper *w, const sb_image *src, int interp_mode, int border_mode, sb_image *dst)
{
    SB_ASSERT(w && src && dst);
    if (!w->have_maps) return fail(SB_ERR_ASSERT, "sb_warper_remap: no cached maps; call sb_warper_build_maps first");
    DeviceGuard g(w->device);
    if (!g.ok) return SB_ERR_CUDA;
    return remap_out(w, src, w->xmap.v, w->ymap.v, interp_mode, border_mode, dst);
}

int sb_warper_warp_backward(sb_warper *w, const sb_image *src, const float K[9], const float R[9], int interp_mode, int border_mode, sb_size dst_size, sb_image *dst)
{
    // warpers_inl.hpp:102-128.  Not on the per-frame path (no caller in the reference apps): the
    // mapForward maps (atan2f/acosf) are built on the host, the remap runs on the device.
    SB_ASSERT(w && src && dst && K && R);
    DeviceGuard g(w->device);
    if (!g.ok) return SB_ERR_CUDA;
    ProjParams p;
    set_params(w, p, K, R);
    sb_point tl, br;
    projector_detect_result_roi(p, dst_size.width, dst_size.height, &tl, &br);
    SB_ASSERT(br.x - tl.x + 1 == src->cols && br.y - tl.y + 1 == src->rows);
    const sb_warper::Key key = make_key(w, dst_size, K, R);
    if (!(w->have_bmaps && std::memcmp(&key, &w->bkey, sizeof key) == 0)) {
        w->have_bmaps = false;
        std::vector<float> hx((size_t)dst_size.width * dst_size.height), hy(hx.size());
        for (int y = 0; y < dst_size.height; ++y)
            for (int x = 0; x < dst_size.width; ++x) {
                float u, v;
                projector_map_forward(p, (float)x, (float)y, &u, &v);
                hx[(size_t)y * dst_size.width + x] = u - tl.x;
                hy[(size_t)y * dst_size.width + x] = v - tl.y;
            }
        sb_image ix = {hx.data(), dst_size.height, dst_size.width, SB_32FC1, (size_t)dst_size.width * 4, -1};
        sb_image iy = {hy.data(), dst_size.height, dst_size.width, SB_32FC1, (size_t)dst_size.width * 4, -1};
        DImage t;
        SB_TRY(to_device(ix, w->bxmap, w->stream, &t));
        SB_TRY(to_device(iy, w->bymap, w->stream, &t));
        SB_CUDA(cudaStreamSynchronize(w->stream));           // (hx / hy go out of scope)
        w->bkey = key; w->have_bmaps = true;
    }
    const DImage vx = w->bxmap.v, vy = w->bymap.v;
    int rc = remap_out(w, src, vx, vy, interp_mode, border_mode, dst);
    cudaStreamSynchronize(w->stream);
    return rc;
}

int sb_remap(const sb_image *src, sb_image *dst, const sb_image *xmap, const sb_image *ymap, int interp_mode, int border_mode, const uint8_t border_value[4], int device)
{
    SB_ASSERT(src && dst && xmap && ymap);
    SB_ASSERT(dst->data != nullptr);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    DevImage s_src, s_x, s_y, s_dst;
    DImage dsrc, dxm, dym, out;
    cudaStream_t s = nullptr;   // legacy default stream: one-shot helper
    SB_TRY(to_device(*src, s_src, s, &dsrc));
    SB_TRY(to_device(*xmap, s_x, s, &dxm));
    if (ymap->data) SB_TRY(to_device(*ymap, s_y, s, &dym));
    SB_ASSERT(dst->rows == dxm.rows && dst->cols == dxm.cols && dst->type == src->type);
    if (dst->device >= 0) { out.data = dst->data; out.rows = dst->rows; out.cols = dst->cols; out.type = dst->type; out.step = dst->step; }
    else { SB_TRY(s_dst.create(dst->rows, dst->cols, dst->type)); out = s_dst.v; }
    SB_TRY(launch_remap(dsrc, out, dxm, dym, interp_mode, border_mode, border_value, s));
    if (dst->device < 0) SB_TRY(from_device(out, dst, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}

int sb_convert_maps(const sb_image *xmap, const sb_image *ymap, sb_image *map1, sb_image *map2, int nn_interpolation, int device)
{
    SB_ASSERT(xmap && ymap && map1 && map1->data);
    SB_ASSERT(nn_interpolation || (map2 && map2->data));
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    DevImage s_x, s_y, s_1, s_2;
    DImage dx, dy, d1, d2;
    cudaStream_t s = nullptr;
    SB_TRY(to_device(*xmap, s_x, s, &dx));
    SB_TRY(to_device(*ymap, s_y, s, &dy));
    SB_ASSERT(map1->type == SB_16SC2 && map1->rows == dx.rows && map1->cols == dx.cols);
    if (map1->device >= 0) { d1.data = map1->data; d1.rows = map1->rows; d1.cols = map1->cols; d1.type = map1->type; d1.step = map1->step; }
    else { SB_TRY(s_1.create(map1->rows, map1->cols, SB_16SC2)); d1 = s_1.v; }
    if (!nn_interpolation) {
        SB_ASSERT(map2->type == SB_16UC1 && map2->rows == dx.rows && map2->cols == dx.cols);
        if (map2->device >= 0) { d2.data = map2->data; d2.rows = map2->rows; d2.cols = map2->cols; d2.type = map2->type; d2.step = map2->step; }
        else { SB_TRY(s_2.create(map2->rows, map2->cols, SB_16UC1)); d2 = s_2.v; }
    }
    SB_TRY(launch_convert_maps(dx, dy, d1, d2, nn_interpolation != 0, s));
    if (map1->device < 0) SB_TRY(from_device(d1, map1, s));
    if (!nn_interpolation && map2->device < 0) SB_TRY(from_device(d2, map2, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}

}  // extern "C"
