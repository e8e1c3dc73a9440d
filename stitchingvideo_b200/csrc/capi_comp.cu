// capi_comp.cu — ExposureCompensator::apply C ABI (exposure_compensate.hpp:51-101,
// exposure_compensate.cpp:150-153, 225-246).  feed() (gain estimation) is calibration and stays
// with the host caller; its result arrives through sb_comp_set_gains / sb_comp_set_gain_maps.
#include "sb_kernels.h"

using namespace sb;

struct sb_comp {
    int device = 0;
    int kind = 0;
    cudaStream_t stream = nullptr;
    std::vector<double> gains;
    std::vector<DevImage> gain_maps;       // block gain maps as fed (small)
    std::vector<DevImage> gain_full;       // resized to the image size, cached (sequence-constant)
    DevImage stage;
};

extern "C" {

int sb_comp_create(int kind, int device, sb_comp **out)
{
    if (!out) return fail(SB_ERR_ASSERT, "out is null");
    *out = nullptr;
    if (kind != SB_COMP_NO && kind != SB_COMP_GAIN && kind != SB_COMP_GAIN_BLOCKS)
        return fail(SB_ERR_BAD_ARG, "unsupported exposure compensation method");
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    sb_comp *c = new sb_comp;
    c->device = device; c->kind = kind;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(SB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = c;
    return SB_OK;
}

void sb_comp_destroy(sb_comp *c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    delete c;
}

int sb_comp_set_gains(sb_comp *c, const double *gains, int n)
{
    SB_ASSERT(c && gains && n >= 0);
    c->gains.assign(gains, gains + n);
    return SB_OK;
}

int sb_comp_get_gains(const sb_comp *c, double *gains, int n)
{
    SB_ASSERT(c && gains && n == (int)c->gains.size());
    std::copy(c->gains.begin(), c->gains.end(), gains);
    return SB_OK;
}

int sb_comp_set_gain_maps(sb_comp *c, const sb_image *maps, int n)
{
    SB_ASSERT(c && maps && n >= 0);
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    c->gain_maps.clear(); c->gain_full.clear();
    c->gain_maps.resize(n); c->gain_full.resize(n);
    for (int i = 0; i < n; ++i) {
        SB_ASSERT(maps[i].type == SB_32FC1 && maps[i].data);
        DImage v;
        DevImage tmp;
        SB_TRY(to_device(maps[i], tmp, c->stream, &v));
        SB_TRY(c->gain_maps[i].create(v.rows, v.cols, SB_32FC1));
        SB_TRY(launch_convert(v, c->gain_maps[i].v, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return SB_OK;
}

int sb_comp_apply(sb_comp *c, int index, sb_point /*corner*/, sb_image *image, const sb_image * /*mask*/)
{
    SB_ASSERT(c && image);
    if (c->kind == SB_COMP_NO) return SB_OK;
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_TRY(check_image(image, "image"));
    DImage d;
    SB_TRY(to_device(*image, c->stage, c->stream, &d));
    if (c->kind == SB_COMP_GAIN) {
        SB_ASSERT(index >= 0 && index < (int)c->gains.size());
        SB_ASSERT(type_depth(image->type) == SB_8U);
        SB_TRY(launch_scale_8u(d, (float)c->gains[index], c->stream));   // image *= gains_(index, 0)
    } else {
        SB_ASSERT(image->type == SB_8UC3);                               // exposure_compensate.cpp:227
        SB_ASSERT(index >= 0 && index < (int)c->gain_maps.size());
        const DImage &gm = c->gain_maps[index].v;
        const DImage *full = &gm;
        if (gm.rows != d.rows || gm.cols != d.cols) {
            DevImage &gf = c->gain_full[index];
            if (gf.v.rows != d.rows || gf.v.cols != d.cols || !gf.v.data) {   // resize once per geometry
                SB_TRY(gf.create(d.rows, d.cols, SB_32FC1));
                SB_TRY(launch_resize_linear_32f(gm, gf.v, c->stream));
            }
            full = &gf.v;
        }
        SB_TRY(launch_mul_map_8u(d, *full, c->stream));
    }
    if (image->device < 0) SB_TRY(from_device(d, image, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

}  // extern "C"
