// capi_common.cu — error state, device buffers, host<->device staging shared by the C ABI.
#include "sb_internal.h"
#include "sb_fused.h"

namespace sb {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (code == SB_ERR_CUDA) cudaGetLastError();   // a reported CUDA error must not resurface at the next launch check of an unrelated call
    return code;
}

int DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap && p) return SB_OK;
    release();
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cap = 0;
        return fail(e == cudaErrorMemoryAllocation ? SB_ERR_NO_MEM : SB_ERR_CUDA, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return SB_OK;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

int DevImage::create(int rows, int cols, int type)
{
    size_t step = aligned_step(cols, type);
    SB_TRY(buf.ensure(step * (size_t)(rows > 0 ? rows : 1)));
    v.data = buf.p; v.rows = rows; v.cols = cols; v.type = type; v.step = step;
    return SB_OK;
}

int DevImage::create_packed(int rows, int cols, int type)
{
    size_t step = ((size_t)cols * elem_size(type) + 15) & ~(size_t)15;
    SB_TRY(buf.ensure(step * (size_t)(rows > 0 ? rows : 1) + 16));
    v.data = buf.p; v.rows = rows; v.cols = cols; v.type = type; v.step = step;
    return SB_OK;
}

int DevImage::create_zero(int rows, int cols, int type, cudaStream_t s)
{
    SB_TRY(create(rows, cols, type));
    SB_CUDA(cudaMemsetAsync(v.data, 0, v.step * (size_t)(rows > 0 ? rows : 1), s));
    return SB_OK;
}

int check_image(const sb_image *img, const char *what)
{
    if (!img) return fail(SB_ERR_ASSERT, "%s: null image", what);
    if (!type_supported(img->type)) return fail(SB_ERR_ASSERT, "%s: unsupported type %d", what, img->type);
    if (img->rows < 0 || img->cols < 0) return fail(SB_ERR_ASSERT, "%s: negative size", what);
    if (img->data && img->step < (size_t)img->cols * elem_size(img->type))
        return fail(SB_ERR_ASSERT, "%s: step %zu smaller than a row", what, img->step);
    return SB_OK;
}

int to_device(const sb_image &img, DevImage &stage, cudaStream_t s, DImage *out)
{
    SB_TRY(check_image(&img, "input image"));
    if (!img.data) return fail(SB_ERR_ASSERT, "input image has no data");
    if (img.device >= 0) {
        out->data = img.data; out->rows = img.rows; out->cols = img.cols; out->type = img.type; out->step = img.step;
        return SB_OK;
    }
    SB_TRY(stage.create_packed(img.rows, img.cols, img.type));
    if (img.rows > 0 && img.cols > 0) {
        if (img.step == stage.v.step && img.step == (size_t)img.cols * elem_size(img.type))      // packed on both sides: one linear DMA
            SB_CUDA(cudaMemcpyAsync(stage.v.data, img.data, img.step * (size_t)img.rows, cudaMemcpyHostToDevice, s));
        else
            SB_CUDA(cudaMemcpy2DAsync(stage.v.data, stage.v.step, img.data, img.step, (size_t)img.cols * elem_size(img.type), img.rows, cudaMemcpyHostToDevice, s));
    }
    *out = stage.v;
    return SB_OK;
}

int from_device(const DImage &src, sb_image *dst, cudaStream_t s)
{
    SB_TRY(check_image(dst, "output image"));
    if (!dst->data) return fail(SB_ERR_ASSERT, "output image has no data");
    if (dst->rows != src.rows || dst->cols != src.cols || dst->type != src.type)
        return fail(SB_ERR_ASSERT, "output image is %dx%d type %d, expected %dx%d type %d", dst->rows, dst->cols, dst->type, src.rows, src.cols, src.type);
    if (src.rows > 0 && src.cols > 0) {
        const cudaMemcpyKind kind = dst->device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        // rows packed back to back on both sides: one linear DMA.  (Equal pitches alone are not enough: a column view of a
        // wider image has the wide image's pitch, and a linear copy would run over the columns next to the view.)
        if (dst->step == src.step && src.step == (size_t)src.cols * elem_size(src.type))
            SB_CUDA(cudaMemcpyAsync(dst->data, src.data, src.step * (size_t)src.rows, kind, s));
        else
            SB_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, src.data, src.step, (size_t)src.cols * elem_size(src.type), src.rows, kind, s));
    }
    return SB_OK;
}

void lend(const DImage &src, int device, sb_image *dst)
{
    dst->data = src.data; dst->rows = src.rows; dst->cols = src.cols; dst->type = src.type; dst->step = src.step;
    dst->device = device;
}

DeviceGuard::DeviceGuard(int device)
{
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(device) == cudaSuccess;
    if (!ok) fail(SB_ERR_CUDA, "cudaSetDevice(%d) failed: no usable CUDA device (there is no CPU fallback)", device);
}
DeviceGuard::~DeviceGuard()
{
    if (prev >= 0 && ok) cudaSetDevice(prev);
}

}  // namespace sb

extern "C" {
const char *sb_last_error(void) { return sb::g_last_error.c_str(); }
const char *sb_version(void) { return "stitchb200 0.1 (sm_100a)"; }
uint64_t sb_kernel_launch_count(void) { return sb::g_launches.load(); }
int sb_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return sb::fail(SB_ERR_ASSERT, "ptr is null");
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return sb::fail(SB_ERR_NO_MEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return SB_OK;
}
void sb_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }
int sb_selftest_division(int device, unsigned long long n, unsigned seed, unsigned long long *mismatches)
{
    if (!mismatches) return sb::fail(SB_ERR_ASSERT, "mismatches is null");
    sb::DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    return sb::selftest_division(n, seed, mismatches);
}
int sb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
}
