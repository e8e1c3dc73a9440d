// sb_internal.h — shared host-side plumbing of libstitchb200 (error state, device buffers, staging).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/stitchb200.h"

namespace sb {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char *fmt, ...);

#define SB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return sb::fail(SB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define SB_TRY(expr)                  \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != SB_OK) return rc_; \
    } while (0)
// CV_Assert-shaped argument check
#define SB_ASSERT(cond)                                                                            \
    do {                                                                                           \
        if (!(cond)) return sb::fail(SB_ERR_ASSERT, "assertion failed: %s (%s:%d)", #cond, __FILE__, __LINE__); \
    } while (0)
// after every kernel launch: count it and surface launch-configuration errors
#define SB_LAUNCHED()                                                                              \
    do {                                                                                           \
        sb::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        cudaError_t e_ = cudaPeekAtLastError();                                                    \
        if (e_ != cudaSuccess)                                                                     \
            return sb::fail(SB_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int type_depth(int type) { return type & 7; }
inline int type_cn(int type) { return ((type >> 3) & 63) + 1; }
inline int depth_size(int depth) { return depth == SB_8U ? 1 : (depth == SB_16S || depth == SB_16U) ? 2 : depth == SB_32F ? 4 : 0; }
inline int elem_size(int type) { return depth_size(type_depth(type)) * type_cn(type); }
inline bool type_supported(int type)
{
    return type == SB_8UC1 || type == SB_8UC3 || type == SB_16SC1 || type == SB_16SC3 || type == SB_32FC1 ||
           type == SB_16UC1 || type == SB_16SC2;      // (the last two: fixed-point remap maps only)
}
inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Where MultiBandBlender::feed places an image of `w` x `h` whose top-left corner is `tl` inside the prepared ROI `roi`
// (blenders.cpp:241-269): the feed rect grows by a gap of 3 * 2^bands, is snapped to multiples of 2^bands relative to the ROI,
// padded to a multiple of 2^bands and shifted back inside.  tl_new = top-left of the padded rect (panorama coordinates),
// width / height its size, top / left / bottom / right the borders copyMakeBorder adds around the image.
struct FeedRect {
    sb_point tl_new;
    int width, height, top, left, bottom, right;
};
inline FeedRect multiband_feed_rect(const sb_rect &roi, sb_point tl, int w, int h, int num_bands)
{
    const int m = 1 << num_bands, gap = 3 * m;
    const int rbr_x = roi.x + roi.width, rbr_y = roi.y + roi.height;
    FeedRect f;
    f.tl_new = {(roi.x > tl.x - gap ? roi.x : tl.x - gap), (roi.y > tl.y - gap ? roi.y : tl.y - gap)};
    sb_point br = {(rbr_x < tl.x + w + gap ? rbr_x : tl.x + w + gap), (rbr_y < tl.y + h + gap ? rbr_y : tl.y + h + gap)};
    f.tl_new.x = roi.x + (((f.tl_new.x - roi.x) >> num_bands) << num_bands);
    f.tl_new.y = roi.y + (((f.tl_new.y - roi.y) >> num_bands) << num_bands);
    f.width = br.x - f.tl_new.x; f.height = br.y - f.tl_new.y;
    f.width += (m - f.width % m) % m;
    f.height += (m - f.height % m) % m;
    br.x = f.tl_new.x + f.width; br.y = f.tl_new.y + f.height;
    const int dy = br.y - rbr_y > 0 ? br.y - rbr_y : 0, dx = br.x - rbr_x > 0 ? br.x - rbr_x : 0;
    f.tl_new.x -= dx; br.x -= dx; f.tl_new.y -= dy; br.y -= dy;
    f.top = tl.y - f.tl_new.y; f.left = tl.x - f.tl_new.x;
    f.bottom = br.y - tl.y - h; f.right = br.x - tl.x - w;
    return f;
}

// Programmatic dependent launch (sm_90+): the grid may be set up and its CTAs made resident while its predecessor in the
// stream is still draining; the kernel calls pdl_wait() (sb_tma.cuh) before it touches memory, so the stream's order of
// memory effects is unchanged - only the launch latency between two dependent kernels (2.5-3 us on B200) is overlapped.
// SB_PDL=0 in the environment launches plainly (A/B measurements).
inline bool pdl_enabled()
{
    static const bool on = !(getenv("SB_PDL") && atoi(getenv("SB_PDL")) == 0);
    return on;
}
inline cudaError_t launch_pdl_c(const void *fn, dim3 grid, dim3 block, void **params, size_t smem, cudaStream_t s)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (!pdl_enabled() || cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
        return cudaLaunchKernel(fn, grid, block, params, smem, s);       // (a stream being captured into a graph keeps plain kernel nodes)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, params);
}
template <typename A> inline cudaError_t launch_pdl(void (*k)(A), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const A &a)
{
    void *params[] = {const_cast<A *>(&a)};
    return launch_pdl_c(reinterpret_cast<const void *>(k), grid, block, params, smem, s);
}

// device image view
struct DImage {
    void *data = nullptr;
    int rows = 0, cols = 0, type = 0;
    size_t step = 0;
    template <typename T> T *ptr() const { return static_cast<T *>(data); }
    bool empty() const { return data == nullptr || rows <= 0 || cols <= 0; }
};

// grow-only device allocation (kept across frames: no cudaMalloc on the per-frame path)
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    ~DevBuf() { release(); }
    int ensure(size_t bytes);
    void release();
};

// owned device image with a 256-byte aligned pitch
struct DevImage {
    DevBuf buf;
    DImage v;
    int create(int rows, int cols, int type);   // contents undefined
    // rows packed back to back (pitch = row bytes rounded up to 16): host<->device transfers of
    // contiguous host images become one linear DMA instead of a pitched 2D copy
    int create_packed(int rows, int cols, int type);
    int create_zero(int rows, int cols, int type, cudaStream_t s);
};

inline size_t aligned_step(int cols, int type) { return ((size_t)cols * elem_size(type) + 255) & ~(size_t)255; }

int check_image(const sb_image *img, const char *what);
// Device view of an sb_image: zero copy for device images, else staged H2D into `stage`.
int to_device(const sb_image &img, DevImage &stage, cudaStream_t s, DImage *out);
// Copy a device image into the caller's sb_image (host or device); sizes/types must match.
int from_device(const DImage &src, sb_image *dst, cudaStream_t s);
// Point a caller sb_image with data == NULL at a handle-owned device image.
void lend(const DImage &src, int device, sb_image *dst);

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device);
    ~DeviceGuard();
};

}  // namespace sb
