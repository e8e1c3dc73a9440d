// sb_pyr.cuh — pyrUp tap arithmetic shared by the pyramid and blend kernels.
#pragma once
#include "sb_device.cuh"

namespace sb {
using namespace sbd;

template <typename T> __device__ __forceinline__ const T *crow(const void *base, size_t step, int y)
{
    return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) + (size_t)y * step);
}
template <typename T> __device__ __forceinline__ T *mrow(void *base, size_t step, int y)
{
    return reinterpret_cast<T *>(reinterpret_cast<char *>(base) + (size_t)y * step);
}

// Value of pyrUp(coarse)(y, x) channel c BEFORE the final cast: sum of <= 9 taps, weights sum to 64.
template <typename T, int CN>
__device__ __forceinline__ int pyr_up_sum(const T *__restrict__ coarse, size_t cstep, int cw, int ch, int y, int x, int c)
{
    const int cx = x >> 1, cy = y >> 1;
    const int xl = cx == 0 ? (cw > 1 ? 1 : 0) : cx - 1, xr = min(cx + 1, cw - 1);
    const int yt = cy == 0 ? (ch > 1 ? 1 : 0) : cy - 1, yb = min(cy + 1, ch - 1);
    const T *r1 = crow<T>(coarse, cstep, cy), *r2 = crow<T>(coarse, cstep, yb);
    int h1, h2;
    if (x & 1) {
        h1 = ((int)r1[cx * CN + c] + (int)r1[xr * CN + c]) * 4;
        h2 = ((int)r2[cx * CN + c] + (int)r2[xr * CN + c]) * 4;
    } else {
        h1 = (int)r1[xl * CN + c] + (int)r1[cx * CN + c] * 6 + (int)r1[xr * CN + c];
        h2 = (int)r2[xl * CN + c] + (int)r2[cx * CN + c] * 6 + (int)r2[xr * CN + c];
    }
    if (y & 1) return (h1 + h2) * 4;
    const T *r0 = crow<T>(coarse, cstep, yt);
    int h0 = (x & 1) ? ((int)r0[cx * CN + c] + (int)r0[xr * CN + c]) * 4
                     : (int)r0[xl * CN + c] + (int)r0[cx * CN + c] * 6 + (int)r0[xr * CN + c];
    return h1 * 6 + h0 + h2;
}
template <typename T> __device__ __forceinline__ int up_cast(int v);
template <> __device__ __forceinline__ int up_cast<uint8_t>(int v) { return sat_u8((v + 32) >> 6); }
template <> __device__ __forceinline__ int up_cast<short>(int v) { return sat_s16((v + 32) >> 6); }


}  // namespace sb
