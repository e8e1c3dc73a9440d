/*
 * so_prims.c — CPU ORACLE (test infrastructure, NOT product code; see stitch_oracle.h).
 *
 * OpenCV 2.4.11 core/imgproc primitive semantics that the reference's compositing
 * path calls.  The 2.4.11 library sources are not in
 * /root/reference (only opencv_core2411.dll / opencv_imgproc2411.dll are), so each
 * function restates the published algorithm (imgwarp.cpp remap, pyramids.cpp
 * pyrDown_/pyrUp_, copy.cpp copyMakeBorder, convert.cpp cvtScale, arithm.cpp
 * add/subtract, distransform.cpp distanceTransform_3x3, imgwarp.cpp resize) and is
 * anchored on the reference's call sites, cited per function.
 * Build: gcc -O2 -ffp-contract=off (no fast-math: every float op is one IEEE op).
 */
#include "stitch_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SO_DEPTH(t) ((t) & 7)
#define SO_CN(t) ((((t) >> 3) & 63) + 1)

static inline int so_elem_size1(int type) {
    switch (SO_DEPTH(type)) { case SO_8U: return 1; case SO_16S: return 2; case SO_32F: return 4; }
    return 0;
}
#define ROW(m, T, y) ((T *)((char *)(m)->data + (size_t)(y) * (m)->step))

const char *so_version(void) { return "stitch-oracle 1 (OpenCV 2.4.11 cv::detail compositing path restatement)"; }

/* ------------------------------------------------------------------ scalars */

/* cvRound(double) of OpenCV 2.4.11 on x86-64 = _mm_cvtsd_si32: round-half-even,
 * "integer indefinite" INT_MIN when out of range or NaN. */
int so_cvround(float v)
{
    if (!(fabsf(v) < 2147483648.0f)) return INT_MIN;
    return (int)nearbyintf(v);          /* default rounding mode: to nearest even */
}

/* static_cast<short>(float) as compiled for x86-64 (cvttss2si r32 then low word):
 * blenders.cpp:139-141,321-323,400-402. */
short so_trunc_short(float v)
{
    int i;
    if (!(fabsf(v) < 2147483648.0f)) i = INT_MIN; else i = (int)v;   /* trunc toward zero */
    return (short)(unsigned short)((unsigned)i & 0xffffu);
}

static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static inline short sat_s16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* cv::borderInterpolate (imgproc filter.cpp) — used by remap, pyrDown, pyrUp, copyMakeBorder */
int so_border_interpolate(int p, int len, int border)
{
    if ((unsigned)p < (unsigned)len) return p;
    switch (border) {
    case SO_BORDER_REPLICATE: return p < 0 ? 0 : len - 1;
    case SO_BORDER_REFLECT:
    case SO_BORDER_REFLECT_101: {
        int delta = border == SO_BORDER_REFLECT_101;
        if (len == 1) return 0;
        do {
            if (p < 0) p = -p - 1 + delta;
            else p = len - 1 - (p - len) - delta;
        } while ((unsigned)p >= (unsigned)len);
        return p;
    }
    case SO_BORDER_WRAP:
        if (p < 0) p -= ((p - len + 1) / len) * len;
        if (p >= len) p %= len;
        return p;
    default: return -1;  /* BORDER_CONSTANT */
    }
}

/* ------------------------------------------------------------------ sinf / cosf
 * Portable restatement of glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c,
 * s_cosf.c, sincosf.h; x86_64 multiarch FMA variant = the one cv2/libm use on any
 * FMA-capable host).  Operation order and the places where the compiler fused
 * multiply-adds were read from the libm.so.6 disassembly; constants from its
 * .rodata.  warpers_inl.hpp:256-259,289-291 call sinf/cosf per pixel; the GPU
 * map-build kernel implements this same sequence so that maps are bit-exact. */
static const double SC_HPI_INV = 0x1.45F306DC9C883p+23;   /* 2/pi * 2^24 */
static const double SC_HPI = 0x1.921FB54442D18p0;
static const double SC_PI63 = 0x1.921FB54442D18p-62;
static const double SC_C[5] = { 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5,
                                -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16 };
static const double SC_S[3] = { -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 };
static const uint32_t SC_INV_PIO4[24] = {
    0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 };

static inline uint32_t sc_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t sc_abstop12(float x) { return (sc_asuint(x) >> 20) & 0x7ff; }

/* neg = 1 selects __sincosf_table[1] (cosine polynomial negated) */
static inline float sc_poly(double x, double x2, int neg, int n)
{
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = fma(x2, SC_S[2], SC_S[1]);
        double x7 = x3 * x2;
        double s = fma(x3, SC_S[0], x);
        return (float)fma(s1, x7, s);
    } else {
        double sg = neg ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c1 = fma(x2, sg * SC_C[1], sg * SC_C[0]);
        double c2 = fma(x2, sg * SC_C[4], sg * SC_C[3]);
        double x6 = x4 * x2;
        double c = fma(x4, sg * SC_C[2], c1);
        return (float)fma(c2, x6, c);
    }
}

static inline double sc_reduce_fast(double x, int *np)
{
    double r = x * SC_HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, SC_HPI, x);
}

static inline double sc_reduce_large(uint32_t xi, int *np)
{
    const uint32_t *arr = &SC_INV_PIO4[(xi >> 26) & 15];
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    *np = (int)n;
    return (double)(int64_t)res0 * SC_PI63;
}

static const double SC_SIGN[4] = { 1.0, -1.0, -1.0, 1.0 };

static float sc_sincos(float y, int want_cos)
{
    double x = y, s;
    int n;
    if (sc_abstop12(y) < sc_abstop12(0x1.921FB6p-1f)) {
        s = x * x;
        if (sc_abstop12(y) < sc_abstop12(0x1p-12f)) return want_cos ? 1.0f : y;
        return sc_poly(x, s, 0, want_cos);
    } else if (sc_abstop12(y) < sc_abstop12(120.0f)) {
        x = sc_reduce_fast(x, &n);
        s = SC_SIGN[n & 3];
        return sc_poly(x * s, x * x, (n & 2) != 0, n ^ want_cos);
    } else if (sc_abstop12(y) < sc_abstop12(INFINITY)) {
        uint32_t xi = sc_asuint(y);
        int sign = xi >> 31;
        x = sc_reduce_large(xi, &n);
        s = SC_SIGN[(n + sign) & 3];
        return sc_poly(x * s, x * x, ((n + sign) & 2) != 0, n ^ want_cos);
    }
    return y - y;   /* NaN for inf/NaN input (__math_invalidf) */
}
float so_sinf(float x) { return sc_sincos(x, 0); }
float so_cosf(float x) { return sc_sincos(x, 1); }

/* ------------------------------------------------------------------ remap
 * cv::remap, 8U, CV_32FC1 x/y maps (imgwarp.cpp RemapInvoker + remapBilinear /
 * remapNearest; Appendix A1).  Call sites: warpers_inl.hpp:96,127; APP64:752. */
static void remap_weights(int fx, int fy, int w[4])
{
    /* initInterTab2D(INTER_LINEAR, fixpt): saturate_cast<short>(float tab * 32768) then the
     * sum-to-32768 fix-up, which only fires for (0,0): {32767,0,0,1}. */
    w[0] = (32 - fx) * (32 - fy) * 32; w[1] = fx * (32 - fy) * 32;
    w[2] = (32 - fx) * fy * 32;        w[3] = fx * fy * 32;
    if (fx == 0 && fy == 0) { w[0] = 32767; w[3] = 1; }
}

int so_remap(const so_mat *src, so_mat *dst, const so_mat *xmap, const so_mat *ymap,
             int interp, int border, const uint8_t border_value[4])
{
    static const uint8_t zero4[4] = { 0, 0, 0, 0 };
    const uint8_t *cval = border_value ? border_value : zero4;
    int cn = SO_CN(src->type);
    if (SO_DEPTH(src->type) != SO_8U || dst->type != src->type) return -1;
    /* fixed-point maps (cv::convertMaps / initUndistortRectifyMap(CV_16SC2), the app's video front end APP64:201-238,
     * 741): map1 CV_16SC2 = integer coordinates, map2 CV_16UC1 (optional) = fy * 32 + fx */
    const int fixed = xmap->type == SO_16SC2;
    if (fixed) {
        if (ymap && ymap->data && (ymap->type != SO_16UC1 || ymap->rows != xmap->rows || ymap->cols != xmap->cols)) return -1;
        if (interp != SO_INTER_NEAREST && !(ymap && ymap->data)) return -1;   /* OpenCV needs the fractional map for INTER_LINEAR */
    } else if (xmap->type != SO_32FC1 || !ymap || ymap->type != SO_32FC1) return -1;
    if (dst->rows != xmap->rows || dst->cols != xmap->cols) return -1;
    int W = src->cols, H = src->rows;

    for (int dy = 0; dy < dst->rows; ++dy) {
        const float *mx = fixed ? NULL : ROW(xmap, const float, dy), *my = fixed ? NULL : ROW(ymap, const float, dy);
        const short *m1 = fixed ? ROW(xmap, const short, dy) : NULL;
        const unsigned short *m2 = (fixed && ymap && ymap->data) ? ROW(ymap, const unsigned short, dy) : NULL;
        uint8_t *D = ROW(dst, uint8_t, dy);
        for (int dx = 0; dx < dst->cols; ++dx, D += cn) {
            if (interp == SO_INTER_NEAREST) {
                int sx = fixed ? m1[2 * dx] : sat_s16(so_cvround(mx[dx])), sy = fixed ? m1[2 * dx + 1] : sat_s16(so_cvround(my[dx]));
                if (m2) {   /* OpenCV's NNDeltaTab_i[fy * 32 + fx] = {fx < 16, fy < 16} is added to the integer coordinate
                             * in int16 arithmetic (sic: the table is set up that way in imgwarp.cpp's initInterTab2D) */
                    const int a = m2[dx] & 1023;
                    sx = (short)(sx + ((a & 31) < 16)); sy = (short)(sy + ((a >> 5) < 16));
                }
                if ((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H) {
                    memcpy(D, ROW(src, const uint8_t, sy) + sx * cn, cn);
                } else if (border == SO_BORDER_REPLICATE) {
                    sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx);
                    sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
                    memcpy(D, ROW(src, const uint8_t, sy) + sx * cn, cn);
                } else if (border == SO_BORDER_CONSTANT) {
                    memcpy(D, cval, cn);
                } else {
                    sx = so_border_interpolate(sx, W, border);
                    sy = so_border_interpolate(sy, H, border);
                    memcpy(D, ROW(src, const uint8_t, sy) + sx * cn, cn);
                }
                continue;
            }
            /* INTER_LINEAR: 5 fractional bits, float32 multiply before rounding */
            int w[4], sx, sy;
            if (fixed) {
                const int fxy = m2 ? (m2[dx] & 1023) : 0;
                remap_weights(fxy & 31, fxy >> 5, w);
                sx = m1[2 * dx]; sy = m1[2 * dx + 1];
            } else {
                int fsx = so_cvround(mx[dx] * 32), fsy = so_cvround(my[dx] * 32);
                remap_weights(fsx & 31, fsy & 31, w);
                sx = sat_s16(fsx >> 5); sy = sat_s16(fsy >> 5);
            }
            int x0, x1, y0, y1;
            if (border == SO_BORDER_CONSTANT && (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0)) {
                memcpy(D, cval, cn);
                continue;
            }
            x0 = so_border_interpolate(sx, W, border);  x1 = so_border_interpolate(sx + 1, W, border);
            y0 = so_border_interpolate(sy, H, border);  y1 = so_border_interpolate(sy + 1, H, border);
            for (int k = 0; k < cn; ++k) {
                int v0 = (x0 >= 0 && y0 >= 0) ? ROW(src, const uint8_t, y0)[x0 * cn + k] : cval[k];
                int v1 = (x1 >= 0 && y0 >= 0) ? ROW(src, const uint8_t, y0)[x1 * cn + k] : cval[k];
                int v2 = (x0 >= 0 && y1 >= 0) ? ROW(src, const uint8_t, y1)[x0 * cn + k] : cval[k];
                int v3 = (x1 >= 0 && y1 >= 0) ? ROW(src, const uint8_t, y1)[x1 * cn + k] : cval[k];
                D[k] = sat_u8((v0 * w[0] + v1 * w[1] + v2 * w[2] + v3 * w[3] + (1 << 14)) >> 15);
            }
        }
    }
    return 0;
}

/* cv::convertMaps CV_32FC1 x/y -> CV_16SC2 + CV_16UC1 (what remap does internally per call, done once):
 * nn_interpolation: map1 = saturate_cast<short>(cvRound(x, y)), map2 untouched; else 5 fractional bits. */
int so_convert_maps(const so_mat *xmap, const so_mat *ymap, so_mat *map1, so_mat *map2, int nn_interpolation)
{
    if (xmap->type != SO_32FC1 || ymap->type != SO_32FC1 || map1->type != SO_16SC2) return -1;
    if (map1->rows != xmap->rows || map1->cols != xmap->cols) return -1;
    if (!nn_interpolation && (!map2 || map2->type != SO_16UC1 || map2->rows != xmap->rows || map2->cols != xmap->cols)) return -1;
    for (int y = 0; y < xmap->rows; ++y) {
        const float *mx = ROW(xmap, const float, y), *my = ROW(ymap, const float, y);
        short *m1 = ROW(map1, short, y);
        unsigned short *m2 = nn_interpolation ? NULL : ROW(map2, unsigned short, y);
        for (int x = 0; x < xmap->cols; ++x) {
            if (nn_interpolation) {
                m1[2 * x] = (short)sat_s16(so_cvround(mx[x])); m1[2 * x + 1] = (short)sat_s16(so_cvround(my[x]));
            } else {
                const int ix = so_cvround(mx[x] * 32), iy = so_cvround(my[x] * 32);
                m1[2 * x] = (short)sat_s16(ix >> 5); m1[2 * x + 1] = (short)sat_s16(iy >> 5);
                m2[x] = (unsigned short)((iy & 31) * 32 + (ix & 31));
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ copyMakeBorder
 * blenders.cpp:273 (BORDER_REFLECT on the 16SC3/8UC3 image), :295 (BORDER_CONSTANT 0 on weights) */
int so_copy_make_border(const so_mat *src, so_mat *dst, int top, int bottom, int left, int right, int border)
{
    int esz = so_elem_size1(src->type) * SO_CN(src->type);
    if (dst->type != src->type || dst->rows != src->rows + top + bottom || dst->cols != src->cols + left + right)
        return -1;
    for (int y = 0; y < dst->rows; ++y) {
        int sy = so_border_interpolate(y - top, src->rows, border);
        char *d = ROW(dst, char, y);
        for (int x = 0; x < dst->cols; ++x) {
            int sx = so_border_interpolate(x - left, src->cols, border);
            if (sx < 0 || sy < 0) memset(d + (size_t)x * esz, 0, esz);
            else memcpy(d + (size_t)x * esz, ROW(src, const char, sy) + (size_t)sx * esz, esz);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ pyrDown
 * pyramids.cpp pyrDown_<CastOp>; Appendix A2.  Call sites blenders.cpp:298,448,454,481. */
static int g_float_pyrdown_order = 0;
void so_set_float_pyrdown_order(int mode) { g_float_pyrdown_order = mode; }

int so_pyr_down(const so_mat *src, so_mat *dst)
{
    int depth = SO_DEPTH(src->type), cn = SO_CN(src->type);
    int sw = src->cols, sh = src->rows, dw = dst->cols, dh = dst->rows;
    if (dst->type != src->type || dw != (sw + 1) / 2 || dh != (sh + 1) / 2) return -1;
    const int K[5] = { 1, 4, 6, 4, 1 };

    if (depth == SO_32F) {
        /* horizontal pass into a float row buffer per needed source row, then vertical.
         * scalar order: row = s2*6 + (s1+s3)*4 + s0 + s4 ; dst = (r2*6 + (r1+r3)*4 + r0 + r4) * (1/256) */
        float *rows = (float *)malloc(sizeof(float) * 5 * (size_t)dw * cn);
        for (int y = 0; y < dh; ++y) {
            for (int i = 0; i < 5; ++i) {
                int sy = so_border_interpolate(2 * y + i - 2, sh, SO_BORDER_REFLECT_101);
                const float *s = ROW(src, const float, sy);
                float *r = rows + (size_t)i * dw * cn;
                for (int x = 0; x < dw; ++x)
                    for (int c = 0; c < cn; ++c) {
                        float v[5];
                        for (int j = 0; j < 5; ++j)
                            v[j] = s[so_border_interpolate(2 * x + j - 2, sw, SO_BORDER_REFLECT_101) * cn + c];
                        r[x * cn + c] = v[2] * 6 + (v[1] + v[3]) * 4 + v[0] + v[4];
                    }
            }
            float *d = ROW(dst, float, y);
            const float *r0 = rows, *r1 = rows + (size_t)dw * cn, *r2 = r1 + (size_t)dw * cn,
                        *r3 = r2 + (size_t)dw * cn, *r4 = r3 + (size_t)dw * cn;
            for (int x = 0; x < dw * cn; ++x) {
                if (g_float_pyrdown_order == 0)
                    d[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (1.f / 256);
                else {   /* PyrDownVec_32f (SSE) order: ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4 */
                    float a = r0[x] + r4[x], b = (r1[x] + r3[x]) + r2[x];
                    a = a + (r2[x] + r2[x]);
                    d[x] = (a + b * 4) * (1.f / 256);
                }
            }
        }
        free(rows);
        return 0;
    }
    if (depth != SO_8U && depth != SO_16S) return -1;
    for (int y = 0; y < dh; ++y) {
        int sy[5];
        for (int i = 0; i < 5; ++i) sy[i] = so_border_interpolate(2 * y + i - 2, sh, SO_BORDER_REFLECT_101);
        for (int x = 0; x < dw; ++x) {
            int sx[5];
            for (int j = 0; j < 5; ++j) sx[j] = so_border_interpolate(2 * x + j - 2, sw, SO_BORDER_REFLECT_101);
            for (int c = 0; c < cn; ++c) {
                int acc = 0;
                for (int i = 0; i < 5; ++i)
                    for (int j = 0; j < 5; ++j) {
                        int p = depth == SO_8U ? ROW(src, const uint8_t, sy[i])[sx[j] * cn + c]
                                               : ROW(src, const short, sy[i])[sx[j] * cn + c];
                        acc += K[i] * K[j] * p;
                    }
                int v = (acc + 128) >> 8;            /* FixedPtCast<int,T,8>; >> is arithmetic */
                if (depth == SO_8U) ROW(dst, uint8_t, y)[x * cn + c] = sat_u8(v);
                else ROW(dst, short, y)[x * cn + c] = sat_s16(v);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ pyrUp
 * pyramids.cpp pyrUp_<CastOp>, exact-2x case; Appendix A3.  Call sites blenders.cpp:463,472,485,527.
 * Width-1 sources are undefined behaviour in 2.4.11 (reads src[x+cn]); restated here with the
 * replicate neighbour, which is what cv2 4.x does. */
int so_pyr_up(const so_mat *src, so_mat *dst)
{
    int depth = SO_DEPTH(src->type), cn = SO_CN(src->type);
    int sw = src->cols, sh = src->rows;
    if (dst->type != src->type || dst->cols != sw * 2 || dst->rows != sh * 2) return -1;
    if (depth != SO_8U && depth != SO_16S) return -1;
    int n = sw * 2 * cn;
    int *rowbuf = (int *)malloc(sizeof(int) * 3 * (size_t)n);
    for (int y = 0; y < sh; ++y) {
        int srow[3] = { y == 0 ? (sh > 1 ? 1 : 0) : y - 1, y, y + 1 < sh ? y + 1 : sh - 1 };
        for (int i = 0; i < 3; ++i) {
            int *r = rowbuf + (size_t)i * n;
            for (int x = 0; x < sw; ++x) {
                int xl = x == 0 ? (sw > 1 ? 1 : 0) : x - 1;     /* reflect-101 on the left */
                int xr = x + 1 < sw ? x + 1 : sw - 1;           /* replicate on the right */
                for (int c = 0; c < cn; ++c) {
                    int a, b, d;
                    if (depth == SO_8U) {
                        const uint8_t *s = ROW(src, const uint8_t, srow[i]);
                        a = s[xl * cn + c]; b = s[x * cn + c]; d = s[xr * cn + c];
                    } else {
                        const short *s = ROW(src, const short, srow[i]);
                        a = s[xl * cn + c]; b = s[x * cn + c]; d = s[xr * cn + c];
                    }
                    r[(2 * x) * cn + c] = a + b * 6 + d;
                    r[(2 * x + 1) * cn + c] = (b + d) * 4;
                }
            }
        }
        const int *r0 = rowbuf, *r1 = rowbuf + n, *r2 = rowbuf + 2 * (size_t)n;
        for (int x = 0; x < n; ++x) {
            int v0 = (r1[x] * 6 + r0[x] + r2[x] + 32) >> 6;     /* FixedPtCast<int,T,6> */
            int v1 = ((r1[x] + r2[x]) * 4 + 32) >> 6;
            if (depth == SO_8U) {
                ROW(dst, uint8_t, 2 * y)[x] = sat_u8(v0); ROW(dst, uint8_t, 2 * y + 1)[x] = sat_u8(v1);
            } else {
                ROW(dst, short, 2 * y)[x] = sat_s16(v0); ROW(dst, short, 2 * y + 1)[x] = sat_s16(v1);
            }
        }
    }
    free(rowbuf);
    return 0;
}

/* ------------------------------------------------------------------ arithmetic / conversion (A4) */
int so_add_16s(const so_mat *a, const so_mat *b, so_mat *dst)
{
    int n = a->cols * SO_CN(a->type);
    for (int y = 0; y < a->rows; ++y) {
        const short *pa = ROW(a, const short, y), *pb = ROW(b, const short, y);
        short *d = ROW(dst, short, y);
        for (int x = 0; x < n; ++x) d[x] = sat_s16((int)pa[x] + pb[x]);
    }
    return 0;
}
int so_subtract_16s(const so_mat *a, const so_mat *b, so_mat *dst)
{
    int n = a->cols * SO_CN(a->type);
    for (int y = 0; y < a->rows; ++y) {
        const short *pa = ROW(a, const short, y), *pb = ROW(b, const short, y);
        short *d = ROW(dst, short, y);
        for (int x = 0; x < n; ++x) d[x] = sat_s16((int)pa[x] - pb[x]);
    }
    return 0;
}
int so_subtract_8u_to_16s(const so_mat *a, const so_mat *b, so_mat *dst)
{
    int n = a->cols * SO_CN(a->type);
    for (int y = 0; y < a->rows; ++y) {
        const uint8_t *pa = ROW(a, const uint8_t, y), *pb = ROW(b, const uint8_t, y);
        short *d = ROW(dst, short, y);
        for (int x = 0; x < n; ++x) d[x] = (short)((int)pa[x] - pb[x]);
    }
    return 0;
}
int so_convert_8u_16s(const so_mat *src, so_mat *dst)   /* stitcher.cpp:285, APP64:755 */
{
    int n = src->cols * SO_CN(src->type);
    for (int y = 0; y < src->rows; ++y) {
        const uint8_t *s = ROW(src, const uint8_t, y);
        short *d = ROW(dst, short, y);
        for (int x = 0; x < n; ++x) d[x] = s[x];
    }
    return 0;
}
int so_convert_16s_8u(const so_mat *src, so_mat *dst)   /* stitcher.cpp:313 result.convertTo(CV_8U) */
{
    int n = src->cols * SO_CN(src->type);
    for (int y = 0; y < src->rows; ++y) {
        const short *s = ROW(src, const short, y);
        uint8_t *d = ROW(dst, uint8_t, y);
        for (int x = 0; x < n; ++x) d[x] = sat_u8(s[x]);
    }
    return 0;
}
/* Mat::operator*=(double) on 8U = convertTo(-1, alpha): cvtScale_<uchar,uchar,float>:
 * saturate_cast<uchar>(src*(float)alpha + 0.f) with cvRound; exposure_compensate.cpp:152 */
int so_scale_8u(so_mat *img, double gain)
{
    float g = (float)gain;
    int n = img->cols * SO_CN(img->type);
    if (SO_DEPTH(img->type) != SO_8U) return -1;
    for (int y = 0; y < img->rows; ++y) {
        uint8_t *p = ROW(img, uint8_t, y);
        for (int x = 0; x < n; ++x) p[x] = sat_u8(so_cvround((float)p[x] * g));
    }
    return 0;
}

/* ------------------------------------------------------------------ distanceTransform(CV_DIST_L1, 3)
 * distransform.cpp distanceTransform_3x3 (16.16 fixed point, 1-px INIT_DIST0 border);
 * call site blenders.cpp:430. */
int so_distance_l1_3x3(const so_mat *mask, so_mat *dist)
{
    const int HV = 1 << 16, DIAG = 2 << 16, INIT0 = INT_MAX >> 2;
    const float scale = 1.f / (1 << 16);
    int w = mask->cols, h = mask->rows, step = w + 2;
    if (mask->type != SO_8UC1 || dist->type != SO_32FC1 || dist->rows != h || dist->cols != w) return -1;
    int *tmp = (int *)malloc(sizeof(int) * (size_t)step * (h + 2));
    for (size_t i = 0; i < (size_t)step * (h + 2); ++i) tmp[i] = INIT0;
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = ROW(mask, const uint8_t, y);
        int *t = tmp + (size_t)(y + 1) * step + 1;
        for (int x = 0; x < w; ++x) {
            if (!s[x]) t[x] = 0;
            else {
                int t0 = t[x - step - 1] + DIAG, v;
                v = t[x - step] + HV;     if (t0 > v) t0 = v;
                v = t[x - step + 1] + DIAG; if (t0 > v) t0 = v;
                v = t[x - 1] + HV;        if (t0 > v) t0 = v;
                t[x] = t0;
            }
        }
    }
    for (int y = h - 1; y >= 0; --y) {
        float *d = ROW(dist, float, y);
        int *t = tmp + (size_t)(y + 1) * step + 1;
        for (int x = w - 1; x >= 0; --x) {
            int t0 = t[x];
            if (t0 > HV) {
                int v;
                v = t[x + step + 1] + DIAG; if (t0 > v) t0 = v;
                v = t[x + step] + HV;       if (t0 > v) t0 = v;
                v = t[x + step - 1] + DIAG; if (t0 > v) t0 = v;
                v = t[x + 1] + HV;          if (t0 > v) t0 = v;
                t[x] = t0;
            }
            if (t0 > INIT0) t0 = INIT0;
            d[x] = (float)(t0 * scale);
        }
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------ resize INTER_LINEAR 32FC1
 * imgwarp.cpp resize (HResizeLinear / VResizeLinear, float); call site exposure_compensate.cpp:233,
 * APP64:316. */
int so_resize_linear_32f(const so_mat *src, so_mat *dst)
{
    int sw = src->cols, sh = src->rows, dw = dst->cols, dh = dst->rows;
    if (src->type != SO_32FC1 || dst->type != SO_32FC1) return -1;
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int *xofs = (int *)malloc(sizeof(int) * dw);
    float *alpha = (float *)malloc(sizeof(float) * 2 * dw);
    float *r0 = (float *)malloc(sizeof(float) * dw), *r1 = (float *)malloc(sizeof(float) * dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx; alpha[2 * dx] = 1.f - fx; alpha[2 * dx + 1] = fx;
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        float b0 = 1.f - fy, b1 = fy;       /* vertical: rows are clipped, beta is not zeroed */
        int sy0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
        int sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
        const float *S0 = ROW(src, const float, sy0), *S1 = ROW(src, const float, sy1);
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            r0[dx] = S0[sx] * alpha[2 * dx] + S0[sx1] * alpha[2 * dx + 1];
            r1[dx] = S1[sx] * alpha[2 * dx] + S1[sx1] * alpha[2 * dx + 1];
        }
        float *d = ROW(dst, float, dy);
        for (int dx = 0; dx < dw; ++dx) d[dx] = r0[dx] * b0 + r1[dx] * b1;
    }
    free(xofs); free(alpha); free(r0); free(r1);
    return 0;
}
