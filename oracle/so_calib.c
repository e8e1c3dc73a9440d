/*
 * so_calib.c — CPU ORACLE (test infrastructure, NOT product code): the once-per-calibration steps next to the
 * per-frame path.
 *   GainCompensator::feed            LIB/src/exposure_compensate.cpp:76-147
 *   BlocksGainCompensator::feed      LIB/src/exposure_compensate.cpp:165-222
 *   seam-mask refinement             LIB/src/stitcher.cpp:291-294  (dilate, resize, &)
 * plus the OpenCV 2.4.11 primitives those lines call (sources not in /root/reference): cv::solve (DECOMP_LU),
 * cv::sepFilter2D with the symmetric 3-tap kernel, cv::dilate with the default 3x3 element, cv::resize
 * INTER_LINEAR on 8UC1.  Pinned against cv2 4.13 by tests/golden/make_golden.py (calib.npz) and
 * tests/test_oracle_vs_cv2.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "stitch_oracle.h"

#define ROW(m, T, y) ((T *)((char *)(m)->data + (size_t)(y) * (m)->step))

/* util.cpp:103-116 overlapRoi */
static int overlap_roi(const int tl1[2], const int tl2[2], int w1, int h1, int w2, int h2, int roi[4])
{
    int x_tl = tl1[0] > tl2[0] ? tl1[0] : tl2[0], y_tl = tl1[1] > tl2[1] ? tl1[1] : tl2[1];
    int x_br = tl1[0] + w1 < tl2[0] + w2 ? tl1[0] + w1 : tl2[0] + w2;
    int y_br = tl1[1] + h1 < tl2[1] + h2 ? tl1[1] + h1 : tl2[1] + h2;
    if (x_tl < x_br && y_tl < y_br) { roi[0] = x_tl; roi[1] = y_tl; roi[2] = x_br - x_tl; roi[3] = y_br - y_tl; return 1; }
    return 0;
}

/* cv::solve(A, b, x) with DECOMP_LU on doubles: Gaussian elimination with partial pivoting (OpenCV's LUImpl:
 * pivot = largest |a| in the column, rows swapped, eliminate below, then back substitution).  Returns 0 when a
 * pivot is smaller than DBL_EPSILON * 100 (cv::solve returns false and leaves x untouched — here x = 0). */
int so_solve_lu(int n, double *A, double *b)
{
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < 2.220446049250313e-16 * 100) { memset(b, 0, sizeof(double) * n); return 0; }
        if (k != i) {
            for (int j = i; j < n; ++j) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; ++j) {
            double alpha = A[j * n + i] * d;
            for (k = i + 1; k < n; ++k) A[j * n + k] += alpha * A[i * n + k];
            b[j] += alpha * b[i];
        }
        A[i * n + i] = -d;
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * b[k];
        b[i] = s * A[i * n + i];
    }
    return 1;
}

/* exposure_compensate.cpp:93-126: overlap statistics of every image pair, in the reference's loop order
 * (row-major over the overlap, sequential double accumulation).  N: n x n int, I: n x n double. */
void so_gain_overlap_stats(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals,
                           int *N, double *I)
{
    memset(N, 0, sizeof(int) * n * n);
    memset(I, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            int roi[4];
            if (!overlap_roi(corners_xy + 2 * i, corners_xy + 2 * j, images[i].cols, images[i].rows, images[j].cols, images[j].rows, roi))
                continue;
            int x1 = roi[0] - corners_xy[2 * i], y1 = roi[1] - corners_xy[2 * i + 1];
            int x2 = roi[0] - corners_xy[2 * j], y2 = roi[1] - corners_xy[2 * j + 1];
            int cnt = 0;
            double Isum1 = 0, Isum2 = 0;
            for (int y = 0; y < roi[3]; ++y) {
                const unsigned char *m1 = ROW(&masks[i], const unsigned char, y1 + y) + x1, *m2 = ROW(&masks[j], const unsigned char, y2 + y) + x2;
                const unsigned char *r1 = ROW(&images[i], const unsigned char, y1 + y) + 3 * x1, *r2 = ROW(&images[j], const unsigned char, y2 + y) + 3 * x2;
                for (int x = 0; x < roi[2]; ++x)
                    if (m1[x] == mask_vals[i] && m2[x] == mask_vals[j]) {
                        ++cnt;
                        Isum1 += sqrt((double)(r1[3 * x] * r1[3 * x] + r1[3 * x + 1] * r1[3 * x + 1] + r1[3 * x + 2] * r1[3 * x + 2]));
                        Isum2 += sqrt((double)(r2[3 * x] * r2[3 * x] + r2[3 * x + 1] * r2[3 * x + 1] + r2[3 * x + 2] * r2[3 * x + 2]));
                    }
            }
            N[i * n + j] = N[j * n + i] = cnt > 1 ? cnt : 1;
            I[i * n + j] = Isum1 / N[i * n + j];
            I[j * n + i] = Isum2 / N[i * n + j];
        }
}

/* exposure_compensate.cpp:128-144: the normal equations of the gain model and their solution */
int so_gain_solve(int n, const int *N, const double *I, double *gains)
{
    const double alpha = 0.01, beta = 100;
    double *A = (double *)calloc((size_t)n * n, sizeof(double));
    memset(gains, 0, sizeof(double) * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            gains[i] += beta * N[i * n + j];
            A[i * n + i] += beta * N[i * n + j];
            if (j == i) continue;
            A[i * n + i] += 2 * alpha * I[i * n + j] * I[i * n + j] * N[i * n + j];
            A[i * n + j] -= 2 * alpha * I[i * n + j] * I[j * n + i] * N[i * n + j];
        }
    int ok = so_solve_lu(n, A, gains);
    free(A);
    return ok ? 0 : -1;
}

/* GainCompensator::feed (exposure_compensate.cpp:76-147).  images 8UC3, masks 8UC1. */
int so_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals, double *gains)
{
    for (int i = 0; i < n; ++i)
        if (images[i].type != SO_8UC3 || masks[i].type != SO_8UC1 || images[i].rows != masks[i].rows || images[i].cols != masks[i].cols) return -1;
    int *N = (int *)malloc(sizeof(int) * n * n);
    double *I = (double *)malloc(sizeof(double) * n * n);
    so_gain_overlap_stats(n, corners_xy, images, masks, mask_vals, N, I);
    int rc = so_gain_solve(n, N, I, gains);
    free(N); free(I);
    return rc;
}

/* cv::sepFilter2D(src, dst, CV_32F, ker, ker) with the symmetric 3-tap kernel {k1, k0, k1}, BORDER_REFLECT_101:
 * OpenCV's SymmRowSmallFilter / SymmColumnSmallFilter evaluate S[i]*k0 + (S[i-1] + S[i+1])*k1 (pinned against cv2). */
int so_sep_filter3_f32(const so_mat *src, so_mat *dst, float k0, float k1)
{
    int w = src->cols, h = src->rows;
    if (src->type != SO_32FC1 || dst->type != SO_32FC1 || dst->rows != h || dst->cols != w) return -1;
    float *tmp = (float *)malloc(sizeof(float) * w * h);
    for (int y = 0; y < h; ++y) {
        const float *S = ROW(src, const float, y);
        for (int x = 0; x < w; ++x) {
            int l = so_border_interpolate(x - 1, w, SO_BORDER_REFLECT_101), r = so_border_interpolate(x + 1, w, SO_BORDER_REFLECT_101);
            tmp[y * w + x] = S[x] * k0 + (S[l] + S[r]) * k1;
        }
    }
    for (int y = 0; y < h; ++y) {
        int u = so_border_interpolate(y - 1, h, SO_BORDER_REFLECT_101), d = so_border_interpolate(y + 1, h, SO_BORDER_REFLECT_101);
        float *D = ROW(dst, float, y);
        for (int x = 0; x < w; ++x) D[x] = tmp[y * w + x] * k0 + (tmp[u * w + x] + tmp[d * w + x]) * k1;
    }
    free(tmp);
    return 0;
}

/* BlocksGainCompensator::feed (exposure_compensate.cpp:165-222): one GainCompensator over all blocks, then the
 * per-image block gain maps smoothed twice.  gain_maps[i] must be CV_32FC1 of the block grid size
 * ((cols + bl_width - 1) / bl_width, (rows + bl_height - 1) / bl_height). */
int so_blocks_gain_feed(int n, const int *corners_xy, const so_mat *images, const so_mat *masks, const unsigned char *mask_vals,
                        int bl_width_, int bl_height_, so_mat *gain_maps)
{
    int total = 0;
    for (int i = 0; i < n; ++i) total += ((images[i].cols + bl_width_ - 1) / bl_width_) * ((images[i].rows + bl_height_ - 1) / bl_height_);
    int *bc = (int *)malloc(sizeof(int) * 2 * total);
    so_mat *bi = (so_mat *)malloc(sizeof(so_mat) * total), *bm = (so_mat *)malloc(sizeof(so_mat) * total);
    unsigned char *bv = (unsigned char *)malloc(total);
    double *gains = (double *)malloc(sizeof(double) * total);
    int b = 0;
    for (int i = 0; i < n; ++i) {
        int bx_n = (images[i].cols + bl_width_ - 1) / bl_width_, by_n = (images[i].rows + bl_height_ - 1) / bl_height_;
        int bl_width = (images[i].cols + bx_n - 1) / bx_n, bl_height = (images[i].rows + by_n - 1) / by_n;
        if (gain_maps[i].type != SO_32FC1 || gain_maps[i].cols != bx_n || gain_maps[i].rows != by_n) return -1;
        for (int by = 0; by < by_n; ++by)
            for (int bx = 0; bx < bx_n; ++bx, ++b) {
                int x0 = bx * bl_width, y0 = by * bl_height;
                int x1 = x0 + bl_width < images[i].cols ? x0 + bl_width : images[i].cols;
                int y1 = y0 + bl_height < images[i].rows ? y0 + bl_height : images[i].rows;
                bc[2 * b] = corners_xy[2 * i] + x0; bc[2 * b + 1] = corners_xy[2 * i + 1] + y0;
                bi[b] = images[i]; bi[b].data = ROW(&images[i], unsigned char, y0) + 3 * x0; bi[b].cols = x1 - x0; bi[b].rows = y1 - y0;
                bm[b] = masks[i]; bm[b].data = ROW(&masks[i], unsigned char, y0) + x0; bm[b].cols = x1 - x0; bm[b].rows = y1 - y0;
                bv[b] = mask_vals[i];
            }
    }
    int rc = so_gain_feed(total, bc, bi, bm, bv, gains);
    b = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        for (int y = 0; y < gain_maps[i].rows; ++y)
            for (int x = 0; x < gain_maps[i].cols; ++x, ++b) ROW(&gain_maps[i], float, y)[x] = (float)gains[b];
        so_sep_filter3_f32(&gain_maps[i], &gain_maps[i], 0.5f, 0.25f);
        so_sep_filter3_f32(&gain_maps[i], &gain_maps[i], 0.5f, 0.25f);
    }
    free(bc); free(bi); free(bm); free(bv); free(gains);
    return rc;
}

/* cv::dilate(src, dst, Mat()): 3x3 rectangle, anchor at the centre, one iteration, the default border (pixels outside
 * the image do not take part in the maximum).  8UC1.  Call site stitcher.cpp:291. */
int so_dilate3x3_8u(const so_mat *src, so_mat *dst)
{
    int w = src->cols, h = src->rows;
    if (src->type != SO_8UC1 || dst->type != SO_8UC1 || dst->rows != h || dst->cols != w || src->data == dst->data) return -1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned char m = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                    unsigned char v = ROW(src, const unsigned char, yy)[xx];
                    if (v > m) m = v;
                }
            ROW(dst, unsigned char, y)[x] = m;
        }
    return 0;
}

/* cv::resize(src, dst, dst.size()) with INTER_LINEAR on 8UC1 (imgwarp.cpp; call site stitcher.cpp:292):
 * equal sizes copy; an exact 2x2 decimation is rerouted to INTER_AREA ((a+b+c+d+2)>>2); otherwise the fixed-point
 * path: coefficients cvRound(f * 2048) as short, horizontal pass in int, vertical pass
 * (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2. */
int so_resize_linear_8u(const so_mat *src, so_mat *dst)
{
    int sw = src->cols, sh = src->rows, dw = dst->cols, dh = dst->rows;
    if (src->type != SO_8UC1 || dst->type != SO_8UC1) return -1;
    if (sw == dw && sh == dh) {
        for (int y = 0; y < sh; ++y) memcpy(ROW(dst, unsigned char, y), ROW(src, const unsigned char, y), sw);
        return 0;
    }
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int iscale_x = (int)nearbyint(scale_x), iscale_y = (int)nearbyint(scale_y);      /* saturate_cast<int>(double) = cvRound */
    int is_area_fast = fabs(scale_x - iscale_x) < 2.220446049250313e-16 && fabs(scale_y - iscale_y) < 2.220446049250313e-16;
    if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
        for (int y = 0; y < dh; ++y) {
            const unsigned char *S0 = ROW(src, const unsigned char, 2 * y), *S1 = ROW(src, const unsigned char, 2 * y + 1);
            unsigned char *D = ROW(dst, unsigned char, y);
            for (int x = 0; x < dw; ++x) D[x] = (unsigned char)((S0[2 * x] + S0[2 * x + 1] + S1[2 * x] + S1[2 * x + 1] + 2) >> 2);
        }
        return 0;
    }
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *ialpha = (short *)malloc(sizeof(short) * 2 * dw);
    int *r0 = (int *)malloc(sizeof(int) * dw), *r1 = (int *)malloc(sizeof(int) * dw);
    int xmax = dw;
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx + 1 >= sw) { if (dx < xmax) xmax = dx; fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)so_cvround((1.f - fx) * 2048); ialpha[2 * dx + 1] = (short)so_cvround(fx * 2048);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = (short)so_cvround((1.f - fy) * 2048), b1 = (short)so_cvround(fy * 2048);
        int sy0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
        int sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
        const unsigned char *S0 = ROW(src, const unsigned char, sy0), *S1 = ROW(src, const unsigned char, sy1);
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            if (dx < xmax) {
                r0[dx] = S0[sx] * ialpha[2 * dx] + S0[sx + 1] * ialpha[2 * dx + 1];
                r1[dx] = S1[sx] * ialpha[2 * dx] + S1[sx + 1] * ialpha[2 * dx + 1];
            } else {
                r0[dx] = S0[sx] * 2048; r1[dx] = S1[sx] * 2048;
            }
        }
        unsigned char *D = ROW(dst, unsigned char, dy);
        for (int dx = 0; dx < dw; ++dx) D[dx] = (unsigned char)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(ialpha); free(r0); free(r1);
    return 0;
}

/* stitcher.cpp:291-294: dilate(seam mask) -> resize to the warped mask's size -> & warped mask */
int so_refine_seam_mask(const so_mat *seam_mask, const so_mat *mask_warped, so_mat *out)
{
    if (out->rows != mask_warped->rows || out->cols != mask_warped->cols || out->type != SO_8UC1) return -1;
    so_mat dil = *seam_mask, rs = *out;
    dil.step = (size_t)seam_mask->cols; dil.data = malloc((size_t)seam_mask->rows * seam_mask->cols);
    rs.step = (size_t)out->cols; rs.data = malloc((size_t)out->rows * out->cols);
    int rc = so_dilate3x3_8u(seam_mask, &dil);
    if (rc == 0) rc = so_resize_linear_8u(&dil, &rs);
    for (int y = 0; rc == 0 && y < out->rows; ++y) {
        const unsigned char *a = ROW(&rs, const unsigned char, y), *b = ROW(mask_warped, const unsigned char, y);
        unsigned char *d = ROW(out, unsigned char, y);
        for (int x = 0; x < out->cols; ++x) d[x] = a[x] & b[x];
    }
    free(dil.data); free(rs.data);
    return rc;
}

/* FeatherBlender::createWeightMaps (blenders.cpp:158-186): per-image feather weights normalised by their sum over the
 * result ROI, so that the final image is a plain weighting of the sources.  Note that `tmp` in the reference is a VIEW
 * of weights_sum: setTo(1, tmp < eps) writes the 1 back into the shared sum before the next image is divided.
 * cv::divide on CV_32F: src2 != 0 ? src1 / src2 : 0 (computed in double and rounded once = the IEEE float quotient). */
int so_feather_create_weight_maps(int n, const so_mat *masks, const int *corners_xy, float sharpness, so_mat *weight_maps, int roi_xywh[4])
{
    int *sizes = (int *)malloc(sizeof(int) * 2 * n);
    for (int i = 0; i < n; ++i) {
        if (so_create_weight_map(&masks[i], sharpness, &weight_maps[i]) != 0) { free(sizes); return -1; }
        sizes[2 * i] = masks[i].cols; sizes[2 * i + 1] = masks[i].rows;
    }
    so_result_roi(corners_xy, sizes, n, roi_xywh);
    free(sizes);
    int W = roi_xywh[2], H = roi_xywh[3];
    float *sum = (float *)calloc((size_t)W * H, sizeof(float));
    for (int i = 0; i < n; ++i) {
        int ox = corners_xy[2 * i] - roi_xywh[0], oy = corners_xy[2 * i + 1] - roi_xywh[1];
        for (int y = 0; y < weight_maps[i].rows; ++y) {
            const float *w = ROW(&weight_maps[i], const float, y);
            for (int x = 0; x < weight_maps[i].cols; ++x) sum[(size_t)(oy + y) * W + ox + x] += w[x];
        }
    }
    for (int i = 0; i < n; ++i) {
        int ox = corners_xy[2 * i] - roi_xywh[0], oy = corners_xy[2 * i + 1] - roi_xywh[1];
        for (int y = 0; y < weight_maps[i].rows; ++y) {
            float *w = ROW(&weight_maps[i], float, y);
            for (int x = 0; x < weight_maps[i].cols; ++x) {
                float *t = &sum[(size_t)(oy + y) * W + ox + x];
                if (*t < 1.1920928955078125e-07f) *t = 1.f;
                w[x] = *t != 0 ? w[x] / *t : 0.f;
            }
        }
    }
    free(sum);
    return 0;
}
