/*
 * so_stitch.c — CPU ORACLE (test infrastructure, NOT product code; see stitch_oracle.h).
 *
 * Line-by-line restatement of the reference's own compositing code
 * (LIB = /root/reference/stitching/OpenCV2.4.11-Stitching-64bit/OpenCV2.4.11-Stitching):
 *   LIB/src/warpers.cpp:50-78,139-212   LIB/include/opencv2/stitching/detail/warpers_inl.hpp:52-300
 *   LIB/src/exposure_compensate.cpp:150-153,225-246
 *   LIB/src/util.cpp:127-140            LIB/src/blenders.cpp:65-186,189-377,383-432,435-489,520-530
 * Build: gcc -O2 -ffp-contract=off.
 */
#include "stitch_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SO_DEPTH(t) ((t) & 7)
#define SO_CN(t) ((((t) >> 3) & 63) + 1)
#define ROW(m, T, y) ((T *)((char *)(m)->data + (size_t)(y) * (m)->step))
#define SO_PI_F ((float)3.1415926535897932384626433832795)   /* static_cast<float>(CV_PI) */

static int mat_alloc(so_mat *m, int rows, int cols, int type)
{
    int d = SO_DEPTH(type), esz = (d == SO_8U ? 1 : d == SO_16S ? 2 : 4) * SO_CN(type);
    m->rows = rows; m->cols = cols; m->type = type; m->step = (size_t)cols * esz;
    m->data = calloc((size_t)rows * cols + 1, esz);          /* create + setTo(0) */
    return m->data ? 0 : -1;
}
static void mat_free(so_mat *m) { free(m->data); memset(m, 0, sizeof *m); }

/* ------------------------------------------------------------------ ProjectorBase::setCameraParams
 * warpers.cpp:50-78.  Mat ops restated from OpenCV 2.4.11 core: 3x3 CV_32F cv::invert (double
 * cofactors, float result) and the len==3 float fast path of cv::gemm (float accumulate). */
static void inv3x3_f(const float S[9], float D[9])
{
#define Sf(r, c) ((double)S[(r) * 3 + (c)])
    double d = Sf(0,0) * (Sf(1,1) * Sf(2,2) - Sf(1,2) * Sf(2,1)) - Sf(0,1) * (Sf(1,0) * Sf(2,2) - Sf(1,2) * Sf(2,0))
             + Sf(0,2) * (Sf(1,0) * Sf(2,1) - Sf(1,1) * Sf(2,0));
    if (d == 0.) { memset(D, 0, 9 * sizeof(float)); return; }   /* invert() zero-fills a singular result */
    d = 1. / d;
    D[0] = (float)((Sf(1,1) * Sf(2,2) - Sf(1,2) * Sf(2,1)) * d);
    D[1] = (float)((Sf(0,2) * Sf(2,1) - Sf(0,1) * Sf(2,2)) * d);
    D[2] = (float)((Sf(0,1) * Sf(1,2) - Sf(0,2) * Sf(1,1)) * d);
    D[3] = (float)((Sf(1,2) * Sf(2,0) - Sf(1,0) * Sf(2,2)) * d);
    D[4] = (float)((Sf(0,0) * Sf(2,2) - Sf(0,2) * Sf(2,0)) * d);
    D[5] = (float)((Sf(0,2) * Sf(1,0) - Sf(0,0) * Sf(1,2)) * d);
    D[6] = (float)((Sf(1,0) * Sf(2,1) - Sf(1,1) * Sf(2,0)) * d);
    D[7] = (float)((Sf(0,1) * Sf(2,0) - Sf(0,0) * Sf(2,1)) * d);
    D[8] = (float)((Sf(0,0) * Sf(1,1) - Sf(0,1) * Sf(1,0)) * d);
#undef Sf
}
static void mul3x3_f(const float A[9], const float B[9], float D[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            D[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

/* exported for oracle/ref_shim: the Mat::inv() / Mat * Mat primitives behind the reference's setCameraParams */
void so_inv3x3_f32(const float S[9], float D[9]) { inv3x3_f(S, D); }
void so_mul3x3_f32(const float A[9], const float B[9], float D[9]) { mul3x3_f(A, B, D); }

void so_projector_set(so_projector *p, int kind, float scale, const float K[9], const float R[9], const float T[3])
{
    float kinv[9];
    p->kind = kind; p->scale = scale;
    memcpy(p->k, K, sizeof p->k);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) p->rinv[i * 3 + j] = R[j * 3 + i];   /* R.t() */
    inv3x3_f(K, kinv);
    mul3x3_f(R, kinv, p->r_kinv);          /* R * K.inv() */
    mul3x3_f(K, p->rinv, p->k_rinv);       /* K * Rinv */
    p->t[0] = T ? T[0] : 0.f; p->t[1] = T ? T[1] : 0.f; p->t[2] = T ? T[2] : 0.f;
}

/* warpers_inl.hpp:206-300 (Plane, Spherical, Cylindrical projectors) */
void so_map_forward(const so_projector *p, float x, float y, float *u, float *v)
{
    const float *m = p->r_kinv;
    float x_ = m[0] * x + m[1] * y + m[2];
    float y_ = m[3] * x + m[4] * y + m[5];
    float z_ = m[6] * x + m[7] * y + m[8];
    if (p->kind == SO_WARP_PLANE) {
        x_ = p->t[0] + x_ / z_ * (1 - p->t[2]);
        y_ = p->t[1] + y_ / z_ * (1 - p->t[2]);
        *u = p->scale * x_; *v = p->scale * y_;
    } else if (p->kind == SO_WARP_SPHERICAL) {
        *u = p->scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = p->scale * (SO_PI_F - acosf(w == w ? w : 0));
    } else {
        *u = p->scale * atan2f(x_, z_);
        *v = p->scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

void so_map_backward(const so_projector *p, float u, float v, float *x, float *y)
{
    const float *m = p->k_rinv;
    float x_, y_, z_, z;
    if (p->kind == SO_WARP_PLANE) {
        u = u / p->scale - p->t[0];
        v = v / p->scale - p->t[1];
        *x = m[0] * u + m[1] * v + m[2] * (1 - p->t[2]);
        *y = m[3] * u + m[4] * v + m[5] * (1 - p->t[2]);
        z  = m[6] * u + m[7] * v + m[8] * (1 - p->t[2]);
        *x /= z; *y /= z;
        return;
    }
    u /= p->scale; v /= p->scale;
    if (p->kind == SO_WARP_SPHERICAL) {
        float sinv = so_sinf(SO_PI_F - v);
        x_ = sinv * so_sinf(u);
        y_ = so_cosf(SO_PI_F - v);
        z_ = sinv * so_cosf(u);
    } else {
        x_ = so_sinf(u); y_ = v; z_ = so_cosf(u);
    }
    *x = m[0] * x_ + m[1] * y_ + m[2] * z_;
    *y = m[3] * x_ + m[4] * y_ + m[5] * z_;
    z  = m[6] * x_ + m[7] * y_ + m[8] * z_;
    if (z > 0) { *x /= z; *y /= z; }
    else *x = *y = -1;
}

/* std::min/std::max keep the first argument when the comparison is false (NaN-safe) */
static inline void upd4(float u, float v, float *tl_u, float *tl_v, float *br_u, float *br_v)
{
    if (u < *tl_u) *tl_u = u;
    if (v < *tl_v) *tl_v = v;
    if (u > *br_u) *br_u = u;
    if (v > *br_v) *br_v = v;
}
#define UPD(u, v) upd4((u), (v), &tl_u, &tl_v, &br_u, &br_v)

/* PlaneWarper::detectResultRoi warpers.cpp:139-168; RotationWarperBase::detectResultRoiByBorder
 * warpers_inl.hpp:169-203 (cylindrical uses it through CylindricalWarper, warpers.hpp:360-363);
 * SphericalWarper::detectResultRoi warpers.cpp:171-212 */
void so_detect_result_roi(const so_projector *p, int src_w, int src_h, int tl[2], int br[2])
{
    float tl_u = 3.402823466e+38f, tl_v = 3.402823466e+38f, br_u = -3.402823466e+38f, br_v = -3.402823466e+38f;
    float u, v;
    if (p->kind == SO_WARP_PLANE) {
        so_map_forward(p, 0, 0, &u, &v); UPD(u, v);
        so_map_forward(p, 0, (float)(src_h - 1), &u, &v); UPD(u, v);
        so_map_forward(p, (float)(src_w - 1), 0, &u, &v); UPD(u, v);
        so_map_forward(p, (float)(src_w - 1), (float)(src_h - 1), &u, &v); UPD(u, v);
    } else {
        for (float x = 0; x < src_w; ++x) {
            so_map_forward(p, x, 0, &u, &v); UPD(u, v);
            so_map_forward(p, x, (float)(src_h - 1), &u, &v); UPD(u, v);
        }
        for (int y = 0; y < src_h; ++y) {
            so_map_forward(p, 0, (float)y, &u, &v); UPD(u, v);
            so_map_forward(p, (float)(src_w - 1), (float)y, &u, &v); UPD(u, v);
        }
    }
    tl[0] = (int)tl_u; tl[1] = (int)tl_v; br[0] = (int)br_u; br[1] = (int)br_v;
    if (p->kind != SO_WARP_SPHERICAL) return;

    tl_u = (float)tl[0]; tl_v = (float)tl[1]; br_u = (float)br[0]; br_v = (float)br[1];
    float x = p->rinv[1], y = p->rinv[4], z = p->rinv[7];
    if (y > 0.f) {
        float x_ = (p->k[0] * x + p->k[1] * y) / z + p->k[2];
        float y_ = p->k[4] * y / z + p->k[5];
        if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
            float pv = (float)(3.1415926535897932384626433832795 * p->scale);   /* CV_PI * scale in double */
            UPD(0.f, pv);
        }
    }
    x = p->rinv[1]; y = -p->rinv[4]; z = p->rinv[7];
    if (y > 0.f) {
        float x_ = (p->k[0] * x + p->k[1] * y) / z + p->k[2];
        float y_ = p->k[4] * y / z + p->k[5];
        if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) UPD(0.f, 0.f);
    }
    tl[0] = (int)tl_u; tl[1] = (int)tl_v; br[0] = (int)br_u; br[1] = (int)br_v;
}

/* RotationWarperBase::buildMaps warpers_inl.hpp:62-85; PlaneWarper::buildMaps warpers.cpp:91-113 */
void so_build_maps(const so_projector *p, const int tl[2], const int br[2], so_mat *xmap, so_mat *ymap)
{
    for (int v = tl[1]; v <= br[1]; ++v) {
        float *mx = ROW(xmap, float, v - tl[1]), *my = ROW(ymap, float, v - tl[1]);
        for (int u = tl[0]; u <= br[0]; ++u)
            so_map_backward(p, (float)u, (float)v, &mx[u - tl[0]], &my[u - tl[0]]);
    }
}

/* ------------------------------------------------------------------ exposure compensation apply */
int so_gain_apply(so_mat *image, double gain)           /* exposure_compensate.cpp:150-153 */
{
    return so_scale_8u(image, gain);
}
int so_blocks_gain_apply(so_mat *image, const so_mat *gain_map)   /* exposure_compensate.cpp:225-246 */
{
    so_mat gm = *gain_map, tmp;
    int own = 0;
    if (image->type != SO_8UC3) return -1;
    if (gain_map->rows != image->rows || gain_map->cols != image->cols) {
        if (mat_alloc(&tmp, image->rows, image->cols, SO_32FC1)) return -1;
        so_resize_linear_32f(gain_map, &tmp);
        gm = tmp; own = 1;
    }
    for (int y = 0; y < image->rows; ++y) {
        const float *g = ROW(&gm, const float, y);
        uint8_t *r = ROW(image, uint8_t, y);
        for (int x = 0; x < image->cols; ++x)
            for (int c = 0; c < 3; ++c) {
                int v = so_cvround((float)r[x * 3 + c] * g[x]);      /* saturate_cast<uchar>(float) */
                r[x * 3 + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
    }
    if (own) mat_free(&tmp);
    return 0;
}

/* ------------------------------------------------------------------ util.cpp:127-140 resultRoi */
void so_result_roi(const int *c, const int *s, int n, int roi[4])
{
    int tlx = INT_MAX, tly = INT_MAX, brx = INT_MIN, bry = INT_MIN;
    for (int i = 0; i < n; ++i) {
        if (c[2 * i] < tlx) tlx = c[2 * i];
        if (c[2 * i + 1] < tly) tly = c[2 * i + 1];
        if (c[2 * i] + s[2 * i] > brx) brx = c[2 * i] + s[2 * i];
        if (c[2 * i + 1] + s[2 * i + 1] > bry) bry = c[2 * i + 1] + s[2 * i + 1];
    }
    roi[0] = tlx; roi[1] = tly; roi[2] = brx - tlx; roi[3] = bry - tly;
}

/* ------------------------------------------------------------------ blenders.cpp auxiliaries */
static const float WEIGHT_EPS = 1e-5f;

void so_normalize_using_weight_map(const so_mat *weight, so_mat *src)   /* blenders.cpp:383-424 */
{
    for (int y = 0; y < src->rows; ++y) {
        short *row = ROW(src, short, y);
        if (weight->type == SO_32FC1) {
            const float *w = ROW(weight, const float, y);
            for (int x = 0; x < src->cols; ++x)
                for (int c = 0; c < 3; ++c)
                    row[x * 3 + c] = so_trunc_short(row[x * 3 + c] / (w[x] + WEIGHT_EPS));
        } else {
            const short *w = ROW(weight, const short, y);
            for (int x = 0; x < src->cols; ++x) {
                int wi = w[x] + 1;
                for (int c = 0; c < 3; ++c)
                    row[x * 3 + c] = (short)((row[x * 3 + c] << 8) / wi);
            }
        }
    }
}

int so_create_weight_map(const so_mat *mask, float sharpness, so_mat *weight)   /* blenders.cpp:427-432 */
{
    if (so_distance_l1_3x3(mask, weight)) return -1;
    for (int y = 0; y < weight->rows; ++y) {
        float *w = ROW(weight, float, y);
        for (int x = 0; x < weight->cols; ++x) {
            float v = w[x] * sharpness;        /* MatExpr weight*sharpness -> convertTo float scale */
            w[x] = v > 1.f ? 1.f : v;          /* threshold(THRESH_TRUNC, 1) */
        }
    }
    return 0;
}

int so_create_laplace_pyr(const so_mat *img, int num_levels, so_mat *pyr)   /* blenders.cpp:435-489 */
{
    int cn = SO_CN(img->type);
    if (SO_DEPTH(img->type) == SO_8U) {
        if (num_levels == 0) return so_convert_8u_16s(img, &pyr[0]);
        so_mat current = *img, down_next, lvl_up, lvl_down;
        int cur_own = 0;
        mat_alloc(&down_next, (img->rows + 1) / 2, (img->cols + 1) / 2, img->type);
        so_pyr_down(img, &down_next);
        for (int i = 1; i < num_levels; ++i) {
            mat_alloc(&lvl_down, (down_next.rows + 1) / 2, (down_next.cols + 1) / 2, img->type);
            so_pyr_down(&down_next, &lvl_down);
            mat_alloc(&lvl_up, current.rows, current.cols, img->type);
            if (so_pyr_up(&down_next, &lvl_up)) return -1;
            so_subtract_8u_to_16s(&current, &lvl_up, &pyr[i - 1]);
            mat_free(&lvl_up);
            if (cur_own) mat_free(&current);
            current = down_next; cur_own = 1;
            down_next = lvl_down;
        }
        mat_alloc(&lvl_up, current.rows, current.cols, img->type);
        if (so_pyr_up(&down_next, &lvl_up)) return -1;
        so_subtract_8u_to_16s(&current, &lvl_up, &pyr[num_levels - 1]);
        so_convert_8u_16s(&down_next, &pyr[num_levels]);
        mat_free(&lvl_up); mat_free(&down_next);
        if (cur_own) mat_free(&current);
        return 0;
    }
    /* 16S branch (:477-488) */
    for (int y = 0; y < img->rows; ++y)
        memcpy(ROW(&pyr[0], char, y), ROW(img, const char, y), (size_t)img->cols * cn * 2);
    for (int i = 0; i < num_levels; ++i)
        if (so_pyr_down(&pyr[i], &pyr[i + 1])) return -1;
    for (int i = 0; i < num_levels; ++i) {
        so_mat tmp;
        mat_alloc(&tmp, pyr[i].rows, pyr[i].cols, pyr[i].type);
        if (so_pyr_up(&pyr[i + 1], &tmp)) return -1;
        so_subtract_16s(&pyr[i], &tmp, &pyr[i]);
        mat_free(&tmp);
    }
    return 0;
}

int so_restore_image_from_laplace_pyr(so_mat *pyr, int n)   /* blenders.cpp:520-530 */
{
    for (int i = n - 1; i > 0; --i) {
        so_mat tmp;
        mat_alloc(&tmp, pyr[i - 1].rows, pyr[i - 1].cols, pyr[i - 1].type);
        if (so_pyr_up(&pyr[i], &tmp)) return -1;
        so_add_16s(&tmp, &pyr[i - 1], &pyr[i - 1]);
        mat_free(&tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------ Blender / FeatherBlender / MultiBandBlender */
struct so_blender {
    int kind, actual_num_bands, num_bands, weight_type;
    float sharpness;
    int roi[4];           /* dst_roi_ x,y,w,h */
    int roi_final[4];     /* dst_roi_final_ (multi-band) */
    int prepared;
    so_mat dst, dst_mask, dst_weight_map;
    so_mat *pyr_laplace, *band_weights;
};

so_blender *so_blender_create(int kind, int num_bands, int weight_type, float sharpness)
{
    if (kind < SO_BLEND_NO || kind > SO_BLEND_MULTI_BAND) return NULL;            /* blenders.cpp:60 */
    if (kind == SO_BLEND_MULTI_BAND && weight_type != SO_32F && weight_type != SO_16S) return NULL;   /* :198 */
    so_blender *b = (so_blender *)calloc(1, sizeof *b);
    b->kind = kind; b->actual_num_bands = num_bands; b->weight_type = weight_type; b->sharpness = sharpness;
    return b;
}

static void blender_release(so_blender *b)
{
    if (b->pyr_laplace) {
        for (int i = 1; i <= b->num_bands; ++i) mat_free(&b->pyr_laplace[i]);     /* [0] aliases dst */
        free(b->pyr_laplace); b->pyr_laplace = NULL;
    }
    if (b->band_weights) {
        for (int i = 0; i <= b->num_bands; ++i) mat_free(&b->band_weights[i]);
        free(b->band_weights); b->band_weights = NULL;
    }
    mat_free(&b->dst); mat_free(&b->dst_mask); mat_free(&b->dst_weight_map);
    b->prepared = 0;
}
void so_blender_destroy(so_blender *b) { if (b) { blender_release(b); free(b); } }

int so_blender_prepare_rect(so_blender *b, int x, int y, int w, int h)
{
    blender_release(b);
    b->roi_final[0] = x; b->roi_final[1] = y; b->roi_final[2] = w; b->roi_final[3] = h;
    if (b->kind == SO_BLEND_MULTI_BAND) {                                         /* blenders.cpp:203-233 */
        double max_len = (double)(w > h ? w : h);
        int nb = (int)ceil(log(max_len) / log(2.0));
        b->num_bands = b->actual_num_bands < nb ? b->actual_num_bands : nb;
        int m = 1 << b->num_bands;
        w += (m - w % m) % m;
        h += (m - h % m) % m;
    }
    /* Blender::prepare(Rect) blenders.cpp:71-78 */
    if (mat_alloc(&b->dst, h, w, SO_16SC3) || mat_alloc(&b->dst_mask, h, w, SO_8UC1)) return -1;
    b->roi[0] = x; b->roi[1] = y; b->roi[2] = w; b->roi[3] = h;
    if (b->kind == SO_BLEND_FEATHER) {                                            /* :115-120 */
        if (mat_alloc(&b->dst_weight_map, h, w, SO_32FC1)) return -1;
    } else if (b->kind == SO_BLEND_MULTI_BAND) {
        int wt = b->weight_type == SO_32F ? SO_32FC1 : SO_16SC1;
        b->pyr_laplace = (so_mat *)calloc(b->num_bands + 1, sizeof(so_mat));
        b->band_weights = (so_mat *)calloc(b->num_bands + 1, sizeof(so_mat));
        b->pyr_laplace[0] = b->dst;
        mat_alloc(&b->band_weights[0], h, w, wt);
        for (int i = 1; i <= b->num_bands; ++i) {
            mat_alloc(&b->pyr_laplace[i], (b->pyr_laplace[i - 1].rows + 1) / 2, (b->pyr_laplace[i - 1].cols + 1) / 2, SO_16SC3);
            mat_alloc(&b->band_weights[i], (b->band_weights[i - 1].rows + 1) / 2, (b->band_weights[i - 1].cols + 1) / 2, wt);
        }
    }
    b->prepared = 1;
    return 0;
}

int so_blender_prepare(so_blender *b, const int *corners_xy, const int *sizes_wh, int n)   /* blenders.cpp:65-68 */
{
    int roi[4];
    so_result_roi(corners_xy, sizes_wh, n, roi);
    return so_blender_prepare_rect(b, roi[0], roi[1], roi[2], roi[3]);
}

void so_blender_result_size(const so_blender *b, int *w, int *h)
{
    *w = b->kind == SO_BLEND_MULTI_BAND ? b->roi_final[2] : b->roi[2];
    *h = b->kind == SO_BLEND_MULTI_BAND ? b->roi_final[3] : b->roi[3];
}
int so_blender_num_bands_effective(const so_blender *b) { return b->num_bands; }

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

static int feed_multiband(so_blender *b, const so_mat *img, const so_mat *mask, int tlx, int tly)
{
    /* blenders.cpp:236-356 */
    int nb = b->num_bands, m = 1 << nb;
    int rx = b->roi[0], ry = b->roi[1], rbrx = b->roi[0] + b->roi[2], rbry = b->roi[1] + b->roi[3];
    int gap = 3 * m;
    int tlnx = imax(rx, tlx - gap), tlny = imax(ry, tly - gap);
    int brnx = imin(rbrx, tlx + img->cols + gap), brny = imin(rbry, tly + img->rows + gap);
    tlnx = rx + (((tlnx - rx) >> nb) << nb);
    tlny = ry + (((tlny - ry) >> nb) << nb);
    int width = brnx - tlnx, height = brny - tlny;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    brnx = tlnx + width; brny = tlny + height;
    int dy = imax(brny - rbry, 0), dx = imax(brnx - rbrx, 0);
    tlnx -= dx; brnx -= dx; tlny -= dy; brny -= dy;
    int top = tly - tlny, left = tlx - tlnx;
    int bottom = brny - tly - img->rows, right = brnx - tlx - img->cols;

    so_mat bordered;
    mat_alloc(&bordered, img->rows + top + bottom, img->cols + left + right, img->type);
    if (so_copy_make_border(img, &bordered, top, bottom, left, right, SO_BORDER_REFLECT)) return -1;
    so_mat *src_pyr = (so_mat *)calloc(nb + 1, sizeof(so_mat));
    for (int i = 0, r = bordered.rows, c = bordered.cols; i <= nb; ++i, r = (r + 1) / 2, c = (c + 1) / 2)
        mat_alloc(&src_pyr[i], r, c, SO_16SC3);
    if (so_create_laplace_pyr(&bordered, nb, src_pyr)) return -1;
    mat_free(&bordered);

    /* weight map Gaussian pyramid (:282-298) */
    so_mat wmap;
    int wt = b->weight_type == SO_32F ? SO_32FC1 : SO_16SC1;
    mat_alloc(&wmap, mask->rows, mask->cols, wt);
    for (int y = 0; y < mask->rows; ++y) {
        const uint8_t *mr = ROW(mask, const uint8_t, y);
        if (wt == SO_32FC1) {
            float *w = ROW(&wmap, float, y);
            for (int x = 0; x < mask->cols; ++x) w[x] = (float)mr[x] * (float)(1. / 255.);   /* convertTo(CV_32F, 1./255.) */
        } else {
            short *w = ROW(&wmap, short, y);
            for (int x = 0; x < mask->cols; ++x) w[x] = (short)(mr[x] + (mr[x] != 0));        /* add(w, 1, w, mask != 0) */
        }
    }
    so_mat *wpyr = (so_mat *)calloc(nb + 1, sizeof(so_mat));
    mat_alloc(&wpyr[0], mask->rows + top + bottom, mask->cols + left + right, wt);
    so_copy_make_border(&wmap, &wpyr[0], top, bottom, left, right, SO_BORDER_CONSTANT);
    mat_free(&wmap);
    for (int i = 0; i < nb; ++i) {
        mat_alloc(&wpyr[i + 1], (wpyr[i].rows + 1) / 2, (wpyr[i].cols + 1) / 2, wt);
        so_pyr_down(&wpyr[i], &wpyr[i + 1]);
    }

    int y_tl = tlny - ry, y_br = brny - ry, x_tl = tlnx - rx, x_br = brnx - rx;
    for (int i = 0; i <= nb; ++i) {
        for (int y = y_tl; y < y_br; ++y) {
            int y_ = y - y_tl;
            const short *s = ROW(&src_pyr[i], const short, y_);
            short *d = ROW(&b->pyr_laplace[i], short, y);
            if (wt == SO_32FC1) {
                const float *w = ROW(&wpyr[i], const float, y_);
                float *dw = ROW(&b->band_weights[i], float, y);
                for (int x = x_tl; x < x_br; ++x) {
                    int x_ = x - x_tl;
                    for (int c = 0; c < 3; ++c)
                        d[x * 3 + c] = (short)(d[x * 3 + c] + so_trunc_short(s[x_ * 3 + c] * w[x_]));
                    dw[x] += w[x_];
                }
            } else {
                const short *w = ROW(&wpyr[i], const short, y_);
                short *dw = ROW(&b->band_weights[i], short, y);
                for (int x = x_tl; x < x_br; ++x) {
                    int x_ = x - x_tl;
                    for (int c = 0; c < 3; ++c)
                        d[x * 3 + c] = (short)(d[x * 3 + c] + (short)((s[x_ * 3 + c] * w[x_]) >> 8));
                    dw[x] = (short)(dw[x] + w[x_]);
                }
            }
        }
        x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
    }
    for (int i = 0; i <= nb; ++i) { mat_free(&src_pyr[i]); mat_free(&wpyr[i]); }
    free(src_pyr); free(wpyr);
    return 0;
}

int so_blender_feed(so_blender *b, const so_mat *img, const so_mat *mask, int tlx, int tly)
{
    if (!b->prepared) return -2;
    if (mask->type != SO_8UC1) return -1;
    if (b->kind == SO_BLEND_MULTI_BAND) {
        if (img->type != SO_16SC3 && img->type != SO_8UC3) return -1;              /* :238 */
        b->pyr_laplace[0] = b->dst;
        return feed_multiband(b, img, mask, tlx, tly);
    }
    if (img->type != SO_16SC3) return -1;                                          /* :83,125 */
    int dx = tlx - b->roi[0], dy = tly - b->roi[1];
    if (b->kind == SO_BLEND_NO) {                                                  /* :81-102 */
        for (int y = 0; y < img->rows; ++y) {
            const short *s = ROW(img, const short, y);
            short *d = ROW(&b->dst, short, dy + y);
            const uint8_t *mr = ROW(mask, const uint8_t, y);
            uint8_t *dm = ROW(&b->dst_mask, uint8_t, dy + y);
            for (int x = 0; x < img->cols; ++x) {
                if (mr[x]) memcpy(d + (dx + x) * 3, s + x * 3, 6);
                dm[dx + x] |= mr[x];
            }
        }
        return 0;
    }
    /* FeatherBlender::feed :123-147 */
    so_mat wm;
    mat_alloc(&wm, mask->rows, mask->cols, SO_32FC1);
    so_create_weight_map(mask, b->sharpness, &wm);
    for (int y = 0; y < img->rows; ++y) {
        const short *s = ROW(img, const short, y);
        short *d = ROW(&b->dst, short, dy + y);
        const float *w = ROW(&wm, const float, y);
        float *dw = ROW(&b->dst_weight_map, float, dy + y);
        for (int x = 0; x < img->cols; ++x) {
            for (int c = 0; c < 3; ++c)
                d[(dx + x) * 3 + c] = (short)(d[(dx + x) * 3 + c] + so_trunc_short(s[x * 3 + c] * w[x]));
            dw[dx + x] += w[x];
        }
    }
    mat_free(&wm);
    return 0;
}

/* copies the result out (the reference hands its buffers to the caller and releases them,
 * blenders.cpp:105-112; prepare must be called again before the next panorama) */
int so_blender_blend(so_blender *b, so_mat *dst, so_mat *dst_mask)
{
    if (!b->prepared) return -2;
    int w, h;
    so_blender_result_size(b, &w, &h);
    if (dst->rows != h || dst->cols != w || dst->type != SO_16SC3) return -1;
    if (dst_mask->rows != h || dst_mask->cols != w || dst_mask->type != SO_8UC1) return -1;
    const so_mat *wsrc = NULL;
    if (b->kind == SO_BLEND_FEATHER) {                                             /* :150-155 */
        so_normalize_using_weight_map(&b->dst_weight_map, &b->dst);
        wsrc = &b->dst_weight_map;
    } else if (b->kind == SO_BLEND_MULTI_BAND) {                                   /* :359-377 */
        b->pyr_laplace[0] = b->dst;
        for (int i = 0; i <= b->num_bands; ++i)
            so_normalize_using_weight_map(&b->band_weights[i], &b->pyr_laplace[i]);
        if (so_restore_image_from_laplace_pyr(b->pyr_laplace, b->num_bands + 1)) return -1;
        wsrc = &b->band_weights[0];
    }
    for (int y = 0; y < h; ++y) {
        uint8_t *dm = ROW(dst_mask, uint8_t, y);
        if (wsrc && wsrc->type == SO_32FC1) {
            const float *wr = ROW(wsrc, const float, y);
            for (int x = 0; x < w; ++x) dm[x] = wr[x] > WEIGHT_EPS ? 255 : 0;
        } else if (wsrc) {
            /* compare(Mat 16S, double 1e-5, CMP_GT): integer weights > 0 */
            const short *wr = ROW(wsrc, const short, y);
            for (int x = 0; x < w; ++x) dm[x] = wr[x] > 0 ? 255 : 0;
        } else
            memcpy(dm, ROW(&b->dst_mask, const uint8_t, y), w);
        /* Blender::blend: dst_.setTo(0, dst_mask_ == 0) */
        const short *s = ROW(&b->dst, const short, y);
        short *d = ROW(dst, short, y);
        for (int x = 0; x < w; ++x) {
            if (dm[x]) memcpy(d + x * 3, s + x * 3, 6);
            else d[x * 3] = d[x * 3 + 1] = d[x * 3 + 2] = 0;
        }
    }
    blender_release(b);
    return 0;
}
