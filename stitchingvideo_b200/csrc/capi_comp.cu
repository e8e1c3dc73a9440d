// capi_comp.cu — ExposureCompensator C ABI (exposure_compensate.hpp:51-101).  apply() is the per-frame part
// (exposure_compensate.cpp:150-153, 225-246).  feed() (gain estimation, :76-147 and :165-222) is calibration: its
// result may be handed in (sb_comp_set_gains / sb_comp_set_gain_maps), or computed here with the overlap statistics
// reduced on the device (sb_comp_feed) — the small linear system is solved on the host like
// cv::solve does.
#include <algorithm>
#include <cmath>

#include "sb_kernels.h"

using namespace sb;

struct sb_comp {
    int device = 0;
    int kind = 0;
    cudaStream_t stream = nullptr;
    std::vector<double> gains;
    std::vector<DevImage> gain_maps;       // block gain maps as fed (small)
    std::vector<DevImage> gain_full;       // resized to the image size, cached (sequence-constant)
    std::vector<std::vector<float>> gain_maps_host;   // block gain maps computed by sb_comp_feed (row-major)
    std::vector<sb_size> gain_map_sizes;
    int bl_width = 32, bl_height = 32;     // BlocksGainCompensator(bl_width = 32, bl_height = 32), exposure_compensate.hpp:92
    DevImage stage;
};

namespace {

struct FeedImage { DImage img, mask; sb_point corner; int val; };

// util.cpp:103-116 overlapRoi
bool overlap_roi(sb_point tl1, sb_point tl2, int w1, int h1, int w2, int h2, sb_rect &roi)
{
    const int x_tl = std::max(tl1.x, tl2.x), y_tl = std::max(tl1.y, tl2.y);
    const int x_br = std::min(tl1.x + w1, tl2.x + w2), y_br = std::min(tl1.y + h1, tl2.y + h2);
    if (x_tl < x_br && y_tl < y_br) { roi = sb_rect{x_tl, y_tl, x_br - x_tl, y_br - y_tl}; return true; }
    return false;
}

// cv::solve(A, b, x, DECOMP_LU) on doubles: Gaussian elimination with partial pivoting, in place
bool solve_lu(int n, std::vector<double> &A, std::vector<double> &b)
{
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (std::fabs(A[(size_t)j * n + i]) > std::fabs(A[(size_t)k * n + i])) k = j;
        if (std::fabs(A[(size_t)k * n + i]) < 2.220446049250313e-16 * 100) return false;
        if (k != i) {
            for (int j = i; j < n; ++j) std::swap(A[(size_t)i * n + j], A[(size_t)k * n + j]);
            std::swap(b[i], b[k]);
        }
        const double d = -1 / A[(size_t)i * n + i];
        for (int j = i + 1; j < n; ++j) {
            const double alpha = A[(size_t)j * n + i] * d;
            if (alpha == 0) continue;                       // (block systems are sparse: adding 0 * row changes nothing)
            for (k = i + 1; k < n; ++k) A[(size_t)j * n + k] += alpha * A[(size_t)i * n + k];
            b[j] += alpha * b[i];
        }
        A[(size_t)i * n + i] = -d;
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= A[(size_t)i * n + k] * b[k];
        b[i] = s * A[(size_t)i * n + i];
    }
    return true;
}

}  // namespace

// The normal equations of the gain model (exposure_compensate.cpp:128-144) from the overlap statistics of the pairs
// (i <= j, each once; i == j is an image's own mask count) and their solution.  Host-only.
//   b_i = A_ii = beta * sum_j N_ij (+ 2 alpha sum_{j != i} I_ij^2 N_ij),  A_ij = -2 alpha I_ij I_ji N_ij
// Up to SB_GAIN_DENSE_MAX unknowns: the reference's own dense LU with partial pivoting (cv::solve).  Beyond that — the
// block compensator has one unknown per 32x32 block, ~13 k for five 1080p cameras, where the reference's dense n x n
// system needs 1.4 GB and an O(n^3) factorisation — the same equations are kept sparse (A is symmetric positive
// definite and strongly diagonally dominant through the beta term) and solved by Jacobi-preconditioned conjugate
// gradients to a relative residual of 1e-15: the gains agree with the dense solution to ~1e-13.
extern "C" int sb_gain_solve(int n, int n_pairs, const int *pi, const int *pj, const double *N, const double *Iij, const double *Iji, double *gains)
{
    SB_ASSERT(n >= 0 && n_pairs >= 0 && gains && (n_pairs == 0 || (pi && pj && N && Iij && Iji)));
    const double alpha = 0.01, beta = 100;
    std::vector<double> diag(n, 0.0), b(n, 0.0);
    for (int k = 0; k < n_pairs; ++k) {
        const int i = pi[k], j = pj[k];
        SB_ASSERT(i >= 0 && j >= i && j < n);
        b[i] += beta * N[k]; diag[i] += beta * N[k];
        if (i == j) continue;
        b[j] += beta * N[k]; diag[j] += beta * N[k];
        diag[i] += 2 * alpha * Iij[k] * Iij[k] * N[k];
        diag[j] += 2 * alpha * Iji[k] * Iji[k] * N[k];
    }
    const int SB_GAIN_DENSE_MAX = 512;
    if (n <= SB_GAIN_DENSE_MAX) {
        std::vector<double> A((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) A[(size_t)i * n + i] = diag[i];
        for (int k = 0; k < n_pairs; ++k) {
            const int i = pi[k], j = pj[k];
            if (i == j) continue;
            A[(size_t)i * n + j] -= 2 * alpha * Iij[k] * Iji[k] * N[k];
            A[(size_t)j * n + i] -= 2 * alpha * Iji[k] * Iij[k] * N[k];
        }
        if (!solve_lu(n, A, b)) return fail(SB_ERR_ASSERT, "exposure compensation: singular normal equations");
        std::copy(b.begin(), b.end(), gains);
        return SB_OK;
    }
    for (int i = 0; i < n; ++i)
        if (!(diag[i] > 0)) return fail(SB_ERR_ASSERT, "exposure compensation: singular normal equations");
    std::vector<double> off(n_pairs, 0.0);
    for (int k = 0; k < n_pairs; ++k) off[k] = pi[k] == pj[k] ? 0.0 : -2 * alpha * Iij[k] * Iji[k] * N[k];
    auto apply = [&](const std::vector<double> &x, std::vector<double> &y) {      // y = A x
        for (int i = 0; i < n; ++i) y[i] = diag[i] * x[i];
        for (int k = 0; k < n_pairs; ++k) { y[pi[k]] += off[k] * x[pj[k]]; y[pj[k]] += off[k] * x[pi[k]]; }
    };
    std::vector<double> x(n), r(n), z(n), p(n), q(n);
    for (int i = 0; i < n; ++i) x[i] = b[i] / diag[i];
    apply(x, q);
    double bnorm = 0, rz = 0;
    for (int i = 0; i < n; ++i) { r[i] = b[i] - q[i]; z[i] = r[i] / diag[i]; p[i] = z[i]; rz += r[i] * z[i]; bnorm += b[i] * b[i]; }
    for (int it = 0; it < 20 * n + 100; ++it) {
        double rn = 0;
        for (int i = 0; i < n; ++i) rn += r[i] * r[i];
        if (rn <= 1e-30 * bnorm) break;
        apply(p, q);
        double pq = 0;
        for (int i = 0; i < n; ++i) pq += p[i] * q[i];
        if (!(pq > 0)) break;
        const double a = rz / pq;
        double rz2 = 0;
        for (int i = 0; i < n; ++i) { x[i] += a * p[i]; r[i] -= a * q[i]; z[i] = r[i] / diag[i]; rz2 += r[i] * z[i]; }
        const double be = rz2 / rz;
        rz = rz2;
        for (int i = 0; i < n; ++i) p[i] = z[i] + be * p[i];
    }
    std::copy(x.begin(), x.end(), gains);
    return SB_OK;
}

namespace {

// GainCompensator::feed (exposure_compensate.cpp:76-147) over `im` (whole images, or the blocks of
// BlocksGainCompensator::feed): pair list on the host, statistics on the device, normal equations + LU on the host.
int gain_feed(const std::vector<FeedImage> &im, cudaStream_t s, std::vector<double> &gains)
{
    const int n = (int)im.size();
    std::vector<OverlapPair> pairs;
    std::vector<std::pair<int, int>> ij;
    int max_h = 1;
    // candidate pairs: the reference's own double loop (one rectangle test per pair; 85 M tests for 13 k blocks, ~0.3 s)
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            sb_rect roi;
            if (!overlap_roi(im[i].corner, im[j].corner, im[i].img.cols, im[i].img.rows, im[j].img.cols, im[j].img.rows, roi)) continue;
            OverlapPair p{};
            const int x1 = roi.x - im[i].corner.x, y1 = roi.y - im[i].corner.y, x2 = roi.x - im[j].corner.x, y2 = roi.y - im[j].corner.y;
            p.img1 = im[i].img.ptr<uint8_t>() + (size_t)y1 * im[i].img.step + 3 * x1; p.istep1 = (unsigned)im[i].img.step;
            p.img2 = im[j].img.ptr<uint8_t>() + (size_t)y2 * im[j].img.step + 3 * x2; p.istep2 = (unsigned)im[j].img.step;
            p.mask1 = im[i].mask.ptr<uint8_t>() + (size_t)y1 * im[i].mask.step + x1; p.mstep1 = (unsigned)im[i].mask.step;
            p.mask2 = im[j].mask.ptr<uint8_t>() + (size_t)y2 * im[j].mask.step + x2; p.mstep2 = (unsigned)im[j].mask.step;
            p.w = roi.width; p.h = roi.height; p.val1 = im[i].val; p.val2 = im[j].val;
            pairs.push_back(p); ij.emplace_back(i, j);
            max_h = std::max(max_h, roi.height);
        }
    const int np = (int)pairs.size();
    DevBuf dpairs, dout;
    std::vector<unsigned long long> out((size_t)np * 5);
    if (np) {
        SB_TRY(dpairs.ensure(sizeof(OverlapPair) * np));
        SB_TRY(dout.ensure(sizeof(unsigned long long) * 5 * np));
        SB_CUDA(cudaMemcpyAsync(dpairs.p, pairs.data(), sizeof(OverlapPair) * np, cudaMemcpyHostToDevice, s));
        SB_TRY(launch_overlap_stats(static_cast<const OverlapPair *>(dpairs.p), np, max_h, static_cast<unsigned long long *>(dout.p), s));
        SB_CUDA(cudaMemcpyAsync(out.data(), dout.p, sizeof(unsigned long long) * 5 * np, cudaMemcpyDeviceToHost, s));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    // N, I (exposure_compensate.cpp:86-87,106,123-124)
    std::vector<int> pi(np), pj(np);
    std::vector<double> N(np), Iij(np), Iji(np);
    auto exact = [](unsigned long long lo, unsigned long long hi) {      // (hi * 2^32 + lo) * 2^-52, rounded once
        const unsigned __int128 t = ((unsigned __int128)hi << 32) + lo;
        return std::ldexp((double)t, -52);
    };
    for (int k = 0; k < np; ++k) {
        pi[k] = ij[k].first; pj[k] = ij[k].second;
        N[k] = (double)std::max<unsigned long long>(1, out[(size_t)k * 5]);
        Iij[k] = exact(out[(size_t)k * 5 + 1], out[(size_t)k * 5 + 2]) / N[k];
        Iji[k] = exact(out[(size_t)k * 5 + 3], out[(size_t)k * 5 + 4]) / N[k];
    }
    gains.assign(n, 0.0);
    return sb_gain_solve(n, np, pi.data(), pj.data(), N.data(), Iij.data(), Iji.data(), gains.data());
}

// cv::sepFilter2D(m, m, CV_32F, ker, ker), ker = {0.25, 0.5, 0.25}, BORDER_REFLECT_101 (exposure_compensate.cpp:217-218):
// S[i] * k0 + (S[i-1] + S[i+1]) * k1, rows then columns (OpenCV's symmetric small-kernel filters)
void smooth_gain_map(std::vector<float> &m, int w, int h)
{
    auto refl = [](int p, int len) { if (len == 1) return 0; if (p < 0) return -p; if (p >= len) return 2 * len - 2 - p; return p; };
    std::vector<float> t(m.size());
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) t[(size_t)y * w + x] = m[(size_t)y * w + x] * 0.5f + (m[(size_t)y * w + refl(x - 1, w)] + m[(size_t)y * w + refl(x + 1, w)]) * 0.25f;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) m[(size_t)y * w + x] = t[(size_t)y * w + x] * 0.5f + (t[(size_t)refl(y - 1, h) * w + x] + t[(size_t)refl(y + 1, h) * w + x]) * 0.25f;
}

}  // namespace

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

int sb_comp_set_block_size(sb_comp *c, int bl_width, int bl_height)
{
    SB_ASSERT(c && bl_width > 0 && bl_height > 0);
    c->bl_width = bl_width; c->bl_height = bl_height;
    return SB_OK;
}

// ExposureCompensator::feed(corners, images, masks) (exposure_compensate.cpp:64-71 -> :76-147 / :165-222)
int sb_comp_feed(sb_comp *c, const sb_point *corners, const sb_image *images, const sb_image *masks, int n)
{
    SB_ASSERT(c && n >= 0 && (n == 0 || (corners && images && masks)));       // corners.size() == images.size() == masks.size()
    if (c->kind == SB_COMP_NO) return SB_OK;
    DeviceGuard g(c->device);
    if (!g.ok) return SB_ERR_CUDA;
    std::vector<DevImage> stage(2 * (size_t)n);
    std::vector<FeedImage> whole(n);
    for (int i = 0; i < n; ++i) {
        SB_TRY(check_image(&images[i], "image"));
        SB_TRY(check_image(&masks[i], "mask"));
        SB_ASSERT(images[i].type == SB_8UC3 && masks[i].type == SB_8UC1 && images[i].rows == masks[i].rows && images[i].cols == masks[i].cols);
        SB_TRY(to_device(images[i], stage[2 * i], c->stream, &whole[i].img));
        SB_TRY(to_device(masks[i], stage[2 * i + 1], c->stream, &whole[i].mask));
        whole[i].corner = corners[i]; whole[i].val = 255;
    }
    if (c->kind == SB_COMP_GAIN) {
        SB_TRY(gain_feed(whole, c->stream, c->gains));
        return SB_OK;
    }
    // BlocksGainCompensator::feed: the block lists of exposure_compensate.cpp:178-201
    std::vector<FeedImage> blocks;
    std::vector<sb_size> per(n);
    for (int i = 0; i < n; ++i) {
        const int cols = whole[i].img.cols, rows = whole[i].img.rows;
        per[i] = sb_size{(cols + c->bl_width - 1) / c->bl_width, (rows + c->bl_height - 1) / c->bl_height};
        const int bw = (cols + per[i].width - 1) / per[i].width, bh = (rows + per[i].height - 1) / per[i].height;
        for (int by = 0; by < per[i].height; ++by)
            for (int bx = 0; bx < per[i].width; ++bx) {
                const int x0 = bx * bw, y0 = by * bh, x1 = std::min(x0 + bw, cols), y1 = std::min(y0 + bh, rows);
                FeedImage f = whole[i];
                f.img.data = whole[i].img.ptr<uint8_t>() + (size_t)y0 * whole[i].img.step + 3 * x0; f.img.cols = x1 - x0; f.img.rows = y1 - y0;
                f.mask.data = whole[i].mask.ptr<uint8_t>() + (size_t)y0 * whole[i].mask.step + x0; f.mask.cols = x1 - x0; f.mask.rows = y1 - y0;
                f.corner = sb_point{corners[i].x + x0, corners[i].y + y0};
                blocks.push_back(f);
            }
    }
    std::vector<double> gains;
    SB_TRY(gain_feed(blocks, c->stream, gains));
    c->gain_maps_host.assign(n, {}); c->gain_map_sizes = per;
    std::vector<sb_image> maps(n);
    size_t b = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<float> &m = c->gain_maps_host[i];
        m.resize((size_t)per[i].width * per[i].height);
        for (size_t k = 0; k < m.size(); ++k) m[k] = (float)gains[b++];
        smooth_gain_map(m, per[i].width, per[i].height);
        smooth_gain_map(m, per[i].width, per[i].height);
        maps[i] = sb_image{m.data(), per[i].height, per[i].width, SB_32FC1, sizeof(float) * per[i].width, -1};
    }
    return sb_comp_set_gain_maps(c, maps.data(), n);
}

int sb_comp_num_gains(const sb_comp *c)
{
    return c ? (int)(c->kind == SB_COMP_GAIN_BLOCKS ? c->gain_maps_host.size() : c->gains.size()) : 0;
}

int sb_comp_gain_map_size(const sb_comp *c, int index, sb_size *size)
{
    SB_ASSERT(c && size && index >= 0 && index < (int)c->gain_map_sizes.size());
    *size = c->gain_map_sizes[index];
    return SB_OK;
}

// gain_maps_[index] as computed by feed: host CV_32FC1 of sb_comp_gain_map_size
int sb_comp_get_gain_map(const sb_comp *c, int index, sb_image *map)
{
    SB_ASSERT(c && map && map->data && map->device < 0 && index >= 0 && index < (int)c->gain_maps_host.size());
    const sb_size sz = c->gain_map_sizes[index];
    SB_ASSERT(map->type == SB_32FC1 && map->rows == sz.height && map->cols == sz.width);
    for (int y = 0; y < sz.height; ++y)
        std::memcpy(static_cast<char *>(map->data) + (size_t)y * map->step, c->gain_maps_host[index].data() + (size_t)y * sz.width, sizeof(float) * sz.width);
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

// cv::dilate(src, dst, Mat()) on 8UC1 (call site stitcher.cpp:291)
int sb_dilate3x3(const sb_image *src, sb_image *dst, int device)
{
    SB_ASSERT(src && dst);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_TRY(check_image(src, "src"));
    SB_ASSERT(src->type == SB_8UC1 && dst->type == SB_8UC1 && dst->rows == src->rows && dst->cols == src->cols);
    DevImage s0, s1;
    DImage a, d;
    SB_TRY(to_device(*src, s0, nullptr, &a));
    if (dst->device >= 0 && dst->data) { d.data = dst->data; d.rows = dst->rows; d.cols = dst->cols; d.type = dst->type; d.step = dst->step; }
    else { SB_TRY(s1.create(dst->rows, dst->cols, SB_8UC1)); d = s1.v; }
    SB_TRY(launch_dilate3x3(a, d, nullptr));
    if (dst->device < 0) SB_TRY(from_device(d, dst, nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

// cv::resize(src, dst, dst.size(), 0, 0, INTER_LINEAR) on 8UC1 (call site stitcher.cpp:292)
int sb_resize_linear_8u(const sb_image *src, sb_image *dst, int device)
{
    SB_ASSERT(src && dst);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_TRY(check_image(src, "src"));
    SB_ASSERT(src->type == SB_8UC1 && dst->type == SB_8UC1 && dst->rows > 0 && dst->cols > 0);
    DevImage s0, s1;
    DImage a, d;
    SB_TRY(to_device(*src, s0, nullptr, &a));
    if (dst->device >= 0 && dst->data) { d.data = dst->data; d.rows = dst->rows; d.cols = dst->cols; d.type = dst->type; d.step = dst->step; }
    else { SB_TRY(s1.create(dst->rows, dst->cols, SB_8UC1)); d = s1.v; }
    SB_TRY(launch_resize_linear_8u(a, d, nullptr, nullptr));
    if (dst->device < 0) SB_TRY(from_device(d, dst, nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

// The seam-mask refinement of the compose loop (stitcher.cpp:291-294; SAMPLE:731-735):
//   dilate(masks_warped[i], dilated, Mat()); resize(dilated, seam_mask, mask_warped.size()); out = seam_mask & mask_warped
// Two launches (the AND is fused into the resize).
int sb_refine_seam_mask(const sb_image *seam_mask, const sb_image *mask_warped, sb_image *out, int device)
{
    SB_ASSERT(seam_mask && mask_warped && out);
    DeviceGuard g(device);
    if (!g.ok) return SB_ERR_CUDA;
    SB_TRY(check_image(seam_mask, "seam_mask"));
    SB_TRY(check_image(mask_warped, "mask_warped"));
    SB_ASSERT(seam_mask->type == SB_8UC1 && mask_warped->type == SB_8UC1 && out->type == SB_8UC1);
    SB_ASSERT(out->rows == mask_warped->rows && out->cols == mask_warped->cols);
    DevImage s0, s1, s2, s3;
    DImage a, m, d;
    SB_TRY(to_device(*seam_mask, s0, nullptr, &a));
    SB_TRY(to_device(*mask_warped, s1, nullptr, &m));
    SB_TRY(s2.create(a.rows, a.cols, SB_8UC1));
    if (out->device >= 0 && out->data) { d.data = out->data; d.rows = out->rows; d.cols = out->cols; d.type = out->type; d.step = out->step; }
    else { SB_TRY(s3.create(out->rows, out->cols, SB_8UC1)); d = s3.v; }
    SB_TRY(launch_dilate3x3(a, s2.v, nullptr));
    SB_TRY(launch_resize_linear_8u(s2.v, d, &m, nullptr));
    if (out->device < 0) SB_TRY(from_device(d, out, nullptr));
    SB_CUDA(cudaStreamSynchronize(nullptr));
    return SB_OK;
}

}  // extern "C"
