// C++ parity test of the host adapter classes (include/stitchb200.hpp) through the C ABI.
//
// Reads like the reference's own per-image loop (LIB/src/stitcher.cpp:221-313):
//     warper->warp(img) ; warper->warp(mask, INTER_NEAREST, BORDER_CONSTANT) ; compensator->apply ;
//     img_warped.convertTo(CV_16S) ; blender->prepare(corners, sizes) ; blender->feed ; blender->blend
// and checks every stage bit-for-bit against the CPU oracle (oracle/stitch_oracle.h, test infrastructure),
// then checks that sb200::Compositor (the fused per-frame path) returns the same panorama.
//   test_adapters            : needs a CUDA device, exits 0 when everything matches
//   test_adapters --no-device: asserts that construction throws CV_GpuApiCallError (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "stitchb200.hpp"
#include "stitch_oracle.h"

using namespace sb200;

static int g_fail = 0;
#define EXPECT(c, ...) do { if (!(c)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } } while (0)

static so_mat so(const Mat &m) { so_mat s; s.data = m.data; s.rows = m.rows; s.cols = m.cols; s.type = m.type(); s.step = m.step; return s; }
static size_t diff(const Mat &a, const Mat &b)
{
    if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) return (size_t)-1;
    size_t d = 0, rowbytes = (size_t)a.cols * elem_size(a.type());
    for (int y = 0; y < a.rows; ++y) {
        const unsigned char *p = a.ptr<unsigned char>(y), *q = b.ptr<unsigned char>(y);
        for (size_t i = 0; i < rowbytes; ++i) d += p[i] != q[i];
    }
    return d;
}

int main(int argc, char **argv)
{
    const int W = 240, H = 136, N = 5;
    const float f = 131.f;
    if (argc > 1 && !std::strcmp(argv[1], "--no-device")) {
        int thrown = 0;
        try { SphericalWarper w(f); } catch (const Exception &e) { thrown += e.code == SB_ERR_CUDA; }
        try { MultiBandBlender b; } catch (const Exception &e) { thrown += e.code == SB_ERR_CUDA; }
        try { Blender::createDefault(7); } catch (const Exception &e) { thrown += e.code == SB_ERR_BAD_ARG; }
        std::printf("no-device: %d of 3 expected exceptions\n", thrown);
        return thrown == 3 ? 0 : 1;
    }
    std::vector<float> K(9 * N, 0.f), R(9 * N, 0.f);
    for (int i = 0; i < N; ++i) {
        float *k = &K[9 * i], *r = &R[9 * i];
        k[0] = f; k[2] = W / 2.f; k[4] = f; k[5] = H / 2.f; k[8] = 1.f;
        const double a = (180.0 + i * 360.0 / N) * M_PI / 180.0;
        r[0] = (float)std::cos(a); r[2] = (float)std::sin(a); r[4] = 1.f; r[6] = (float)-std::sin(a); r[8] = (float)std::cos(a);
    }
    const double gains[N] = {0.95, 1.02, 1.00, 0.98, 1.05};
    // seeded frames: LCG noise smoothed once horizontally
    std::vector<Mat> frames(N);
    unsigned lcg = 12345u;
    for (int i = 0; i < N; ++i) {
        frames[i].create(H, W, SB_8UC3);
        for (int y = 0; y < H; ++y) {
            unsigned char *p = frames[i].ptr<unsigned char>(y);
            int prev[3] = {128, 128, 128};
            for (int x = 0; x < W * 3; ++x) {
                lcg = lcg * 1664525u + 1013904223u;
                prev[x % 3] = (prev[x % 3] * 3 + (int)(lcg >> 24)) / 4;
                p[x] = (unsigned char)prev[x % 3];
            }
        }
    }
    try {
        SphericalWarperCreator creator;
        std::unique_ptr<RotationWarper> warper = creator.create(f);
        std::unique_ptr<ExposureCompensator> comp = ExposureCompensator::createDefault(ExposureCompensator::GAIN);
        comp->setGains(std::vector<double>(gains, gains + N));
        Mat ones(H, W, SB_8UC1);
        std::memset(ones.data, 255, (size_t)H * W);

        for (int wt = 0; wt < 2; ++wt) {
            const int weight_type = wt ? SB_16S : SB_32F;
            MultiBandBlender blender(false, 5, weight_type);
            so_blender *ob = so_blender_create(SO_BLEND_MULTI_BAND, 5, weight_type, 0.02f);
            std::vector<Point> corners(N);
            std::vector<Size> sizes(N);
            std::vector<Mat> warped(N), masks(N), warped_s(N);
            for (int i = 0; i < N; ++i) {
                const float *Ki = &K[9 * i], *Ri = &R[9 * i];
                corners[i] = warper->warp(frames[i], Ki, Ri, SB_INTER_LINEAR, SB_BORDER_REFLECT, warped[i]);     // stitcher.cpp:275
                warper->warp(ones, Ki, Ri, SB_INTER_NEAREST, SB_BORDER_CONSTANT, masks[i]);                      // :278-280
                sizes[i] = warped[i].size();
                // oracle: buildMaps + remap
                so_projector p;
                so_projector_set(&p, SO_WARP_SPHERICAL, f, Ki, Ri, nullptr);
                int tl[2], br[2];
                so_detect_result_roi(&p, W, H, tl, br);
                EXPECT(tl[0] == corners[i].x && tl[1] == corners[i].y, "corner of camera %d", i);
                Mat xm(br[1] - tl[1] + 1, br[0] - tl[0] + 1, SB_32FC1), ym(br[1] - tl[1] + 1, br[0] - tl[0] + 1, SB_32FC1);
                so_mat sxm = so(xm), sym = so(ym);
                so_build_maps(&p, tl, br, &sxm, &sym);
                Mat gx, gy;
                const Rect roi = warper->buildMaps(Size(W, H), Ki, Ri, gx, gy);                                  // warpers_inl.hpp:62-85
                EXPECT(roi.x == tl[0] && roi.width == br[0] - tl[0], "buildMaps roi camera %d", i);
                EXPECT(diff(gx, xm) == 0 && diff(gy, ym) == 0, "maps of camera %d differ", i);
                Mat ow(xm.rows, xm.cols, SB_8UC3), om(xm.rows, xm.cols, SB_8UC1);
                so_mat sf = so(frames[i]), sow = so(ow), s1 = so(ones), som = so(om);
                const uint8_t bv[4] = {0, 0, 0, 0};
                so_remap(&sf, &sow, &sxm, &sym, SO_INTER_LINEAR, SO_BORDER_REFLECT, bv);
                so_remap(&s1, &som, &sxm, &sym, SO_INTER_NEAREST, SO_BORDER_CONSTANT, bv);
                EXPECT(diff(warped[i], ow) == 0, "warped image %d differs (%zu)", i, diff(warped[i], ow));
                EXPECT(diff(masks[i], om) == 0, "warped mask %d differs", i);
                comp->apply(i, corners[i], warped[i], masks[i]);                                                 // :283
                so_gain_apply(&sow, gains[i]);
                EXPECT(diff(warped[i], ow) == 0, "gain-applied image %d differs", i);
                warped_s[i].create(warped[i].rows, warped[i].cols, SB_16SC3);                                    // :285 convertTo(CV_16S)
                for (int y = 0; y < warped[i].rows; ++y)
                    for (int x = 0; x < warped[i].cols * 3; ++x) warped_s[i].ptr<short>(y)[x] = warped[i].ptr<unsigned char>(y)[x];
            }
            if (wt == 0) {   // ExposureCompensator::feed as stitcher.cpp:209 calls it, on the warped images (exposure_compensate.cpp:76-147)
                GainCompensator est;
                est.feed(corners, warped, masks);
                std::vector<so_mat> si, sm;
                std::vector<int> c2;
                std::vector<unsigned char> vals(N, 255);
                for (int i = 0; i < N; ++i) { si.push_back(so(warped[i])); sm.push_back(so(masks[i])); c2.push_back(corners[i].x); c2.push_back(corners[i].y); }
                std::vector<double> og(N), gg = est.gains();
                EXPECT(so_gain_feed(N, c2.data(), si.data(), sm.data(), vals.data(), og.data()) == 0, "oracle gain feed");
                for (int i = 0; i < N; ++i) EXPECT(std::fabs(gg[i] - og[i]) <= 1e-11 * std::fabs(og[i]), "estimated gain %d: %.17g vs %.17g", i, gg[i], og[i]);
            }
            blender.prepare(corners, sizes);                                                                     // :296-300
            std::vector<int> cxy, swh;
            for (int i = 0; i < N; ++i) { cxy.push_back(corners[i].x); cxy.push_back(corners[i].y); swh.push_back(sizes[i].width); swh.push_back(sizes[i].height); }
            so_blender_prepare(ob, cxy.data(), swh.data(), N);
            for (int i = 0; i < N; ++i) {
                blender.feed(warped_s[i], masks[i], corners[i]);                                                 // :303
                so_mat si = so(warped_s[i]), sm = so(masks[i]);
                so_blender_feed(ob, &si, &sm, corners[i].x, corners[i].y);
            }
            Mat result, result_mask;
            blender.blend(result, result_mask);                                                                  // :307
            int ow_, oh_;
            so_blender_result_size(ob, &ow_, &oh_);
            Mat oresult(oh_, ow_, SB_16SC3), omask(oh_, ow_, SB_8UC1);
            so_mat sr = so(oresult), sm = so(omask);
            so_blender_blend(ob, &sr, &sm);
            so_blender_destroy(ob);
            EXPECT(diff(result, oresult) == 0, "blend result differs (weight_type %d): %zu", weight_type, diff(result, oresult));
            EXPECT(diff(result_mask, omask) == 0, "blend mask differs (weight_type %d)", weight_type);
            bool threw = false;
            try { blender.blend(result, result_mask); } catch (const Exception &e) { threw = e.code == SB_ERR_ASSERT; }
            EXPECT(threw, "second blend() without prepare() must throw CV_StsAssert");

            // the fused frame loop gives the same panorama (16S output to compare like with like)
            Compositor::Config cfg;
            cfg.src_size = Size(W, H); cfg.warper_kind = SB_WARP_SPHERICAL; cfg.warper_scale = f; cfg.K = K; cfg.R = R;
            cfg.blender_kind = SB_BLEND_MULTI_BAND; cfg.num_bands = 5; cfg.weight_type = weight_type;
            cfg.gains.assign(gains, gains + N); cfg.output_type = SB_16SC3;
            Compositor compositor(cfg);
            Mat pano, pano_mask;
            compositor.compose(frames, pano, pano_mask);
            EXPECT(diff(pano, oresult) == 0, "compositor panorama differs (weight_type %d): %zu", weight_type, diff(pano, oresult));
            EXPECT(diff(pano_mask, omask) == 0, "compositor mask differs (weight_type %d)", weight_type);
            for (int i = 0; i < N; ++i) {
                const Rect r = compositor.cameraRoi(i);
                EXPECT(r.x == corners[i].x && r.y == corners[i].y && r.width == sizes[i].width && r.height == sizes[i].height, "camera roi %d", i);
            }
        }
        // error behaviour mirrors the reference
        int code = 0;
        try { Blender::createDefault(7); } catch (const Exception &e) { code = e.code; }
        EXPECT(code == SB_ERR_BAD_ARG, "Blender::createDefault(7) -> %d", code);
        code = 0;
        try { MultiBandBlender b(false, 5, SB_8U); } catch (const Exception &e) { code = e.code; }
        EXPECT(code == SB_ERR_ASSERT, "weight_type assert -> %d", code);
        code = 0;
        try { FeatherBlender fb; Mat a(4, 4, SB_16SC3), m(4, 4, SB_8UC1); fb.feed(a, m, Point(0, 0)); } catch (const Exception &e) { code = e.code; }
        EXPECT(code == SB_ERR_ASSERT, "feed before prepare -> %d", code);
    } catch (const Exception &e) {
        std::printf("FAIL: exception %d: %s\n", e.code, e.what());
        return 2;
    }
    if (g_fail) std::printf("test_adapters: %d failures\n", g_fail);
    else std::printf("test_adapters: ok (%llu kernel launches)\n", (unsigned long long)sb_kernel_launch_count());
    return g_fail ? 1 : 0;
}
