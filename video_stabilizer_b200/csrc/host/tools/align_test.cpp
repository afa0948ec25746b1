// align_test — the reference's manual test executable (align_test.cpp) against this
// implementation, using only the public headers it uses (imgproc.hpp, alignment.hpp).
// Same checks: pyramid + gradients + ImageWarp shift at every level (align_test.cpp:43-247),
// the transform-algebra tests (:261-601), the ImageWarp square test (:358-400) and an aligned
// image pair (:625-691) — on procedural images, with the shift verified directly instead of by
// cv::phaseCorrelate.  Unlike upstream it exits non-zero when a check fails.
#include <stdio.h>

#include <chrono>
#include <random>

#include "alignment.hpp"
#include "synth_frames.hpp"

static int g_fail = 0;
#define CHECK(cond, ...) do { if (cond) { printf("[PASS] "); } else { printf("[FAIL] "); g_fail++; } printf(__VA_ARGS__); printf("\n"); } while (0)

static bool near(double a, double b, double eps = 1e-5) { return fabs((float)a - (float)b) <= eps; }

static void TestPyrDown()
{
    cv::Mat gray = synth::to_gray(synth::make_canvas(1280, 720, 7));
    std::vector<Halide::Runtime::Buffer<uint8_t>> scale(6);
    scale[0] = mat_to_halide_buffer_u8(gray);
    for (int i = 1; i < 6; i++) {
        scale[i] = Halide::Runtime::Buffer<uint8_t>(scale[i - 1].width() / 2, scale[i - 1].height() / 2);
        CHECK(PyrDown(scale[i - 1], scale[i]), "PyrDown level %d -> %dx%d", i, scale[i].width(), scale[i].height());
    }
    SimilarityTransform shift;
    shift.TX = 4; shift.TY = 4;
    for (int i = 0; i < 6; i++) {
        Halide::Runtime::Buffer<float> gx(scale[i].width(), scale[i].height()), gy(scale[i].width(), scale[i].height());
        CHECK(GradXY(scale[i], gx, gy), "GradXY level %d", i);
        Halide::Runtime::Buffer<float> warped(scale[i].width(), scale[i].height());
        bool ok = ImageWarp(scale[i], shift.inverse(), warped);
        // content moves by (+4,+4): warped(x,y) == original(x-4,y-4) away from the border
        int bad = 0;
        for (int y = 8; ok && y < scale[i].height() - 8; y++)
            for (int x = 8; x < scale[i].width() - 8; x++)
                if (warped(x, y) != (float)scale[i](x - 4, y - 4)) bad++;
        CHECK(ok && bad == 0, "ImageWarp shift (4,4) at level %d (%d mismatching pixels)", i, bad);
        if (i < 2) {
            int tile = 0;
            Halide::Runtime::Buffer<uint16_t> lmx, lmy;
            auto t0 = std::chrono::steady_clock::now();
            ok = GradArgMax(gx, gy, tile, lmx, lmy);
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            CHECK(ok && lmx.width() == scale[i].width() / tile, "GradArgMax level %d: tile %d, %dx%d tiles, %.0f us", i, tile,
                  lmx.width(), lmx.height(), us);
        }
    }
}

static void TestSimilarityTransformsAll()
{
    {   // inverse / compose by hand (align_test.cpp:261-346)
        SimilarityTransform t;
        t.A = 0.1; t.B = -0.05; t.TX = 10; t.TY = -5;
        Point p{3.0, 4.0};
        Point q = t.inverse().warp(t.warp(p));
        CHECK(near(q.x, p.x) && near(q.y, p.y), "inverse(T)(T(p)) == p");
        SimilarityTransform u;
        u.A = -0.02; u.B = 0.03; u.TX = -1; u.TY = 2;
        Point a = u.warp(t.warp(p)), b = t.compose(u).warp(p);
        CHECK(near(a.x, b.x) && near(a.y, b.y), "compose applies this transform first");
    }
    {   // randomized (align_test.cpp:444-601), seeds as upstream
        std::mt19937 rng(12345);
        std::uniform_real_distribution<double> ab(-0.5, 0.5), tr(-100, 100), pt(-1000, 1000);
        int bad = 0;
        for (int i = 0; i < 50; i++) {
            SimilarityTransform t;
            t.A = ab(rng); t.B = ab(rng); t.TX = tr(rng); t.TY = tr(rng);
            for (int j = 0; j < 10; j++) {
                Point p{pt(rng), pt(rng)};
                Point q = t.inverse().warp(t.warp(p));
                if (fabs(q.x - p.x) > 1e-6 * 1000 || fabs(q.y - p.y) > 1e-6 * 1000) bad++;
            }
        }
        CHECK(bad == 0, "randomized inverse, seed 12345");
        std::mt19937 rng2(6789);
        bad = 0;
        for (int i = 0; i < 50; i++) {
            SimilarityTransform t[3];
            for (auto& s : t) { s.A = ab(rng2) * 0.4; s.B = ab(rng2) * 0.4; s.TX = tr(rng2) * 0.5; s.TY = tr(rng2) * 0.5; }
            Point p{pt(rng2) * 0.5, pt(rng2) * 0.5};
            Point l = t[0].compose(t[1]).compose(t[2]).warp(p), r = t[0].compose(t[1].compose(t[2])).warp(p);
            if (fabs(l.x - r.x) > 1e-6 || fabs(l.y - r.y) > 1e-6) bad++;
        }
        CHECK(bad == 0, "randomized compose associativity, seed 6789");
        std::mt19937 rng3(9999);
        bad = 0;
        for (int i = 0; i < 50; i++) {
            SimilarityTransform t;
            t.A = ab(rng3); t.B = ab(rng3); t.TX = tr(rng3); t.TY = tr(rng3);
            SimilarityTransform id = t.compose(t.inverse());
            if (fabs(id.A) > 1e-9 || fabs(id.B) > 1e-9 || fabs(id.TX) > 1e-9 || fabs(id.TY) > 1e-9) bad++;
        }
        CHECK(bad == 0, "T compose inverse(T) == identity, seed 9999");
    }
}

static void TestImageWarpCorrectness()
{
    const int W = 64, H = 64;
    cv::Mat synthetic(H, W, CV_8UC1, cv::Scalar(0));
    for (int y = 20; y < 30; y++)
        for (int x = 20; x < 30; x++) synthetic.ptr<uint8_t>(y)[x] = 255;
    auto in = mat_to_halide_buffer_u8(synthetic);
    SimilarityTransform t;
    t.TX = 5; t.TY = 7;
    Halide::Runtime::Buffer<float> out(W, H);
    bool ok = ImageWarp(in, t.inverse(), out);
    int x0 = W, x1 = -1, y0 = H, y1 = -1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (out(x, y) > 127.f) { x0 = std::min(x0, x); x1 = std::max(x1, x); y0 = std::min(y0, y); y1 = std::max(y1, y); }
    CHECK(ok && x0 == 25 && x1 == 34 && y0 == 27 && y1 == 36, "ImageWarp moves the square by (5,7): x %d..%d y %d..%d", x0, x1, y0, y1);
}

static void AlignImagePair()
{
    const int W = 1920, H = 1080;
    cv::Mat canvas = synth::make_canvas(W, H, 11);
    SimilarityTransform truth;   // camera motion between the two frames
    truth.A = 0.0008; truth.B = -0.0006; truth.TX = 2.3; truth.TY = -1.7;
    cv::Mat templ = canvas, input = warpBySimilarityTransform(canvas, truth);
    VideoAligner aligner;
    SimilarityTransform t;
    bool first = aligner.AlignNextFrame(templ, t);
    CHECK(!first, "first frame returns false");
    auto t0 = std::chrono::steady_clock::now();
    bool ok = aligner.AlignNextFrame(input, t);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("Alignment success=%d transform: %s (%.2f ms)\n", ok, t.toString().c_str(), ms);
    // the aligner reports previous -> current; content moved by `truth`
    SimilarityTransform err = t.compose(truth.inverse());
    double d = err.maxCornerDisplacement(W, H);
    CHECK(ok && d < 0.5, "recovered transform within %.3f px of the known one", d);
}

int main()
{
    try {
        TestPyrDown();
        TestSimilarityTransformsAll();
        TestImageWarpCorrectness();
        AlignImagePair();
    } catch (const std::exception& e) {
        printf("[FAIL] exception: %s\n", e.what());
        return 2;
    }
    printf("%s (%d failed)\n", g_fail ? "SOME CHECKS FAILED" : "ALL CHECKS PASSED", g_fail);
    return g_fail ? 1 : 0;
}
