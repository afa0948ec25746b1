// video_test — the reference's video driver (video_test.cpp) against this implementation:
// a synthetic jittered clip instead of ../recordings/*.mp4, VideoStabilizerParams exactly as
// video_test.cpp:53-54 (crop_pixels = 0), every frame through VideoStabilizer::processFrame.
// It then feeds the same clip to the batched ClipStabilizer and checks that both produce the
// same frames, and reports frames/s for both.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "clip_stabilizer.hpp"
#include "stabilizer.hpp"
#include "synth_frames.hpp"

int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080, N = argc > 3 ? atoi(argv[3]) : 60;
    try {
        printf("Generating %d synthetic %dx%d frames...\n", N, W, H);
        cv::Mat canvas = synth::make_canvas(W, H, 3);
        synth::Jitter jitter(4);
        std::vector<cv::Mat> frames;
        for (int i = 0; i < N; i++) frames.push_back(synth::render(canvas, jitter.next()));

        VideoStabilizerParams params;
        params.crop_pixels = 0;   // video_test.cpp:54
        VideoStabilizer stabilizer(params);
        std::vector<cv::Mat> out;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < N; i++) {
            cv::Mat processed = stabilizer.processFrame(frames[i]);
            if (!processed.empty()) out.push_back(processed);
            if ((i + 1) % 100 == 0) printf("Processed %d frames...\n", i + 1);
        }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("VideoStabilizer::processFrame: %d frames in, %zu out, %.1f frames/s\n", N, out.size(), N / s);
        if ((int)out.size() != N - params.lag) { printf("[FAIL] expected %d output frames\n", N - params.lag); return 1; }

        // batched path on the same clip
        const size_t fb = (size_t)W * H * 3;
        std::vector<uint8_t> in(fb * N), res(fb * N);
        for (int i = 0; i < N; i++)
            for (int y = 0; y < H; y++) memcpy(&in[fb * i + (size_t)y * W * 3], frames[i].ptr(y), (size_t)W * 3);
        vstab::ClipStabilizer batch(0, W, H, N, params);
        t0 = std::chrono::steady_clock::now();
        int produced = batch.feed(in.data(), N, (int64_t)W * 3, (int64_t)fb, VS_MEM_HOST, res.data(), (int64_t)fb, VS_MEM_HOST);
        s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("ClipStabilizer::feed: %d frames out, %.1f frames/s (pageable host buffers)\n", produced, N / s);
        int mismatched = produced == (int)out.size() ? 0 : 1;
        for (int i = 0; i < produced && !mismatched; i++)
            for (int y = 0; y < H; y++)
                if (memcmp(out[i].ptr(y), &res[fb * i + (size_t)y * W * 3], (size_t)W * 3) != 0) { mismatched = 1; break; }
        printf("%s batched output %s frame-by-frame output\n", mismatched ? "[FAIL]" : "[PASS]", mismatched ? "differs from" : "equals");
        return mismatched;
    } catch (const std::exception& e) {
        printf("[FAIL] exception: %s\n", e.what());
        return 2;
    }
}
