// stream_bench — throughput of the drop-in per-frame API exactly as a C++ caller uses it: T host threads, one
// VideoStabilizer each (the reference's own scale-out recipe, grid_search_align.cpp:159-210), all on one GPU; every
// frame comes from ordinary pageable cv::Mat memory and the stabilized frame is a cv::Mat the caller keeps.
//   stream_bench [width height frames threads...]      one JSON line per thread count
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <thread>
#include <vector>

#include "stabilizer.hpp"
#include "synth_frames.hpp"

int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080, N = argc > 3 ? atoi(argv[3]) : 96;
    std::vector<int> counts;
    for (int i = 4; i < argc; i++) counts.push_back(atoi(argv[i]));
    if (counts.empty()) counts = {1, 2, 4, 8};
    try {
        cv::Mat canvas = synth::make_canvas(W, H, 3);
        synth::Jitter jitter(4);
        std::vector<cv::Mat> frames;
        for (int i = 0; i < N; i++) frames.push_back(synth::render(canvas, jitter.next()).clone());   // clone: plain pageable memory
        VideoStabilizerParams params;
        params.crop_pixels = 0;   // video_test.cpp:54
        for (int T : counts) {
            std::vector<std::unique_ptr<VideoStabilizer>> stabs;
            for (int t = 0; t < T; t++) {
                stabs.emplace_back(new VideoStabilizer(params));
                for (int i = 0; i < 12; i++) stabs.back()->processFrame(frames[i]);   // warm up: ring, staging, first outputs
            }
            std::vector<long> produced(T, 0);
            std::vector<unsigned> checksum(T, 0);
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> ths;
            for (int t = 0; t < T; t++)
                ths.emplace_back([&, t] {
                    for (int i = 12; i < N; i++) {
                        cv::Mat out = stabs[t]->processFrame(frames[i]);
                        if (!out.empty()) { produced[t]++; checksum[t] += out.ptr(out.rows / 2)[out.cols]; }
                    }
                });
            for (auto& th : ths) th.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            long total = 0;
            for (long p : produced) total += p;
            printf("{\"api\": \"VideoStabilizer::processFrame\", \"size\": \"%dx%d\", \"threads\": %d, \"frames_per_s\": %.1f, "
                   "\"ms_per_frame_per_stream\": %.3f, \"frames_out\": %ld, \"input\": \"pageable cv::Mat\"}\n",
                   W, H, T, T * (N - 12) / s, 1e3 * s / (N - 12), total);
            fflush(stdout);
        }
    } catch (const std::exception& e) {
        printf("[FAIL] exception: %s\n", e.what());
        return 2;
    }
    return 0;
}
