// grid_search_align — the reference's parameter sweep (grid_search_align.cpp) on the batched API: a synthetic jittered
// clip, its jitter, the 54 VideoAlignerParams combinations in one solver launch, the best combination by output / input
// jitter ratio.   grid_search_align [width height frames]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "grid_search.hpp"
#include "synth_frames.hpp"

int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1280, H = argc > 2 ? atoi(argv[2]) : 720, N = argc > 3 ? atoi(argv[3]) : 48;
    try {
        cv::Mat canvas = synth::make_canvas(W, H, 5);
        synth::Jitter jitter(6);
        const size_t fb = (size_t)W * H * 3;
        std::vector<uint8_t> frames(fb * N);
        for (int i = 0; i < N; i++) {
            cv::Mat f = synth::render(canvas, jitter.next());
            for (int y = 0; y < H; y++) memcpy(&frames[fb * i + (size_t)y * W * 3], f.ptr(y), (size_t)W * 3);
        }
        const auto combos = vstab::AlignerGridSearch::reference_grid();
        vstab::AlignerGridSearch gs(0, W, H, N, (int)combos.size());
        vstab::JitterScore in;
        const auto t0 = std::chrono::steady_clock::now();
        const auto res = gs.run(frames.data(), N, (int64_t)W * 3, (int64_t)fb, combos, &in);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("Input median jitter: %g px (%d pairs, %d not aligned)\n", in.median_px, in.pairs, in.failed);
        size_t best = 0;
        for (size_t i = 0; i < res.size(); i++) {
            printf("[%zu/%zu] PC=%d thr=%g frac=%g maxDisp=%g  outJit=%g  ratio=%g  failed=%d\n", i + 1, res.size(),
                   (int)res[i].combo.phase_correlate, res[i].combo.threshold, res[i].combo.smallest_fraction,
                   res[i].combo.max_displacement, res[i].out.median_px, res[i].ratio, res[i].failed_alignments);
            if (res[i].ratio < res[best].ratio) best = i;
        }
        printf("\nBest params: phase_correlate=%d  threshold=%g  smallest_fraction=%g  max_displacement=%g  ratio=%g\n",
               (int)res[best].combo.phase_correlate, res[best].combo.threshold, res[best].combo.smallest_fraction,
               res[best].combo.max_displacement, res[best].ratio);
        printf("%zu combinations x %d frames in %.3f s, %ld kernel launches\n", res.size(), N, s, gs.launches());
        return 0;
    } catch (const std::exception& e) {
        printf("[FAIL] exception: %s\n", e.what());
        return 2;
    }
}
