// grid_search.hpp — quality tooling on the batched API.
//
// The reference tunes VideoAlignerParams with grid_search_align.cpp: for each of 54 parameter combinations (:134-146) a
// worker thread runs a whole VideoStabilizer over the clip (smoother off, lag 1; :167-180) and scores the output by its
// jitter, the median optical-flow magnitude between consecutive frames (measure_jitter, :27-60; eval_jitter.cpp:21-75).
// Here the same experiment is a handful of launches: every pair of the clip is aligned under ALL combinations in one
// solver launch (vs_clip_align_sweep), the 54 trajectories run on the host, the 54 x (n-1) output frames are warped in
// batched launches, and the jitter of every output sequence comes from one more batched alignment.
//
// Jitter here is the median flow magnitude of the global motion the aligner recovers between two frames, sampled on a
// grid of points (the reference estimates the same statistic with dense Farneback flow, which needs OpenCV's video
// module: not on the hot path, not rebuilt).  Pairs the aligner cannot align do not contribute.
#pragma once

#include <stdint.h>

#include <vector>

#include "stabilizer.hpp"
#include "vstab.h"

namespace vstab {

// median over a 32 x 18 grid of |W(p) - p| for a centre-based similarity on a w x h frame
double flow_median_px(const SimilarityTransform& T, int w, int h);

struct JitterScore {
    double median_px = 0;     // median over the aligned pairs of flow_median_px
    int pairs = 0, failed = 0;
};

class AlignerGridSearch {
public:
    struct Combo {
        bool phase_correlate;
        double threshold;
        float smallest_fraction;
        double max_displacement;
    };
    struct Result {
        Combo combo;
        JitterScore out;
        double ratio = 0;          // output jitter / input jitter (grid_search_align.cpp:187)
        int failed_alignments = 0; // pairs of the input clip this combination could not align
    };
    // the reference's grid: 2 x 3 x 3 x 3 = 54 combinations, in its loop order (grid_search_align.cpp:134-146)
    static std::vector<Combo> reference_grid();

    AlignerGridSearch(int device, int width, int height, int max_frames, int max_combos, int crop_pixels = 32);
    ~AlignerGridSearch();
    AlignerGridSearch(const AlignerGridSearch&) = delete;
    AlignerGridSearch& operator=(const AlignerGridSearch&) = delete;

    // jitter of n host BGR frames (default aligner parameters)
    JitterScore measure_jitter(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride);
    // the sweep; input_jitter (optional) receives the clip's own score
    std::vector<Result> run(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride,
                            const std::vector<Combo>& combos, JitterScore* input_jitter);
    // measurements of the last run(): [combo][pair] (pair i = frame i+1 against frame i)
    const std::vector<double>& sweep_transforms() const { return m_T; }
    const std::vector<int32_t>& sweep_status() const { return m_status; }
    long launches() const;

private:
    int m_w, m_h, m_crop, m_max_frames, m_max_combos, m_group;
    vs_ctx* m_ctx = nullptr;
    vs_clip* m_in = nullptr;       // the input clip
    vs_clip* m_out = nullptr;      // stabilized sequences of a group of combinations
    uint8_t* m_dev_out = nullptr;  // warped frames of a group before they enter m_out
    std::vector<double> m_T;
    std::vector<int32_t> m_status;
    void check(int rc, const char* what) const;
    JitterScore score(vs_clip* clip, int first_slot, int n, int w, int h, int sequences, int seq_stride, std::vector<JitterScore>* per_seq);
};

}  // namespace vstab
