// alignment.hpp — VideoAligner, drop-in for the reference's alignment.hpp (same
// VideoAlignerParams fields and defaults, same AlignNextFrame signature and return rule;
// reference alignment.hpp:5-41, :51-58).  The reference's protected host state
// (alignment.hpp:60-98: pyramids, gradients, keypoints, Jacobians, selections) lives on the
// GPU here, behind an opaque handle.
#pragma once

#include <memory>

#include "imgproc.hpp"

struct VideoAlignerParams {
    // Initialise from cv::phaseCorrelate on pyramid level 2 (alignment.cpp:369-388).  Off by default
    // upstream; here it runs on the device (vs_phasecorr.cu).
    bool phase_correlate = false;
    double phase_correlate_threshold = 0.5;

    // Stop a level's Gauss-Newton loop when no frame corner moved more than this (pixels).
    double threshold = 0.02;

    // Fraction of each level's keypoints kept (those with the smallest warp residual).
    float smallest_fraction = 0.8f;

    // Iterations per pyramid level before the pair is declared lost.
    int max_iters = 64;

    // Smallest pyramid level (the pyramid stops before a level would be narrower/shorter).
    int pyramid_min_width = 20;
    int pyramid_min_height = 20;

    // Largest converged corner displacement accepted at any level (pixels at that level).
    double max_displacement = 10.0;
};

// Aligns each frame against its predecessor.  Every other frame (the 2nd, 4th, ... since the
// last size change) is a keyframe whose gradient keypoints serve the pair before and the
// pair after it; the transform is solved coarse to fine by sparse inverse-compositional
// Lucas-Kanade, entirely on the device.
class VideoAligner {
public:
    VideoAligner();
    ~VideoAligner();
    VideoAligner(VideoAligner&&) noexcept;
    VideoAligner& operator=(VideoAligner&&) noexcept;

    // Returns false for the first frame (and the first after a size change), when the solve
    // does not converge, when it moves further than max_displacement, or on a device error.
    // `transform` maps the previous frame to this one (centre-based TX,TY); on false it holds
    // whatever the solve had reached (identity for a first frame).
    bool AlignNextFrame(const cv::Mat& frame, SimilarityTransform& transform,
                        const VideoAlignerParams& params = VideoAlignerParams());

protected:
    friend class VideoStabilizer;
    struct Impl;
    std::unique_ptr<Impl> impl_;
};
