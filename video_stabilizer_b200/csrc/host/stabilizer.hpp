// stabilizer.hpp — VideoStabilizer, drop-in for the reference's stabilizer.hpp (same
// VideoStabilizerParams fields and defaults, same constructor and processFrame;
// reference stabilizer.hpp:13-39).
#pragma once

#include <deque>

#include "alignment.hpp"
#include "smoother.hpp"

struct VideoStabilizerParams {
    VideoAlignerParams aligner;

    int lag = 10;              // frames of delay between input and stabilized output
    int smoother_memory = 5;   // handed to the smoother as its look-ahead (see StabilizerTrajectory)
    double lambda = 4.0;       // smoothing strength of the L1 smoother

    bool enable_smoother = true;   // false: remove the raw measured motion instead

    int crop_pixels = 32;      // border removed from every output frame

    // accumulated-correction decay: min_decay below min_disp px of corner displacement,
    // max_decay above max_disp, linear in between
    double min_disp = 48.0, max_disp = 64.0;
    double min_decay = 0.9, max_decay = 0.7;
};

namespace vstab {

// The sequential part of the trajectory for one push (reference stabilizer.cpp:39-88): reset on failure, accumulate the
// jitter of the frame `lag` pushes back, decay, invert.  See trajectory.hpp for the smoothed value it takes.
class TrajectoryChain {
public:
    explicit TrajectoryChain(const VideoStabilizerParams& params) : m_params(params) {}
    // push n: `success` is AlignNextFrame's return value for frame n; when n >= lag, `oldest` is the measurement of frame
    // n - lag and `smoothed` the smoother's output at this push (identity while it has none); returns true and the
    // correction of frame n - lag then.
    bool step(long n, bool success, const SimilarityTransform& oldest, const SimilarityTransform& smoothed,
              int frame_width, int frame_height, SimilarityTransform& correction);
    const SimilarityTransform& accumulated() const { return m_accum; }

private:
    VideoStabilizerParams m_params;
    SimilarityTransform m_accum;
};

// The sequential, host-side half of the stabilizer (reference stabilizer.cpp:19-88):
// measurement in, correction for the frame `lag` frames back out.  Shared by
// VideoStabilizer::processFrame and the batched clip pipeline so both produce the same
// trajectory.
class StabilizerTrajectory {
public:
    explicit StabilizerTrajectory(const VideoStabilizerParams& params);
    // Feed the measurement of the newest frame (success = AlignNextFrame's return value).
    // Returns true when the frame `lag` frames back is due; `correction` is then the
    // transform to warp it by (inverse of the decayed accumulated jitter).
    bool push(const SimilarityTransform& measurement, bool success, int frame_width, int frame_height,
              SimilarityTransform& correction);
    const SimilarityTransform& accumulated() const { return m_chain.accumulated(); }

private:
    VideoStabilizerParams m_params;
    L1SmootherCenter m_smoother;
    TrajectoryChain m_chain;
    std::deque<SimilarityTransform> m_measurements;
    long m_pushes = 0;
};

}  // namespace vstab

class VideoStabilizer {
public:
    VideoStabilizer(const VideoStabilizerParams& params = VideoStabilizerParams());
    ~VideoStabilizer();

    // Feed one BGR frame.  Returns the stabilized (and cropped) frame from `lag` frames ago,
    // or an empty cv::Mat while the delay line is still filling.
    cv::Mat processFrame(const cv::Mat& inputFrame);

protected:
    struct Pending {
        int slot = -1;        // ring slot on the GPU, valid while generation matches
        int generation = -1;
        cv::Mat host;         // only filled when the ring had to be dropped (frame size changed)
    };
    VideoStabilizerParams m_params;
    VideoAligner aligner;
    int m_frameIndex = 0;
    vstab::StabilizerTrajectory m_trajectory;
    std::deque<Pending> m_frameBuffer;
};
