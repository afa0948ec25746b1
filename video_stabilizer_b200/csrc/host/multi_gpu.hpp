// multi_gpu.hpp — one video, several GPUs of one box, one process (north_star: "work is
// partitioned by frame-chunk across the GPUs, with per-GPU streams and only the small per-frame
// transforms gathered to the host; the sequential L1 smoother stays host C++; no collective").
//
//   phase 1 (one host thread per GPU, concurrently): upload the GPU's frame chunk (+ the one
//            halo frame before it), BGR->gray pyramids, keyframe features of the odd frames
//            (frame parity is global: chunks start on even frames), ONE solver launch over the
//            chunk's pairs; 40 bytes per pair land in a shared host table
//   phase 2 (host, sequential): StabilizerTrajectory over the whole table
//   phase 3 (per GPU, concurrently): warp (crop fused) the chunk's frames that became due, from
//            the frames still resident on that GPU, and copy them to their place in the output
// The result equals feeding the whole video to one VideoStabilizer frame by frame.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "stabilizer.hpp"
#include "vstab.h"

namespace vstab {

class MultiGpuStabilizer {
public:
    // devices: CUDA device ordinals, one worker each (the same ordinal may appear twice: two
    // workers then share a GPU, which is how the path is tested on a single device).
    // max_frames: longest video stabilize() will be given.
    MultiGpuStabilizer(const std::vector<int>& devices, int width, int height, int max_frames, const VideoStabilizerParams& params);
    ~MultiGpuStabilizer();
    MultiGpuStabilizer(const MultiGpuStabilizer&) = delete;
    MultiGpuStabilizer& operator=(const MultiGpuStabilizer&) = delete;

    // frames: n interleaved BGR host frames (row_stride / frame_stride bytes); out: n - lag dense
    // stabilized frames ((w-2c) x (h-2c) x 3, out_frame_stride bytes apart, host).  Returns the
    // number of frames written.  Throws std::runtime_error on a device error.
    int stabilize(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, uint8_t* out, int64_t out_frame_stride);

    // [(first, last_exclusive)] per worker for an n-frame video: contiguous, even-aligned starts
    static std::vector<std::pair<int, int>> frame_chunks(int n, int workers);

    int out_width() const { return m_w - 2 * m_crop; }
    int out_height() const { return m_h - 2 * m_crop; }
    const std::vector<SimilarityTransform>& measurements() const { return m_meas; }
    const std::vector<uint8_t>& successes() const { return m_ok; }

private:
    struct Worker {
        int device = 0;
        vs_ctx* ctx = nullptr;
        vs_clip* clip = nullptr;
        int capacity = 0;
        std::string error;
    };
    int m_w, m_h, m_crop, m_max_frames;
    VideoStabilizerParams m_params;
    std::vector<Worker> m_workers;
    std::vector<SimilarityTransform> m_meas;
    std::vector<uint8_t> m_ok;
    void release();
};

}  // namespace vstab
