// clip_stabilizer.hpp — batched form of VideoStabilizer for throughput: many frames per
// call, every stage one launch over the whole batch (north_star: "many frame pairs are
// batched per launch so the 2k-point sparse solves fill the SMs").
//
// feed(n frames) produces the frames n successive VideoStabilizer::processFrame calls would have
// produced, in the same order (the selections and iteration counts are identical; the solver's f64 sums
// run in the order of its CTA size, which follows the number of pairs in flight, so measured transforms
// agree to ~1e-9 px and the warped frames are the same bytes):
//   H2D (if host input) -> BGR->gray + pyramid for the n frames -> keyframe features for the
//   odd frames -> ONE solver launch over the n alignment pairs -> D2H of 4 doubles + status
//   per pair -> sequential host trajectory (L1 smoother, accumulate, decay) -> ONE warp launch
//   (crop fused) over the frames that became due -> D2H (if host output).
// Frames stay resident in a device ring of chunk_frames + lag + 1 slots, so a pair may span
// two feed() calls and delayed frames are warped without a second upload.
#pragma once

#include <stdint.h>

#include <vector>

#include "stabilizer.hpp"
#include "vstab.h"

namespace vstab {

class ClipStabilizer {
public:
    // Throws std::runtime_error when the device or its memory is unavailable.
    // nv12: the frames fed and produced are NV12 (VS_CLIP_NV12 in vstab.h: `height` rows of Y then height / 2 rows of
    // interleaved UV; 3/2 bytes per pixel, even width / height / crop) instead of interleaved BGR.
    ClipStabilizer(int device, int width, int height, int chunk_frames, const VideoStabilizerParams& params, bool nv12 = false);
    ~ClipStabilizer();
    ClipStabilizer(const ClipStabilizer&) = delete;
    ClipStabilizer& operator=(const ClipStabilizer&) = delete;

    // Start a new video: frame numbering (and with it the keyframe parity), the trajectory
    // and the delay line start over.
    void reset();

    // Feed the next n (<= chunk_frames) frames.  `frames`: interleaved BGR (or NV12 frames), row_stride /
    // frame_stride in bytes, in host (VS_MEM_HOST) or device (VS_MEM_DEVICE) memory.
    // Stabilized frames that became due are written densely ((w-2c) x (h-2c) x 3 each; NV12: x 3/2,
    // out_frame_stride bytes apart) to `out` in `out_mem`; returns how many (<= n).
    // With host input AND host output the call is software-pipelined over sub-chunks of
    // pipeline_frames(): the H2D copy of sub-chunk k+1 (copy-in stream) and the D2H copy of
    // the frames warped for sub-chunk k-1 (copy-out stream) overlap the kernels of sub-chunk k,
    // so a PCIe-fed clip runs at the speed of the slower PCIe direction rather than the sum of
    // both copies and the compute.  Pinned host buffers are needed for that overlap.
    int feed(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem,
             uint8_t* out, int64_t out_frame_stride, int out_mem);
    int pipeline_frames() const { return m_sub; }
    void set_pipeline_frames(int frames) { m_sub = frames < 1 ? 1 : frames; }
    // pieces a device-resident chunk of >= 128 pairs is cut into (one solver stream each); 1 = all stages back to back
    int solver_lanes() const { return m_lanes; }
    void set_solver_lanes(int lanes) { m_lanes = lanes < 1 ? 1 : lanes; }

    // Same, for frames that are already in the ring: slots are assigned in feed order,
    // frame f of the video lives in slot f % ring_capacity(); upload_only() places frames
    // there without processing them (lets a benchmark exclude H2D from its timed region).
    void upload_only(long first_frame, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem);
    int feed_resident(int n, uint8_t* out, int64_t out_frame_stride, int out_mem);

    int out_width() const { return m_w - 2 * m_crop; }
    int out_height() const { return m_h - 2 * m_crop; }
    size_t out_frame_bytes() const { return (size_t)out_width() * out_height() * 3 / (m_nv12 ? 2 : 1); }
    bool nv12() const { return m_nv12; }
    int ring_capacity() const { return m_capacity; }
    long frames_fed() const { return m_fed; }
    vs_ctx* context() const { return m_ctx; }
    vs_clip* clip() const { return m_clip; }

    // per-frame records of the last feed() (index i = i-th frame of that call)
    const std::vector<SimilarityTransform>& measurements() const { return m_meas; }
    const std::vector<uint8_t>& successes() const { return m_ok; }
    // corrections of the frames produced by the last feed(), in output order
    const std::vector<SimilarityTransform>& corrections() const { return m_corr; }

private:
    int m_w, m_h, m_chunk, m_capacity, m_crop;
    bool m_nv12 = false;
    int m_lanes = 3;       // device-resident chunks of >= 128 pairs: up to this many pieces on the clip's solver lanes
    int m_sub = 32;        // sub-chunk of the host-to-host pipeline
    int m_warp_batch = 96; // largest batch of due frames per warp launch while the host trajectory is still running (16, 32, 64, 96, 96 ...)
    VideoStabilizerParams m_params;
    vs_ctx* m_ctx = nullptr;
    vs_clip* m_clip = nullptr;
    StabilizerTrajectory m_trajectory;
    long m_fed = 0;        // frames fed since reset
    long m_emitted = 0;    // frames produced since reset
    std::vector<SimilarityTransform> m_meas, m_corr;
    std::vector<uint8_t> m_ok;
    std::vector<vs_pair> m_pairs;
    std::vector<double> m_T;
    std::vector<int32_t> m_status, m_slots;

    void check(int rc, const char* what) const;
    template <typename F> void for_slot_runs(long first_frame, int n, F f) const;
    int process(int n, uint8_t* out, int64_t out_frame_stride, int out_mem, bool append_records, bool async_to_host);
    int feed_pipelined(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, uint8_t* out, int64_t out_frame_stride);
};

}  // namespace vstab
