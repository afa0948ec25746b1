// partitioned.hpp — one worker's share of ONE video that is partitioned by frame chunk over several GPUs
// (BASELINE.json north_star: "frame pairs and warps are independent, so work is partitioned by frame-chunk across the 8
// GPUs of one box, with per-GPU streams and only the small per-frame transforms gathered to the host").
//
// The video is cut into sub-chunks of `sub_frames` frames (even, so every sub-chunk starts on an even frame: keyframes are
// the odd frames of the VIDEO, alignment.cpp:357); `block` consecutive sub-chunks form a chunk and chunks go round-robin to
// the workers.  block * sub_frames = frames / workers gives every GPU one contiguous chunk (the north_star layout, used for
// device-resident runs: the sub-chunks inside it only pipeline the stages); block = 1 with small sub-chunks interleaves the
// workers along the video, which is what a host-fed stream wants (every GPU then uploads, computes and downloads all the
// time).  A worker holds its own frames plus one halo frame (the last frame of a foreign predecessor sub-chunk).
//
// Per own sub-chunk [a, b), on the worker's GPU: pyramids, keyframe features of the odd frames, ONE solver launch for the
// pairs (f-1 -> f), f in [a, b), on a solver lane of its own (the next sub-chunk's pyramids run beside it).  On the host
// (PartitionedTrajectory): publish the measurements (TrajectoryExchange), compute the smoothed transforms of pushes [a, b)
// on the worker pool and publish them, advance the sequential chain (every worker walks the whole chain: tens of
// nanoseconds per frame) and warp the own frames that became due.  There is no collective and no device-to-device
// traffic; what crosses between workers is 65 bytes per frame of host memory.  The output equals VideoStabilizer fed
// frame by frame.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "exchange.hpp"
#include "stabilizer.hpp"
#include "trajectory.hpp"
#include "vstab.h"

namespace vstab {

struct VideoPartition {
    long total = 0;
    int world = 1, sub = 2, block = 1;
    VideoPartition() {}
    VideoPartition(long total_frames, int workers, int sub_frames, int block_subchunks);
    int count() const { return (int)((total + sub - 1) / sub); }
    long first(int j) const { return (long)j * sub; }
    long last(int j) const { return std::min(total, (long)(j + 1) * sub); }
    int owner(int j) const { return (j / block) % world; }
    int subchunk_of(long frame) const { return (int)(frame / sub); }
};

// The host half of a worker: which frames it owns, and the distributed trajectory (no GPU involved).
class PartitionedTrajectory {
public:
    struct Sub {
        int j;              // sub-chunk index in the video
        long a, b;          // frames [a, b)
        int local0;         // first local entry (the halo when it has one)
        int halo;           // 1 when local0 is a halo frame
        int own0;           // own-frame index of frame a
    };
    // exchange_name: POSIX shared-memory name ("/something") shared by the workers of the video; may be empty when
    // world == 1.  Rank 0 creates the segment (the others wait for it to appear).
    PartitionedTrajectory(int rank, int world, int width, int height, long total_frames, int sub_frames, int block_subchunks,
                          const VideoStabilizerParams& params, const std::string& exchange_name, int host_threads);

    const VideoPartition& partition() const { return m_plan; }
    const std::vector<Sub>& subs() const { return m_subs; }
    // local frame list: the worker's own frames in video order, each foreign-preceded run headed by its halo frame
    int local_count() const { return (int)m_local.size(); }
    long local_frame(int i) const { return m_local[i]; }
    bool local_is_halo(int i) const { return m_halo[i] != 0; }
    int local_of_own(int own_index) const;
    int own_count() const { return m_own_frames; }
    int output_count() const { return m_outputs; }      // own frames f with f + lag < total

    // every worker: begin_video, then submit() for each own sub-chunk in order, then flush, then end_video
    void begin_video();
    // AlignNextFrame results of the pairs (f-1 -> f), f in [max(a,1), b) of own sub-chunk i: T = 4 doubles per pair.
    // Publishes them, computes and publishes the smoothed transforms of pushes [a, b), advances the chain to push b - 1.
    void submit(int i, const double* T, const int32_t* status);
    void flush();           // advance the chain until the last own output is decided
    void end_video();
    // corrections decided so far: own outputs [0, due()) in own-frame order
    int due() const { return m_due; }
    const SimilarityTransform& correction(int own_index) const { return m_corr[own_index]; }
    // the part of the video's tables this worker has seen (valid between begin_video and end_video; copies survive after)
    long seen_frames() const { return m_chain_pos; }
    const std::vector<SimilarityTransform>& measurements() const { return m_meas_copy; }
    const std::vector<uint8_t>& successes() const { return m_ok_copy; }
    const std::vector<SimilarityTransform>& corrections() const { return m_corr; }

private:
    int m_w, m_h, m_rank;
    VideoStabilizerParams m_params;
    VideoPartition m_plan;
    std::unique_ptr<TrajectoryExchange> m_xchg;
    std::unique_ptr<WorkerPool> m_pool;
    std::vector<Sub> m_subs;
    std::vector<long> m_local;
    std::vector<uint8_t> m_halo;
    int m_own_frames = 0, m_outputs = 0;
    uint64_t m_generation = 0;
    TrajectoryChain m_chain;
    long m_chain_pos = 0;
    int m_sm_waited = -1;               // highest sub-chunk whose smoothed transforms are known to be published
    int m_due = 0;
    std::vector<SimilarityTransform> m_corr, m_meas_copy;
    std::vector<uint8_t> m_ok_copy, m_corr_ready;

    const Sub* own_sub_of(long frame) const;
    void advance_chain(long upto);
};

class PartitionedStabilizer {
public:
    // resident = true: the ring holds all of this worker's frames (upload_resident + stabilize(nullptr ...));
    // false: frames stream through a ring of (lanes + 1) sub-chunks.
    // lanes: sub-chunks in flight on the GPU, one solver stream each (0 = default: the sub-chunks of a chunk, at most
    // VS_CLIP_SOLVER_LANES, when resident; 2 when streamed)
    // nv12: the video's frames, in and out, are NV12 (VS_CLIP_NV12 in vstab.h) instead of interleaved BGR
    PartitionedStabilizer(int device, int rank, int world, int width, int height, long total_frames, int sub_frames,
                          int block_subchunks, const VideoStabilizerParams& params, const std::string& exchange_name,
                          bool resident, int host_threads = 4, int lanes = 0, bool nv12 = false);
    ~PartitionedStabilizer();
    PartitionedStabilizer(const PartitionedStabilizer&) = delete;
    PartitionedStabilizer& operator=(const PartitionedStabilizer&) = delete;

    const PartitionedTrajectory& trajectory() const { return m_traj; }
    // place the local frames in the ring without processing them (resident mode): `frames` holds local_count() frames
    void upload_resident(const uint8_t* frames, int64_t row_stride, int64_t frame_stride, int mem);
    // Stabilize the video (every worker calls this once per video, in the same order).  frames: the worker's local frames
    // (host memory; nullptr in resident mode).  out: output_count() dense frames in own-frame order, `out_frame_stride`
    // bytes apart, in out_mem.  Returns output_count().
    int stabilize(const uint8_t* frames, int64_t row_stride, int64_t frame_stride, uint8_t* out, int64_t out_frame_stride, int out_mem);

    // sub-chunks in flight on the GPU (one solver lane each), 1 .. lanes the ring was built for; 1 = every stage back to back
    void set_lanes(int lanes) { m_lanes = std::max(1, std::min(lanes, m_max_lanes)); }
    int lanes() const { return m_lanes; }
    int out_width() const { return m_w - 2 * m_crop; }
    int out_height() const { return m_h - 2 * m_crop; }
    size_t out_frame_bytes() const { return (size_t)out_width() * out_height() * 3 / (m_nv12 ? 2 : 1); }
    vs_ctx* context() const { return m_ctx; }
    vs_clip* clip() const { return m_clip; }

private:
    int m_w, m_h, m_crop;
    bool m_resident;
    bool m_nv12 = false;
    int m_lanes = 3, m_max_lanes = 3;
    VideoStabilizerParams m_params;
    PartitionedTrajectory m_traj;
    vs_ctx* m_ctx = nullptr;
    vs_clip* m_clip = nullptr;
    int m_capacity = 0;
    int m_emitted = 0;                  // own frames warped so far (this video)
    std::vector<vs_pair> m_pairs;
    std::vector<double> m_T;
    std::vector<int32_t> m_status, m_slots;

    void check(int rc, const char* what) const;
    int slot_of_local(int i) const { return i % m_capacity; }
    template <typename F> void for_slot_runs(int local0, int n, F f) const;
    void issue(int i, const uint8_t* frames, int64_t row_stride, int64_t frame_stride);
    void warp_due(uint8_t* out, int64_t out_frame_stride, int out_mem, bool async_to_host);
};

}  // namespace vstab
