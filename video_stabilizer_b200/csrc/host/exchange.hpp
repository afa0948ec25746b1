// exchange.hpp — the small per-frame table through which the workers of ONE video share their alignment results
// (north_star: "only the small per-frame transforms gathered to the host ... no NCCL collective on the hot path").
//
// A video is cut into sub-chunks of frames; every sub-chunk belongs to one worker (a rank = one process per GPU, or a
// thread of one process).  A worker publishes, per sub-chunk [a, b):
//   measurement[f], ok[f], f in [a, b)   AlignNextFrame's result for frame f: {A, B, TX, TY} and its return value
//   smoothed[n], n in [a, b)             the smoothed transform the stabilizer uses at push n (trajectory.hpp)
// each behind a release-store of a flag; readers acquire the flag and then read the rows.  The table lives in POSIX shared
// memory when the workers are processes (name = "/..." as for shm_open) and on the heap when they are threads of one
// process (empty name).  Successive videos (generations) alternate between two copies of the table, and a worker may only
// start generation g once every worker has finished generation g - 2, so a fast worker never overwrites rows a slow one is
// still reading.  Every wait spins with a deadline and throws std::runtime_error past it: a lost peer fails loudly.
#pragma once

#include <stdint.h>

#include <atomic>
#include <string>

#include "imgproc.hpp"

namespace vstab {

class TrajectoryExchange {
public:
    // create = true: this worker creates and initialises the table (exactly one worker, before the others attach).
    TrajectoryExchange(const std::string& shm_name, bool create, int rank, int world, long max_frames, int max_subchunks,
                       double timeout_seconds = 30.0);
    ~TrajectoryExchange();
    TrajectoryExchange(const TrajectoryExchange&) = delete;
    TrajectoryExchange& operator=(const TrajectoryExchange&) = delete;

    int rank() const { return m_rank; }
    int world() const { return m_world; }

    // start generation g (1, 2, 3 ... the same sequence on every worker): waits until all workers finished g - 2
    void begin(uint64_t generation);
    // this worker has read everything it needs of the current generation
    void finish();

    // tables of the current generation, indexed by frame / push
    SimilarityTransform* measurements() { return m_meas[m_gen & 1]; }
    uint8_t* ok() { return m_ok[m_gen & 1]; }
    SimilarityTransform* smoothed() { return m_sm[m_gen & 1]; }
    void publish_raw(int subchunk);
    void publish_smoothed(int subchunk);
    void wait_raw(int subchunk);
    void wait_smoothed(int subchunk);

private:
    struct Header;
    void wait_flag(const std::atomic<uint64_t>& flag, uint64_t want, const char* what, int index) const;
    std::string m_name;
    bool m_owner = false;
    int m_rank, m_world;
    long m_max_frames;
    int m_max_sub;
    double m_timeout;
    uint64_t m_gen = 0;
    void* m_base = nullptr;
    size_t m_bytes = 0;
    Header* m_hdr = nullptr;
    std::atomic<uint64_t>* m_done = nullptr;          // [world]
    std::atomic<uint64_t>* m_raw_flag[2] = {};        // [parity][max_subchunks]
    std::atomic<uint64_t>* m_sm_flag[2] = {};
    SimilarityTransform* m_meas[2] = {};
    uint8_t* m_ok[2] = {};
    SimilarityTransform* m_sm[2] = {};
};

}  // namespace vstab
