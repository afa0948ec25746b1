// trajectory.hpp — the host trajectory of the stabilizer (reference stabilizer.cpp:19-88) split into its two parts:
//
//   * the SMOOTHED value used at push n (stabilizer.cpp:35 -> smoother.cpp:74-127).  At push n the smoother finalises
//     measurement s = n - lagAhead from the raw window [max(0, s - lagBehind), s + lagAhead] with a fresh 100-iteration
//     relaxation: it depends on the raw measurements only, so the values of different pushes are independent and run on
//     several host threads (WorkerPool) or, for one video spread over several GPUs, on several ranks;
//   * the CHAIN (stabilizer.cpp:39-88): accumulate the jitter of the frame `lag` pushes back, decay, invert.  Strictly
//     sequential, a few dozen flops per frame.
//
// StabilizerTrajectory (stabilizer.hpp) is the streaming composition of the two; the batched and partitioned pipelines
// call them separately.  All three produce the same bits: one relaxation routine, one chain routine.
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "stabilizer.hpp"

namespace vstab {

// The local `smoothed` of VideoStabilizer::processFrame at push n (stabilizer.cpp:34-35): identity until the smoother
// has lagAhead later measurements, then the finalised value of measurement n - lagAhead.
SimilarityTransform smoothed_at_push(const SimilarityTransform* meas, long n, const VideoStabilizerParams& params);

// A few persistent host threads for the independent relaxations: parallel_for(n, f) calls f(i) for i in [0, n) on the
// calling thread and the workers and returns when all are done.  Waiting workers spin briefly before they sleep, so a
// burst of calls (one per chunk of a clip) costs microseconds, not thread wake-ups.
class WorkerPool {
public:
    explicit WorkerPool(int threads);      // threads <= 1: everything runs on the caller
    ~WorkerPool();
    WorkerPool(const WorkerPool&) = delete;
    WorkerPool& operator=(const WorkerPool&) = delete;
    int threads() const { return (int)m_workers.size() + 1; }
    void parallel_for(long n, const std::function<void(long)>& f);

private:
    void worker();
    std::vector<std::thread> m_workers;
    std::mutex m_mutex;
    std::condition_variable m_wake;
    const std::function<void(long)>* m_fn = nullptr;
    std::atomic<long> m_next{0}, m_n{0}, m_done{0};
    std::atomic<unsigned long> m_epoch{0};
    std::atomic<int> m_active{0};          // workers inside the current job
    bool m_stop = false;
};

// smoothed_at_push for pushes [n0, n1) into out[0 .. n1 - n0), on the pool (or the caller when pool is null)
void smooth_pushes(const SimilarityTransform* meas, long n0, long n1, const VideoStabilizerParams& params,
                   SimilarityTransform* out, WorkerPool* pool);

}  // namespace vstab
