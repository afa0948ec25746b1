// trajectory.cpp — see trajectory.hpp.
#include "trajectory.hpp"

#include <immintrin.h>

#include <algorithm>

namespace vstab {

SimilarityTransform smoothed_at_push(const SimilarityTransform* meas, long n, const VideoStabilizerParams& params)
{
    // upstream passes (lag, smoother_memory) as (lagBehind, lagAhead): reference stabilizer.cpp:4
    if (!params.enable_smoother || n < params.smoother_memory) return SimilarityTransform();
    return smoother_finalize(meas, n - params.smoother_memory, params.lag, params.smoother_memory, params.lambda);
}

bool TrajectoryChain::step(long n, bool success, const SimilarityTransform& oldest, const SimilarityTransform& smoothed,
                           int frame_width, int frame_height, SimilarityTransform& correction)
{
    if (!success) m_accum = SimilarityTransform();          // stabilizer.cpp:39-41
    if (n < m_params.lag) return false;                     // the delay line is still filling (stabilizer.cpp:48)

    const SimilarityTransform jitter = m_params.enable_smoother ? oldest.compose(smoothed.inverse()) : oldest;
    SimilarityTransform next = m_accum.compose(jitter);

    const double displacement = next.maxCornerDisplacement(frame_width, frame_height);
    double decay;
    if (displacement > m_params.max_disp) {
        decay = m_params.max_decay;
    } else if (displacement > m_params.min_disp) {
        double f = (displacement - m_params.min_disp) / (m_params.max_disp - m_params.min_disp);
        f = std::max(0.0, std::min(1.0, f));
        decay = m_params.min_decay * (1.0 - f) + m_params.max_decay * f;
    } else {
        decay = m_params.min_decay;
    }
    next.TX *= decay;
    next.TY *= decay;
    next.A *= decay;
    next.B *= decay;
    m_accum = next;
    correction = next.inverse();
    return true;
}

// ---------------------------------------------------------------- worker pool
WorkerPool::WorkerPool(int threads)
{
    for (int i = 1; i < threads; i++) m_workers.emplace_back([this] { worker(); });
}

WorkerPool::~WorkerPool()
{
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        m_stop = true;
        m_epoch.fetch_add(1, std::memory_order_release);
    }
    m_wake.notify_all();
    for (auto& t : m_workers) t.join();
}

void WorkerPool::worker()
{
    unsigned long seen = 0;
    for (;;) {
        // spin briefly for the next job, then sleep
        unsigned long e = m_epoch.load(std::memory_order_acquire);
        for (int spin = 0; e == seen && spin < 20000; spin++) {
            _mm_pause();
            e = m_epoch.load(std::memory_order_acquire);
        }
        if (e == seen) {
            std::unique_lock<std::mutex> lock(m_mutex);
            m_wake.wait(lock, [&] { return m_epoch.load(std::memory_order_acquire) != seen; });
            e = m_epoch.load(std::memory_order_acquire);
        }
        seen = e;
        const std::function<void(long)>* fn;
        long n;
        {
            std::lock_guard<std::mutex> lock(m_mutex);      // the job's fields are published and retired under the mutex
            if (m_stop) return;
            fn = m_fn;
            if (!fn) continue;                              // woke up after the job had ended
            n = m_n.load(std::memory_order_acquire);
            m_active.fetch_add(1, std::memory_order_acq_rel);
        }
        long done = 0;
        for (long i = m_next.fetch_add(1, std::memory_order_acq_rel); i < n; i = m_next.fetch_add(1, std::memory_order_acq_rel)) {
            (*fn)(i);
            done++;
        }
        m_done.fetch_add(done, std::memory_order_acq_rel);
        m_active.fetch_sub(1, std::memory_order_acq_rel);
    }
}

void WorkerPool::parallel_for(long n, const std::function<void(long)>& f)
{
    if (n <= 0) return;
    if (m_workers.empty() || n == 1) {
        for (long i = 0; i < n; i++) f(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        m_fn = &f;
        m_n.store(n, std::memory_order_release);
        m_next.store(0, std::memory_order_release);
        m_done.store(0, std::memory_order_release);
        m_epoch.fetch_add(1, std::memory_order_release);
    }
    m_wake.notify_all();
    long done = 0;
    for (long i = m_next.fetch_add(1, std::memory_order_acq_rel); i < n; i = m_next.fetch_add(1, std::memory_order_acq_rel)) {
        f(i);
        done++;
    }
    m_done.fetch_add(done, std::memory_order_acq_rel);
    // every item is finished and no worker still holds the job (f goes out of scope with the caller)
    while (m_done.load(std::memory_order_acquire) < n) _mm_pause();
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        m_fn = nullptr;
    }
    while (m_active.load(std::memory_order_acquire) != 0) _mm_pause();
}

void smooth_pushes(const SimilarityTransform* meas, long n0, long n1, const VideoStabilizerParams& params,
                   SimilarityTransform* out, WorkerPool* pool)
{
    const long count = n1 - n0;
    if (count <= 0) return;
    if (!pool || pool->threads() <= 1 || count < 8) {
        for (long i = 0; i < count; i++) out[i] = smoothed_at_push(meas, n0 + i, params);
        return;
    }
    // blocks of a few pushes per work item: an item is ~2 us of arithmetic per push
    const long block = 8, items = (count + block - 1) / block;
    pool->parallel_for(items, [&](long it) {
        const long a = it * block, b = std::min(count, a + block);
        for (long i = a; i < b; i++) out[i] = smoothed_at_push(meas, n0 + i, params);
    });
}

}  // namespace vstab
