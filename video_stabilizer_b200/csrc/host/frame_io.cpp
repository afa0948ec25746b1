// frame_io.cpp — see frame_io.hpp.
#include "frame_io.hpp"

#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "trajectory.hpp"
#include "vstab.h"

namespace vstab {

namespace {

struct CopyThreads {
    std::mutex busy;
    WorkerPool pool;
    CopyThreads() : pool((int)std::max(1u, std::min(4u, std::thread::hardware_concurrency() / 2))) {}
};

CopyThreads& copy_threads()
{
    static CopyThreads* t = new CopyThreads();      // never destroyed: threads must not be joined during static teardown
    return *t;
}

struct PinnedPool {
    std::mutex m;
    std::multimap<size_t, void*> free_list;          // capacity -> buffer
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCached = (size_t)1 << 30;
};

PinnedPool& pinned_pool()
{
    static PinnedPool* p = new PinnedPool();         // intentionally leaked: outstanding frames may outlive main()
    return *p;
}

}  // namespace

void copy_rows_parallel(uint8_t* dst, size_t dst_stride, const uint8_t* src, size_t src_stride, size_t row_bytes, int rows)
{
    auto copy = [&](int r0, int r1) {
        if (dst_stride == row_bytes && src_stride == row_bytes) {
            memcpy(dst + (size_t)r0 * row_bytes, src + (size_t)r0 * row_bytes, (size_t)(r1 - r0) * row_bytes);
        } else {
            for (int r = r0; r < r1; r++) memcpy(dst + (size_t)r * dst_stride, src + (size_t)r * src_stride, row_bytes);
        }
    };
    CopyThreads& t = copy_threads();
    const int parts = t.pool.threads();
    // small frames, one copy thread, or the threads busy with another stabilizer's frame: copy here
    if (parts <= 1 || (size_t)rows * row_bytes < ((size_t)1 << 20) || !t.busy.try_lock()) {
        copy(0, rows);
        return;
    }
    std::lock_guard<std::mutex> lock(t.busy, std::adopt_lock);
    t.pool.parallel_for(parts, [&](long i) { copy((int)((long)rows * i / parts), (int)((long)rows * (i + 1) / parts)); });
}

std::shared_ptr<void> pinned_frame(size_t bytes)
{
    PinnedPool& p = pinned_pool();
    void* buf = nullptr;
    size_t cap = bytes;
    {
        std::lock_guard<std::mutex> lock(p.m);
        auto it = p.free_list.lower_bound(bytes);
        if (it != p.free_list.end() && it->first <= bytes + bytes / 4) {
            cap = it->first;
            buf = it->second;
            p.cached_bytes -= cap;
            p.free_list.erase(it);
        }
    }
    if (!buf && vs_pinned_alloc(bytes, &buf) != VS_OK)
        throw std::runtime_error(std::string("pinned_frame: ") + vs_last_error(nullptr));
    return std::shared_ptr<void>(buf, [cap](void* q) {
        PinnedPool& pool = pinned_pool();
        {
            std::lock_guard<std::mutex> lock(pool.m);
            if (pool.cached_bytes + cap <= PinnedPool::kMaxCached) {
                pool.free_list.emplace(cap, q);
                pool.cached_bytes += cap;
                return;
            }
        }
        vs_pinned_free(q);
    });
}

}  // namespace vstab
