// exchange.cpp — see exchange.hpp.
#include "exchange.hpp"

#include <fcntl.h>
#include <immintrin.h>
#include <sched.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <new>
#include <stdexcept>

namespace vstab {

struct TrajectoryExchange::Header {
    std::atomic<uint64_t> magic;
    uint64_t world, max_frames, max_sub;
};

namespace {
constexpr uint64_t kMagic = 0x56535458434847ULL;   // "VSTXCHG"
size_t align64(size_t v) { return (v + 63) / 64 * 64; }
static_assert(sizeof(std::atomic<uint64_t>) == 8, "flags are plain 64-bit words in shared memory");
static_assert(sizeof(SimilarityTransform) == 32, "SimilarityTransform is four doubles");
}  // namespace

TrajectoryExchange::TrajectoryExchange(const std::string& shm_name, bool create, int rank, int world, long max_frames,
                                       int max_subchunks, double timeout_seconds)
    : m_name(shm_name), m_owner(create), m_rank(rank), m_world(world), m_max_frames(max_frames), m_max_sub(max_subchunks),
      m_timeout(timeout_seconds)
{
    if (world < 1 || rank < 0 || rank >= world || max_frames < 1 || max_subchunks < 1)
        throw std::runtime_error("TrajectoryExchange: bad arguments");
    const size_t o_done = align64(sizeof(Header));
    const size_t o_flags = o_done + align64(sizeof(uint64_t) * world);
    const size_t flag_bytes = align64(sizeof(uint64_t) * max_subchunks);
    const size_t tf_bytes = align64(sizeof(SimilarityTransform) * max_frames), ok_bytes = align64((size_t)max_frames);
    const size_t o_meas = o_flags + 4 * flag_bytes;
    const size_t o_sm = o_meas + 2 * tf_bytes;
    const size_t o_ok = o_sm + 2 * tf_bytes;
    m_bytes = o_ok + 2 * ok_bytes;

    if (m_name.empty()) {
        if (world != 1) throw std::runtime_error("TrajectoryExchange: several workers need a named shared segment");
        m_base = ::operator new(m_bytes, std::align_val_t(64));
        memset(m_base, 0, m_bytes);
    } else {
        int fd = -1;
        if (create) {
            shm_unlink(m_name.c_str());
            fd = shm_open(m_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)m_bytes) != 0) {
                if (fd >= 0) close(fd);
                throw std::runtime_error("TrajectoryExchange: cannot create shared segment " + m_name);
            }
        } else {
            // the creator may still be on its way: wait for the segment to exist at its full size
            const auto t0 = std::chrono::steady_clock::now();
            for (;;) {
                fd = shm_open(m_name.c_str(), O_RDWR, 0600);
                struct stat st;
                if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= m_bytes) break;
                if (fd >= 0) close(fd);
                fd = -1;
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > m_timeout)
                    throw std::runtime_error("TrajectoryExchange: shared segment " + m_name + " did not appear");
                usleep(1000);
            }
        }
        m_base = mmap(nullptr, m_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m_base == MAP_FAILED) {
            m_base = nullptr;
            throw std::runtime_error("TrajectoryExchange: mmap of " + m_name + " failed");
        }
    }
    char* b = (char*)m_base;
    m_hdr = (Header*)b;
    m_done = (std::atomic<uint64_t>*)(b + o_done);
    for (int p = 0; p < 2; p++) {
        m_raw_flag[p] = (std::atomic<uint64_t>*)(b + o_flags + (size_t)p * flag_bytes);
        m_sm_flag[p] = (std::atomic<uint64_t>*)(b + o_flags + (size_t)(2 + p) * flag_bytes);
        m_meas[p] = (SimilarityTransform*)(b + o_meas + (size_t)p * tf_bytes);
        m_sm[p] = (SimilarityTransform*)(b + o_sm + (size_t)p * tf_bytes);
        m_ok[p] = (uint8_t*)(b + o_ok + (size_t)p * ok_bytes);
    }
    if (create) {
        // a fresh segment is zero-filled: flags and `done` start at generation 0
        m_hdr->world = (uint64_t)world; m_hdr->max_frames = (uint64_t)max_frames; m_hdr->max_sub = (uint64_t)max_subchunks;
        m_hdr->magic.store(kMagic, std::memory_order_release);
    } else {
        wait_flag(m_hdr->magic, kMagic, "initialisation of the table", 0);
        if (m_hdr->world != (uint64_t)world || m_hdr->max_frames != (uint64_t)max_frames || m_hdr->max_sub != (uint64_t)max_subchunks)
            throw std::runtime_error("TrajectoryExchange: the workers disagree about the table's geometry");
    }
}

TrajectoryExchange::~TrajectoryExchange()
{
    if (!m_base) return;
    if (m_name.empty()) {
        ::operator delete(m_base, std::align_val_t(64));
    } else {
        munmap(m_base, m_bytes);
        if (m_owner) shm_unlink(m_name.c_str());
    }
}

void TrajectoryExchange::wait_flag(const std::atomic<uint64_t>& flag, uint64_t want, const char* what, int index) const
{
    if (flag.load(std::memory_order_acquire) >= want) return;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; spins++) {
        if (flag.load(std::memory_order_acquire) >= want) return;
        if ((spins & 63) == 63) {
            if ((spins & 0xffff) == 0xffff &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > m_timeout)
                throw std::runtime_error(std::string("TrajectoryExchange: rank ") + std::to_string(m_rank) + " timed out waiting for " +
                                         what + " " + std::to_string(index) + " (a peer is gone or behind by more than the deadline)");
            sched_yield();      // the peers may share this core (tests run several workers on few cores)
        } else {
            _mm_pause();
        }
    }
}

void TrajectoryExchange::begin(uint64_t generation)
{
    if (generation >= 2)
        for (int q = 0; q < m_world; q++) wait_flag(m_done[q], generation - 2, "the end of an earlier video on rank", q);
    m_gen = generation;
}

void TrajectoryExchange::finish() { m_done[m_rank].store(m_gen, std::memory_order_release); }

void TrajectoryExchange::publish_raw(int j) { m_raw_flag[m_gen & 1][j].store(m_gen, std::memory_order_release); }
void TrajectoryExchange::publish_smoothed(int j) { m_sm_flag[m_gen & 1][j].store(m_gen, std::memory_order_release); }
void TrajectoryExchange::wait_raw(int j) { wait_flag(m_raw_flag[m_gen & 1][j], m_gen, "the measurements of sub-chunk", j); }
void TrajectoryExchange::wait_smoothed(int j) { wait_flag(m_sm_flag[m_gen & 1][j], m_gen, "the smoothed transforms of sub-chunk", j); }

}  // namespace vstab
