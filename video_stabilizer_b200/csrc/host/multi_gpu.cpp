// multi_gpu.cpp — see multi_gpu.hpp.
#include "multi_gpu.hpp"

#include <algorithm>
#include <stdexcept>
#include <thread>

#include "aligner_impl.hpp"

namespace vstab {

std::vector<std::pair<int, int>> MultiGpuStabilizer::frame_chunks(int n, int workers)
{
    std::vector<int> bounds(1, 0);
    for (int r = 1; r < workers; r++) {
        int b = (int)((long long)n * r / workers) & ~1;     // even boundary: keyframes are the odd frames of the video
        bounds.push_back(std::max(b, bounds.back()));
    }
    bounds.push_back(n);
    std::vector<std::pair<int, int>> chunks;
    for (int r = 0; r < workers; r++) chunks.emplace_back(bounds[r], bounds[r + 1]);
    return chunks;
}

MultiGpuStabilizer::MultiGpuStabilizer(const std::vector<int>& devices, int width, int height, int max_frames,
                                       const VideoStabilizerParams& params)
    : m_w(width), m_h(height), m_crop(std::max(0, params.crop_pixels)), m_max_frames(max_frames), m_params(params)
{
    if (devices.empty() || width <= 0 || height <= 0 || max_frames <= 0) throw std::runtime_error("MultiGpuStabilizer: bad arguments");
    if (2 * m_crop >= width || 2 * m_crop >= height) throw std::runtime_error("MultiGpuStabilizer: crop_pixels removes the whole frame");
    vs_align_params cp;
    to_c_params(params.aligner, &cp);
    const int workers = (int)devices.size();
    // a chunk holds at most ceil(n / workers) + 2 frames (even alignment) plus the halo frame
    const int capacity = (max_frames + workers - 1) / workers + 3;
    m_workers.resize(workers);
    try {
        for (int i = 0; i < workers; i++) {
            Worker& wk = m_workers[i];
            wk.device = devices[i];
            wk.capacity = capacity;
            if (vs_ctx_create(wk.device, &wk.ctx) != VS_OK)
                throw std::runtime_error(std::string("MultiGpuStabilizer: cannot create a GPU context: ") + vs_last_error(nullptr));
            if (vs_clip_create(wk.ctx, width, height, capacity, capacity, &cp, 0, &wk.clip) != VS_OK)
                throw std::runtime_error(std::string("MultiGpuStabilizer: ") + vs_last_error(wk.ctx));
        }
    } catch (...) {
        release();      // the destructor does not run for a partially constructed object
        throw;
    }
}

void MultiGpuStabilizer::release()
{
    for (Worker& wk : m_workers) {
        if (wk.clip) vs_clip_destroy(wk.clip);
        if (wk.ctx) vs_ctx_destroy(wk.ctx);
        wk.clip = nullptr;
        wk.ctx = nullptr;
    }
}

MultiGpuStabilizer::~MultiGpuStabilizer() { release(); }

int MultiGpuStabilizer::stabilize(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, uint8_t* out,
                                  int64_t out_frame_stride)
{
    if (n < 0 || n > m_max_frames) throw std::runtime_error("MultiGpuStabilizer: more frames than max_frames");
    m_meas.assign(n, SimilarityTransform());
    m_ok.assign(n, 0);
    if (n == 0) return 0;
    const int workers = (int)m_workers.size();
    const auto chunks = frame_chunks(n, workers);
    std::vector<int> up_first(workers);      // first frame resident on each worker (its chunk start, or the halo frame before it)

    auto run_all = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int r = 0; r < workers; r++) th.emplace_back([&, r] {
            try {
                fn(r);
            } catch (const std::exception& e) {
                m_workers[r].error = e.what();
            }
        });
        for (auto& t : th) t.join();
        for (Worker& wk : m_workers)
            if (!wk.error.empty()) {
                std::string msg = wk.error;
                wk.error.clear();
                throw std::runtime_error(msg);
            }
    };
    auto check = [&](Worker& wk, int rc, const char* what) {
        if (rc != VS_OK) throw std::runtime_error(std::string("MultiGpuStabilizer: ") + what + ": " + vs_last_error(wk.ctx));
    };

    // ---- phase 1: per-GPU ingest and alignment of the chunk's pairs
    run_all([&](int r) {
        Worker& wk = m_workers[r];
        const int first = chunks[r].first, last = chunks[r].second;
        up_first[r] = first;
        if (last <= first) return;
        const int up0 = (r > 0 && first > 0) ? first - 1 : first;
        up_first[r] = up0;
        const int cnt = last - up0;
        if (cnt > wk.capacity) throw std::runtime_error("MultiGpuStabilizer: chunk larger than the worker's ring");
        // asynchronous copies on the clip's copy-in stream (PCIe rate when the caller's frames are pinned); the pyramids of
        // a sub-chunk are built while the next one is still arriving
        const int sub = 32;
        auto upload = [&](int s0) {
            check(wk, vs_clip_upload_async(wk.clip, s0, std::min(sub, cnt - s0), frames + (size_t)frame_stride * (up0 + s0),
                                           row_stride, frame_stride), "upload");
        };
        upload(0);
        for (int s0 = 0; s0 < cnt; s0 += sub) {
            check(wk, vs_clip_wait_uploads(wk.clip), "wait for uploads");     // covers sub-chunk s0: the next one is not issued yet
            if (s0 + sub < cnt) upload(s0 + sub);
            check(wk, vs_clip_build_pyramids(wk.clip, s0, std::min(sub, cnt - s0)), "pyramids");
        }
        std::vector<int32_t> keys;
        std::vector<vs_pair> pairs;
        for (int f = std::max(first, 1); f < last; f++) {
            vs_pair p;
            if (f & 1) { p.template_slot = f - 1 - up0; p.keyframe_slot = f - up0; p.invert = 0; }
            else       { p.template_slot = f - up0; p.keyframe_slot = f - 1 - up0; p.invert = 1; }
            pairs.push_back(p);
        }
        for (int f = up0; f < last; f++)
            if (f & 1) keys.push_back(f - up0);
        if (!keys.empty()) check(wk, vs_clip_build_keyframes(wk.clip, keys.data(), (int)keys.size()), "keyframes");
        if (pairs.empty()) return;
        std::vector<double> T(pairs.size() * 4);
        std::vector<int32_t> st(pairs.size());
        check(wk, vs_clip_align(wk.clip, pairs.data(), (int)pairs.size(), T.data(), st.data(), nullptr, VS_MEM_HOST), "align");
        for (size_t i = 0; i < pairs.size(); i++) {
            const int f = std::max(first, 1) + (int)i;       // disjoint frame ranges per worker: no locking needed
            m_meas[f].A = T[4 * i]; m_meas[f].B = T[4 * i + 1]; m_meas[f].TX = T[4 * i + 2]; m_meas[f].TY = T[4 * i + 3];
            m_ok[f] = st[i] != 0;
        }
    });

    // ---- phase 2: the sequential trajectory on the host over the gathered table
    StabilizerTrajectory traj(m_params);
    std::vector<SimilarityTransform> corr;
    for (int f = 0; f < n; f++) {
        SimilarityTransform c;
        if (traj.push(m_meas[f], m_ok[f] != 0, m_w, m_h, c)) corr.push_back(c);
    }
    const int produced = (int)corr.size();          // correction i belongs to frame i
    if (produced == 0) return 0;
    if (!out) throw std::runtime_error("MultiGpuStabilizer: output buffer is NULL");

    // ---- phase 3: per-GPU warp of the due frames that live on that GPU
    run_all([&](int r) {
        Worker& wk = m_workers[r];
        const int first = chunks[r].first, last = std::min(chunks[r].second, produced);
        if (last <= first) return;
        std::vector<int32_t> slots;
        std::vector<double> T;
        for (int f = first; f < last; f++) {
            slots.push_back(f - up_first[r]);
            T.insert(T.end(), {corr[f].A, corr[f].B, corr[f].TX, corr[f].TY});
        }
        // warps in batches, each batch's frames copied out on the copy-out stream while the next batch is warped
        const int batch = 32, cnt = (int)slots.size();
        for (int b0 = 0; b0 < cnt; b0 += batch) {
            const int bc = std::min(batch, cnt - b0);
            check(wk, vs_clip_warp_to_host_async(wk.clip, slots.data() + b0, bc, T.data() + 4 * (size_t)b0, VS_WARP_CV_EXACT_BILINEAR,
                                                 VS_BORDER_CONSTANT0, m_crop, out + (size_t)out_frame_stride * (first + b0),
                                                 out_frame_stride), "warp");
        }
        check(wk, vs_clip_sync_transfers(wk.clip), "transfers");
    });
    return produced;
}

}  // namespace vstab
