// partitioned.cpp — see partitioned.hpp.
#include "partitioned.hpp"

#include <stdexcept>

#include "aligner_impl.hpp"

namespace vstab {

VideoPartition::VideoPartition(long total_frames, int workers, int sub_frames, int block_subchunks)
    : total(total_frames), world(workers), sub(sub_frames), block(block_subchunks)
{
    if (total < 1 || world < 1 || sub < 2 || (sub & 1) || block < 1)
        throw std::runtime_error("VideoPartition: sub-chunks must be an even number of frames (keyframes are the odd frames of the video)");
}

// ================================================================== host half
PartitionedTrajectory::PartitionedTrajectory(int rank, int world, int width, int height, long total_frames, int sub_frames,
                                             int block_subchunks, const VideoStabilizerParams& params,
                                             const std::string& exchange_name, int host_threads)
    : m_w(width), m_h(height), m_rank(rank), m_params(params), m_plan(total_frames, world, sub_frames, block_subchunks),
      m_chain(params)
{
    if (width <= 0 || height <= 0 || rank < 0 || rank >= world) throw std::runtime_error("PartitionedTrajectory: bad arguments");
    if (params.lag < 0 || sub_frames < params.lag)
        throw std::runtime_error("PartitionedTrajectory: a sub-chunk must be at least `lag` frames long");
    for (int j = 0; j < m_plan.count(); j++) {
        if (m_plan.owner(j) != rank) continue;
        Sub s;
        s.j = j; s.a = m_plan.first(j); s.b = m_plan.last(j);
        s.halo = (s.a > 0 && (m_subs.empty() || m_subs.back().j != j - 1)) ? 1 : 0;
        s.local0 = (int)m_local.size();
        s.own0 = m_own_frames;
        if (s.halo) { m_local.push_back(s.a - 1); m_halo.push_back(1); }
        for (long f = s.a; f < s.b; f++) {
            m_local.push_back(f); m_halo.push_back(0);
            m_own_frames++;
            if (f + params.lag < total_frames) m_outputs++;
        }
        m_subs.push_back(s);
    }
    m_xchg.reset(new TrajectoryExchange(exchange_name, rank == 0, rank, world, total_frames, m_plan.count()));
    m_pool.reset(new WorkerPool(std::max(1, host_threads)));
    m_corr.resize(m_own_frames);
    m_corr_ready.resize(m_own_frames);
}

const PartitionedTrajectory::Sub* PartitionedTrajectory::own_sub_of(long frame) const
{
    if (frame < 0 || frame >= m_plan.total) return nullptr;
    const int j = m_plan.subchunk_of(frame);
    if (m_plan.owner(j) != m_rank) return nullptr;
    auto it = std::lower_bound(m_subs.begin(), m_subs.end(), j, [](const Sub& s, int jj) { return s.j < jj; });
    return it != m_subs.end() && it->j == j ? &*it : nullptr;
}

int PartitionedTrajectory::local_of_own(int own_index) const
{
    auto it = std::upper_bound(m_subs.begin(), m_subs.end(), own_index, [](int v, const Sub& s) { return v < s.own0; });
    const Sub& s = *(it - 1);
    return s.local0 + s.halo + (own_index - s.own0);
}

void PartitionedTrajectory::begin_video()
{
    m_xchg->begin(++m_generation);
    m_chain = TrajectoryChain(m_params);
    m_chain_pos = 0; m_sm_waited = -1; m_due = 0;
    std::fill(m_corr_ready.begin(), m_corr_ready.end(), 0);
}

void PartitionedTrajectory::advance_chain(long upto)
{
    const SimilarityTransform* meas = m_xchg->measurements();
    const SimilarityTransform* sm = m_xchg->smoothed();
    const uint8_t* ok = m_xchg->ok();
    for (long n = m_chain_pos; n < upto; n++) {
        const int jj = m_plan.subchunk_of(n);
        while (m_sm_waited < jj) m_xchg->wait_smoothed(++m_sm_waited);   // published after the sub-chunk's measurements
        SimilarityTransform corr;
        const SimilarityTransform oldest = n >= m_params.lag ? meas[n - m_params.lag] : SimilarityTransform();
        if (!m_chain.step(n, ok[n] != 0, oldest, sm[n], m_w, m_h, corr)) continue;
        const long f = n - m_params.lag;
        if (const Sub* s = own_sub_of(f)) {
            const int idx = s->own0 + (int)(f - s->a);
            m_corr[idx] = corr;
            m_corr_ready[idx] = 1;
        }
    }
    m_chain_pos = std::max(m_chain_pos, upto);
    while (m_due < m_outputs && m_corr_ready[m_due]) m_due++;
}

void PartitionedTrajectory::submit(int i, const double* T, const int32_t* status)
{
    const Sub& s = m_subs.at(i);
    const long p0 = std::max(s.a, 1L);
    SimilarityTransform* meas = m_xchg->measurements();
    uint8_t* ok = m_xchg->ok();
    for (long f = s.a; f < s.b; f++) {
        SimilarityTransform m;          // frame 0 has no predecessor: identity, false (alignment.cpp:231-234)
        bool good = false;
        if (f >= p0) {
            const size_t p = (size_t)(f - p0);
            m.A = T[4 * p]; m.B = T[4 * p + 1]; m.TX = T[4 * p + 2]; m.TY = T[4 * p + 3];
            good = status[p] != 0;
        }
        meas[f] = m;
        ok[f] = good ? 1 : 0;
    }
    m_xchg->publish_raw(s.j);
    // the smoothed transforms of pushes [a, b) read the raw measurements from a - lagAhead - lagBehind on
    if (m_params.enable_smoother) {
        const long lo = std::max(0L, s.a - m_params.smoother_memory - m_params.lag);
        for (int jj = m_plan.subchunk_of(lo); jj < s.j; jj++) m_xchg->wait_raw(jj);
    }
    smooth_pushes(meas, s.a, s.b, m_params, m_xchg->smoothed() + s.a, m_pool.get());
    m_xchg->publish_smoothed(s.j);
    advance_chain(s.b);
}

void PartitionedTrajectory::flush()
{
    // the last `lag` own frames are decided by pushes of the following (foreign) sub-chunk
    if (!m_subs.empty()) advance_chain(std::min(m_plan.total, m_subs.back().b + m_params.lag));
}

void PartitionedTrajectory::end_video()
{
    m_meas_copy.assign(m_xchg->measurements(), m_xchg->measurements() + m_chain_pos);
    m_ok_copy.assign(m_xchg->ok(), m_xchg->ok() + m_chain_pos);
    m_xchg->finish();
}

// ================================================================== GPU half
PartitionedStabilizer::PartitionedStabilizer(int device, int rank, int world, int width, int height, long total_frames,
                                             int sub_frames, int block_subchunks, const VideoStabilizerParams& params,
                                             const std::string& exchange_name, bool resident, int host_threads, int lanes, bool nv12)
    : m_w(width), m_h(height), m_crop(std::max(0, params.crop_pixels)), m_resident(resident), m_nv12(nv12), m_params(params),
      m_traj(rank, world, width, height, total_frames, sub_frames, block_subchunks, params, exchange_name, host_threads)
{
    if (2 * m_crop >= width || 2 * m_crop >= height) throw std::runtime_error("PartitionedStabilizer: crop_pixels removes the whole frame");
    if (lanes <= 0) lanes = resident ? std::max(2, block_subchunks) : 2;
    m_lanes = m_max_lanes = std::max(1, std::min(lanes, VS_CLIP_SOLVER_LANES));
    const int locals = m_traj.local_count();
    m_capacity = std::max(2, resident ? locals : std::min(locals, (m_max_lanes + 1) * (sub_frames + 1)));
    if (vs_ctx_create(device, &m_ctx) != VS_OK)
        throw std::runtime_error(std::string("PartitionedStabilizer: cannot create a GPU context: ") + vs_last_error(nullptr));
    vs_align_params cp;
    to_c_params(params.aligner, &cp);
    if (vs_clip_create(m_ctx, width, height, m_capacity, m_max_lanes * sub_frames, &cp, nv12 ? VS_CLIP_NV12 : 0, &m_clip) != VS_OK) {
        const std::string msg = std::string("PartitionedStabilizer: ") + vs_last_error(m_ctx);
        vs_ctx_destroy(m_ctx);
        m_ctx = nullptr;
        throw std::runtime_error(msg);
    }
    m_T.resize((size_t)std::max(sub_frames, m_capacity) * 4);
    m_status.resize(sub_frames);
}

PartitionedStabilizer::~PartitionedStabilizer()
{
    if (m_clip) vs_clip_destroy(m_clip);
    if (m_ctx) vs_ctx_destroy(m_ctx);
}

void PartitionedStabilizer::check(int rc, const char* what) const
{
    if (rc != VS_OK) throw std::runtime_error(std::string("PartitionedStabilizer: ") + what + ": " + vs_last_error(m_ctx));
}

// local entries [local0, local0 + n) occupy contiguous slot runs of the ring
template <typename F>
void PartitionedStabilizer::for_slot_runs(int local0, int n, F f) const
{
    int done = 0;
    while (done < n) {
        const int slot = slot_of_local(local0 + done);
        const int run = std::min(n - done, m_capacity - slot);
        f(slot, local0 + done, run);
        done += run;
    }
}

void PartitionedStabilizer::upload_resident(const uint8_t* frames, int64_t row_stride, int64_t frame_stride, int mem)
{
    if (!m_resident) throw std::runtime_error("PartitionedStabilizer: upload_resident needs a resident ring");
    for_slot_runs(0, m_traj.local_count(), [&](int slot, int local, int run) {
        check(vs_clip_upload(m_clip, slot, run, frames + (size_t)frame_stride * local, row_stride, frame_stride, mem), "upload");
    });
}

// GPU work of own sub-chunk i: (upload,) pyramids, keyframe features, the solver launch on lane i % lanes
void PartitionedStabilizer::issue(int i, const uint8_t* frames, int64_t row_stride, int64_t frame_stride)
{
    const PartitionedTrajectory::Sub& s = m_traj.subs()[i];
    const int cnt = s.halo + (int)(s.b - s.a);
    if (frames) {
        for_slot_runs(s.local0, cnt, [&](int slot, int local, int run) {
            check(vs_clip_upload_async(m_clip, slot, run, frames + (size_t)frame_stride * local, row_stride, frame_stride), "upload");
        });
        check(vs_clip_wait_uploads(m_clip), "wait for uploads");
    }
    for_slot_runs(s.local0, cnt, [&](int slot, int, int run) { check(vs_clip_build_pyramids(m_clip, slot, run), "pyramids"); });
    m_slots.clear();
    for (int e = s.local0; e < s.local0 + cnt; e++)
        if (m_traj.local_frame(e) & 1) m_slots.push_back(slot_of_local(e));
    if (!m_slots.empty()) check(vs_clip_build_keyframes(m_clip, m_slots.data(), (int)m_slots.size()), "keyframes");
    // every pair (f-1 -> f); roles as reference alignment.cpp:357,396-397,690-693.  Frame f-1 is the local entry before f
    // (the halo, or the last frame of the previous own sub-chunk)
    m_pairs.clear();
    for (long f = std::max(s.a, 1L); f < s.b; f++) {
        const int e = s.local0 + s.halo + (int)(f - s.a);
        const int cur = slot_of_local(e), prev = slot_of_local(e - 1);
        vs_pair p;
        if (f & 1) { p.template_slot = prev; p.keyframe_slot = cur; p.invert = 0; }
        else       { p.template_slot = cur; p.keyframe_slot = prev; p.invert = 1; }
        m_pairs.push_back(p);
    }
    const int lane = i % m_lanes;
    check(vs_clip_align_async(m_clip, m_pairs.data(), (int)m_pairs.size(), lane * m_traj.partition().sub, lane), "align");
}

// warp the own frames whose correction is decided, in order (crop fused)
void PartitionedStabilizer::warp_due(uint8_t* out, int64_t out_frame_stride, int out_mem, bool async_to_host)
{
    const int upto = m_traj.due();
    while (m_emitted < upto) {
        const int cnt = std::min(upto - m_emitted, m_capacity);
        if (!out) throw std::runtime_error("PartitionedStabilizer: output buffer is NULL");
        m_slots.clear();
        for (int k = 0; k < cnt; k++) {
            const int idx = m_emitted + k;
            m_slots.push_back(slot_of_local(m_traj.local_of_own(idx)));
            const SimilarityTransform& c = m_traj.correction(idx);
            m_T[4 * k] = c.A; m_T[4 * k + 1] = c.B; m_T[4 * k + 2] = c.TX; m_T[4 * k + 3] = c.TY;
        }
        uint8_t* dst = out + (size_t)out_frame_stride * m_emitted;
        if (async_to_host)
            check(vs_clip_warp_to_host_async(m_clip, m_slots.data(), cnt, m_T.data(), VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0,
                                             m_crop, dst, out_frame_stride), "warp");
        else
            check(vs_clip_warp(m_clip, m_slots.data(), cnt, m_T.data(), VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0, m_crop, dst,
                               out_frame_stride, out_mem), "warp");
        m_emitted += cnt;
    }
}

int PartitionedStabilizer::stabilize(const uint8_t* frames, int64_t row_stride, int64_t frame_stride, uint8_t* out,
                                     int64_t out_frame_stride, int out_mem)
{
    if (!frames && !m_resident) throw std::runtime_error("PartitionedStabilizer: a streaming ring needs the frames");
    if (frames && m_resident) throw std::runtime_error("PartitionedStabilizer: a resident ring takes its frames through upload_resident()");
    m_traj.begin_video();
    m_emitted = 0;
    const bool async_to_host = frames != nullptr && out_mem == VS_MEM_HOST;
    const int K = (int)m_traj.subs().size(), D = m_lanes;
    for (int i = 0; i < std::min(D - 1, K); i++) issue(i, frames, row_stride, frame_stride);
    for (int i = 0; i < K; i++) {
        if (i + D - 1 < K) issue(i + D - 1, frames, row_stride, frame_stride);   // its GPU work runs beside the host work below
        check(vs_clip_align_wait(m_clip, i % m_lanes, m_T.data(), m_status.data()), "align");
        m_traj.submit(i, m_T.data(), m_status.data());
        warp_due(out, out_frame_stride, out_mem, async_to_host);
    }
    m_traj.flush();
    warp_due(out, out_frame_stride, out_mem, async_to_host);
    if (async_to_host) check(vs_clip_sync_transfers(m_clip), "transfers");
    m_traj.end_video();
    if (m_emitted != m_traj.output_count()) throw std::runtime_error("PartitionedStabilizer: internal error, not every own frame was produced");
    return m_emitted;
}

}  // namespace vstab
