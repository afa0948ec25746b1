// clip_stabilizer.cpp — see clip_stabilizer.hpp.
#include "clip_stabilizer.hpp"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "aligner_impl.hpp"

namespace vstab {

ClipStabilizer::ClipStabilizer(int device, int width, int height, int chunk_frames, const VideoStabilizerParams& params, bool nv12)
    : m_w(width), m_h(height), m_chunk(chunk_frames), m_capacity(chunk_frames + std::max(0, params.lag) + 1),
      m_crop(std::max(0, params.crop_pixels)), m_nv12(nv12), m_params(params), m_trajectory(params)
{
    if (width <= 0 || height <= 0 || chunk_frames <= 0) throw std::runtime_error("ClipStabilizer: bad geometry");
    if (2 * m_crop >= width || 2 * m_crop >= height) throw std::runtime_error("ClipStabilizer: crop_pixels removes the whole frame");
    if (vs_ctx_create(device, &m_ctx) != VS_OK)
        throw std::runtime_error(std::string("ClipStabilizer: cannot create a GPU context: ") + vs_last_error(nullptr));
    vs_align_params cp;
    to_c_params(params.aligner, &cp);
    if (vs_clip_create(m_ctx, width, height, m_capacity, chunk_frames, &cp, nv12 ? VS_CLIP_NV12 : 0, &m_clip) != VS_OK) {
        std::string msg = std::string("ClipStabilizer: ") + vs_last_error(m_ctx);
        vs_ctx_destroy(m_ctx);
        m_ctx = nullptr;
        throw std::runtime_error(msg);
    }
    m_pairs.reserve(chunk_frames); m_T.resize((size_t)chunk_frames * 4); m_status.resize(chunk_frames); m_slots.reserve(chunk_frames);
}

ClipStabilizer::~ClipStabilizer()
{
    if (m_clip) vs_clip_destroy(m_clip);
    if (m_ctx) vs_ctx_destroy(m_ctx);
}

void ClipStabilizer::reset()
{
    m_fed = m_emitted = 0;
    m_trajectory = StabilizerTrajectory(m_params);
    m_meas.clear(); m_ok.clear(); m_corr.clear();
}

void ClipStabilizer::check(int rc, const char* what) const
{
    if (rc != VS_OK) throw std::runtime_error(std::string("ClipStabilizer: ") + what + ": " + vs_last_error(m_ctx));
}

// frames [first_frame, first_frame+n) occupy at most two contiguous slot runs of the ring
template <typename F>
void ClipStabilizer::for_slot_runs(long first_frame, int n, F f) const
{
    int done = 0;
    while (done < n) {
        const int slot = (int)((first_frame + done) % m_capacity);
        const int run = std::min(n - done, m_capacity - slot);
        f(slot, done, run);
        done += run;
    }
}

void ClipStabilizer::upload_only(long first_frame, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem)
{
    if (n < 0 || n > m_capacity) throw std::runtime_error("ClipStabilizer: more frames than ring slots");
    for_slot_runs(first_frame, n, [&](int slot, int done, int run) {
        check(vs_clip_upload(m_clip, slot, run, frames + (size_t)frame_stride * done, row_stride, frame_stride, mem), "upload");
    });
}

int ClipStabilizer::feed(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem,
                         uint8_t* out, int64_t out_frame_stride, int out_mem)
{
    if (n < 0 || n > m_chunk) throw std::runtime_error("ClipStabilizer: feed() takes at most chunk_frames frames");
    if (mem == VS_MEM_HOST && out_mem == VS_MEM_HOST && n > m_sub)
        return feed_pipelined(frames, n, row_stride, frame_stride, out, out_frame_stride);
    upload_only(m_fed, frames, n, row_stride, frame_stride, mem);
    return process(n, out, out_frame_stride, out_mem, false, false);
}

int ClipStabilizer::feed_pipelined(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride,
                                   uint8_t* out, int64_t out_frame_stride)
{
    // NOTE: m_fed advances inside process(); uploads address frames by their position in this call
    const long base = m_fed;
    auto upload_chunk = [&](int first, int count) {
        for_slot_runs(base + first, count, [&](int slot, int done, int run) {
            check(vs_clip_upload_async(m_clip, slot, run, frames + (size_t)frame_stride * (first + done), row_stride, frame_stride),
                  "upload");
        });
    };
    int produced = 0;
    upload_chunk(0, std::min(m_sub, n));
    for (int first = 0; first < n; first += m_sub) {
        const int count = std::min(m_sub, n - first);
        check(vs_clip_wait_uploads(m_clip), "wait for uploads");          // compute of this sub-chunk waits for its frames only
        if (first + count < n) upload_chunk(first + count, std::min(m_sub, n - first - count));   // overlaps the kernels below
        produced += process(count, out ? out + (size_t)out_frame_stride * produced : nullptr, out_frame_stride, VS_MEM_HOST,
                            first > 0, true);
    }
    check(vs_clip_sync_transfers(m_clip), "transfers");
    return produced;
}

int ClipStabilizer::feed_resident(int n, uint8_t* out, int64_t out_frame_stride, int out_mem)
{
    if (n < 0 || n > m_chunk) throw std::runtime_error("ClipStabilizer: feed_resident() takes at most chunk_frames frames");
    return process(n, out, out_frame_stride, out_mem, false, false);
}

int ClipStabilizer::process(int n, uint8_t* out, int64_t out_frame_stride, int out_mem, bool append_records, bool async_to_host)
{
    const size_t rec0 = append_records ? m_meas.size() : 0;
    m_meas.resize(rec0 + n);
    m_ok.resize(rec0 + n);
    if (!append_records) m_corr.clear();
    if (n == 0) return 0;
    const long f0 = m_fed;

    // ---- every pair (f-1 -> f); roles as reference alignment.cpp:357,396-397,690-693
    m_pairs.clear();
    for (long f = std::max(f0, 1L); f < f0 + n; f++) {
        vs_pair p;
        const int cur = (int)(f % m_capacity), prev = (int)((f - 1) % m_capacity);
        if (f & 1) { p.template_slot = prev; p.keyframe_slot = cur; p.invert = 0; }
        else       { p.template_slot = cur; p.keyframe_slot = prev; p.invert = 1; }
        m_pairs.push_back(p);
    }
    const int np = (int)m_pairs.size();
    const long first_pair = std::max(f0, 1L);

    // pyramids of frames [a, b), keyframe features of the odd ones
    auto ingest = [&](long a, long b) {
        if (b <= a) return;
        for_slot_runs(a, (int)(b - a), [&](int slot, int, int run) { check(vs_clip_build_pyramids(m_clip, slot, run), "pyramids"); });
        m_slots.clear();
        for (long f = a; f < b; f++)
            if (f & 1) m_slots.push_back((int32_t)(f % m_capacity));
        if (!m_slots.empty()) check(vs_clip_build_keyframes(m_clip, m_slots.data(), (int)m_slots.size()), "keyframes");
    };

    // Large device-resident chunks run as up to four pieces on the clip's solver lanes.  A solve lasts as long as its
    // slowest pair (~0.9 ms) however many pairs it has and is bound by gather latency; the other stages are bound by
    // bandwidth.  So the solve of a piece runs beside the pyramids / keyframe features of the next pieces and beside the
    // warps of the previous ones, and the chunk takes about ingest + one solve + the last piece's warps.
    const bool overlap = async_to_host || out_mem == VS_MEM_DEVICE;
    int lanes = 1;
    if (m_lanes > 1 && overlap && !async_to_host && np >= 128) lanes = std::min({m_lanes, VS_CLIP_SOLVER_LANES, np / 64});
    int lane_first[VS_CLIP_SOLVER_LANES + 1] = {0};   // first pair of each piece; piece l = pairs [lane_first[l], lane_first[l+1])
    if (lanes > 1) {
        long a = f0;
        for (int l = 0; l < lanes; l++) {
            const long b = l + 1 == lanes ? f0 + n : (f0 + (long)n * (l + 1) / lanes) & ~1L;   // pieces end on even frames
            lane_first[l + 1] = (int)(std::max(b, first_pair) - first_pair);
            ingest(a, b);
            check(vs_clip_align_async(m_clip, m_pairs.data() + lane_first[l], lane_first[l + 1] - lane_first[l], lane_first[l], l), "align");
            a = b;
        }
    } else {
        ingest(f0, f0 + n);
        if (np) check(vs_clip_align(m_clip, m_pairs.data(), np, m_T.data(), m_status.data(), nullptr, VS_MEM_HOST), "align");
    }
    int lane_waited = 0;              // pieces whose results are on the host

    // ---- sequential host trajectory, with the warps of the frames already decided launched in
    //      batches while the host is still smoothing the later ones (asynchronous outputs only)
    std::vector<int32_t> due_slots;
    std::vector<double> due_T;
    int launched = 0;
    int batch = std::min(16, m_warp_batch);   // first launch early (the GPU idles until then), later ones grow to m_warp_batch
    auto launch_warps = [&](int upto) {
        const int cnt = upto - launched;
        if (cnt <= 0) return;
        if (!out) throw std::runtime_error("ClipStabilizer: output buffer is NULL");
        uint8_t* dst = out + (size_t)out_frame_stride * launched;
        if (async_to_host)
            check(vs_clip_warp_to_host_async(m_clip, due_slots.data() + launched, cnt, due_T.data() + 4 * (size_t)launched,
                                             VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0, m_crop, dst, out_frame_stride), "warp");
        else
            check(vs_clip_warp(m_clip, due_slots.data() + launched, cnt, due_T.data() + 4 * (size_t)launched,
                               VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0, m_crop, dst, out_frame_stride, out_mem), "warp");
        launched = upto;
    };
    const int first_pair_frame = (int)(std::max(f0, 1L) - f0);
    for (int i = 0; i < n; i++) {
        SimilarityTransform meas;
        bool ok = false;
        if (i >= first_pair_frame) {
            const int p = i - first_pair_frame;
            while (lanes > 1 && lane_waited < lanes && p >= lane_first[lane_waited]) {
                check(vs_clip_align_wait(m_clip, lane_waited, m_T.data() + 4 * (size_t)lane_first[lane_waited],
                                         m_status.data() + lane_first[lane_waited]), "align");
                lane_waited++;
            }
            meas.A = m_T[4 * p]; meas.B = m_T[4 * p + 1]; meas.TX = m_T[4 * p + 2]; meas.TY = m_T[4 * p + 3];
            ok = m_status[p] != 0;
        }
        m_meas[rec0 + i] = meas;
        m_ok[rec0 + i] = ok ? 1 : 0;
        SimilarityTransform corr;
        if (m_trajectory.push(meas, ok, m_w, m_h, corr)) {
            m_corr.push_back(corr);
            due_slots.push_back((int32_t)((m_emitted + (long)due_slots.size()) % m_capacity));
            due_T.insert(due_T.end(), {corr.A, corr.B, corr.TX, corr.TY});
            if (overlap && (int)due_slots.size() - launched >= batch) {
                launch_warps((int)due_slots.size());
                batch = std::min(2 * batch, m_warp_batch);
            }
        }
    }
    m_fed = f0 + n;

    // ---- the remaining (or, for synchronous host output, all) due frames in one launch (crop fused)
    const int produced = (int)due_slots.size();
    launch_warps(produced);
    m_emitted += produced;
    return produced;
}

}  // namespace vstab
