// grid_search.cpp — see grid_search.hpp.
#include "grid_search.hpp"

#include <math.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "aligner_impl.hpp"

namespace vstab {

namespace {
double median_of(std::vector<double>& v)
{
    if (v.empty()) return 0.0;
    const size_t n = v.size() / 2;
    std::nth_element(v.begin(), v.begin() + n, v.end());
    double med = v[n];
    if (v.size() % 2 == 0) {                      // grid_search_align.cpp:14-24
        std::nth_element(v.begin(), v.begin() + n - 1, v.end());
        med = 0.5 * (med + v[n - 1]);
    }
    return med;
}

// pairs (f-1 -> f) of a sequence stored in consecutive slots; keyframes are the odd frames of the SEQUENCE
void sequence_pairs(int slot0, int n, std::vector<vs_pair>& pairs, std::vector<int32_t>& keys)
{
    for (int f = 1; f < n; f++) {
        vs_pair p;
        if (f & 1) { p.template_slot = slot0 + f - 1; p.keyframe_slot = slot0 + f; p.invert = 0; keys.push_back(slot0 + f); }
        else       { p.template_slot = slot0 + f; p.keyframe_slot = slot0 + f - 1; p.invert = 1; }
        pairs.push_back(p);
    }
}
}  // namespace

double flow_median_px(const SimilarityTransform& T, int w, int h)
{
    std::vector<double> mags;
    const int gx = 32, gy = 18;
    const double cx = w * 0.5, cy = h * 0.5;
    for (int j = 0; j < gy; j++)
        for (int i = 0; i < gx; i++) {
            Point p;
            p.x = (i + 0.5) * w / gx;
            p.y = (j + 0.5) * h / gy;
            const Point q = T.warp(p, cx, cy);
            mags.push_back(hypot(q.x - p.x, q.y - p.y));
        }
    // eval_jitter.cpp:59-61: the upper median (no averaging for even counts)
    const size_t n = mags.size() / 2;
    std::nth_element(mags.begin(), mags.begin() + n, mags.end());
    return mags[n];
}

std::vector<AlignerGridSearch::Combo> AlignerGridSearch::reference_grid()
{
    std::vector<Combo> combos;
    for (bool pc : {false, true})
        for (double thr : {0.02, 0.03, 0.05})
            for (float frac : {0.3f, 0.5f, 0.8f})
                for (double md : {6.0, 8.0, 10.0}) combos.push_back({pc, thr, frac, md});
    return combos;
}

AlignerGridSearch::AlignerGridSearch(int device, int width, int height, int max_frames, int max_combos, int crop_pixels)
    : m_w(width), m_h(height), m_crop(std::max(0, crop_pixels)), m_max_frames(max_frames), m_max_combos(max_combos)
{
    if (width <= 0 || height <= 0 || max_frames < 2 || max_combos < 1) throw std::runtime_error("AlignerGridSearch: bad arguments");
    if (2 * m_crop >= width || 2 * m_crop >= height) throw std::runtime_error("AlignerGridSearch: crop_pixels removes the whole frame");
    if (vs_ctx_create(device, &m_ctx) != VS_OK)
        throw std::runtime_error(std::string("AlignerGridSearch: cannot create a GPU context: ") + vs_last_error(nullptr));
    // output sequences are scored a group of combinations at a time (about 2 GB of frames per group)
    const size_t out_frame = (size_t)(width - 2 * m_crop) * (height - 2 * m_crop) * 3;
    m_group = (int)std::max<size_t>(1, std::min<size_t>(max_combos, ((size_t)2 << 30) / (out_frame * max_frames)));
    vs_align_params ap;
    vs_align_params_default(&ap);
    int rc = vs_clip_create(m_ctx, width, height, max_frames, std::max(max_frames, (max_frames - 1) * max_combos), &ap, 0, &m_in);
    if (rc == VS_OK)
        rc = vs_clip_create(m_ctx, width - 2 * m_crop, height - 2 * m_crop, m_group * max_frames, m_group * max_frames, &ap, 0, &m_out);
    if (rc == VS_OK) rc = vs_dev_alloc(m_ctx, out_frame * m_group * max_frames, (void**)&m_dev_out);
    if (rc != VS_OK) {
        const std::string msg = std::string("AlignerGridSearch: ") + vs_last_error(m_ctx);
        if (m_out) vs_clip_destroy(m_out);
        if (m_in) vs_clip_destroy(m_in);
        vs_ctx_destroy(m_ctx);
        throw std::runtime_error(msg);
    }
}

AlignerGridSearch::~AlignerGridSearch()
{
    if (m_dev_out) vs_dev_free(m_ctx, m_dev_out);
    if (m_out) vs_clip_destroy(m_out);
    if (m_in) vs_clip_destroy(m_in);
    if (m_ctx) vs_ctx_destroy(m_ctx);
}

long AlignerGridSearch::launches() const { return (long)vs_ctx_launch_count(m_ctx); }

void AlignerGridSearch::check(int rc, const char* what) const
{
    if (rc != VS_OK) throw std::runtime_error(std::string("AlignerGridSearch: ") + what + ": " + vs_last_error(m_ctx));
}

// jitter of `sequences` sequences of n frames each, sequence q in slots [first_slot + q * seq_stride, ... + n): pyramids
// and keyframe features of all of them, ONE solver launch over all their pairs, medians on the host
JitterScore AlignerGridSearch::score(vs_clip* clip, int first_slot, int n, int w, int h, int sequences, int seq_stride,
                                     std::vector<JitterScore>* per_seq)
{
    JitterScore all;
    if (n < 2 || sequences < 1) return all;
    std::vector<vs_pair> pairs;
    std::vector<int32_t> keys;
    for (int q = 0; q < sequences; q++) {
        check(vs_clip_build_pyramids(clip, first_slot + q * seq_stride, n), "pyramids");
        sequence_pairs(first_slot + q * seq_stride, n, pairs, keys);
    }
    check(vs_clip_build_keyframes(clip, keys.data(), (int)keys.size()), "keyframes");
    std::vector<double> T(pairs.size() * 4);
    std::vector<int32_t> st(pairs.size());
    check(vs_clip_align(clip, pairs.data(), (int)pairs.size(), T.data(), st.data(), nullptr, VS_MEM_HOST), "align");
    std::vector<double> every;
    for (int q = 0; q < sequences; q++) {
        std::vector<double> meds;
        JitterScore s;
        for (int i = 0; i < n - 1; i++) {
            const size_t k = (size_t)q * (n - 1) + i;
            s.pairs++;
            if (!st[k]) { s.failed++; continue; }
            SimilarityTransform t;
            t.A = T[4 * k]; t.B = T[4 * k + 1]; t.TX = T[4 * k + 2]; t.TY = T[4 * k + 3];
            meds.push_back(flow_median_px(t, w, h));
        }
        every.insert(every.end(), meds.begin(), meds.end());
        s.median_px = median_of(meds);
        all.pairs += s.pairs; all.failed += s.failed;
        if (per_seq) per_seq->push_back(s);
    }
    all.median_px = median_of(every);
    return all;
}

JitterScore AlignerGridSearch::measure_jitter(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride)
{
    if (n < 0 || n > m_max_frames) throw std::runtime_error("AlignerGridSearch: more frames than max_frames");
    check(vs_clip_upload(m_in, 0, n, frames, row_stride, frame_stride, VS_MEM_HOST), "upload");
    return score(m_in, 0, n, m_w, m_h, 1, 0, nullptr);
}

std::vector<AlignerGridSearch::Result> AlignerGridSearch::run(const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride,
                                                              const std::vector<Combo>& combos, JitterScore* input_jitter)
{
    const int S = (int)combos.size();
    if (n < 3 || n > m_max_frames || S < 1 || S > m_max_combos) throw std::runtime_error("AlignerGridSearch: bad clip or grid size");
    const JitterScore in = measure_jitter(frames, n, row_stride, frame_stride);     // uploads; pyramids and features are built
    if (input_jitter) *input_jitter = in;

    // ---- every pair under every combination: one launch
    std::vector<vs_pair> pairs;
    std::vector<int32_t> keys;
    sequence_pairs(0, n, pairs, keys);
    std::vector<vs_sweep_params> sets(S);
    for (int s = 0; s < S; s++) {
        sets[s].threshold = combos[s].threshold; sets[s].max_displacement = combos[s].max_displacement;
        sets[s].smallest_fraction = combos[s].smallest_fraction; sets[s].max_iters = VideoAlignerParams().max_iters;
        sets[s].phase_correlate = combos[s].phase_correlate ? 1 : 0;
    }
    const int np = n - 1;
    m_T.assign((size_t)S * np * 4, 0.0);
    m_status.assign((size_t)S * np, 0);
    check(vs_clip_align_sweep(m_in, pairs.data(), np, sets.data(), S, m_T.data(), m_status.data(), VS_MEM_HOST), "align sweep");

    // ---- per combination: the stabilizer's trajectory as grid_search_align.cpp:167-180 configures it (smoother off,
    //      lag 1), then the warped output sequences a group of combinations at a time, then their jitter
    VideoStabilizerParams sp;
    sp.enable_smoother = false;
    sp.lag = 1;
    sp.smoother_memory = 0;
    sp.crop_pixels = m_crop;
    const int ow = m_w - 2 * m_crop, oh = m_h - 2 * m_crop;
    const size_t out_frame = (size_t)ow * oh * 3;
    const int n_out = n - sp.lag;
    std::vector<Result> results(S);
    for (int g0 = 0; g0 < S; g0 += m_group) {
        const int G = std::min(m_group, S - g0);
        for (int q = 0; q < G; q++) {
            const int s = g0 + q;
            results[s].combo = combos[s];
            StabilizerTrajectory traj(sp);
            std::vector<int32_t> slots;
            std::vector<double> corr;
            for (int f = 0; f < n; f++) {
                SimilarityTransform meas;
                bool ok = false;
                if (f > 0) {
                    const size_t k = (size_t)s * np + (f - 1);
                    meas.A = m_T[4 * k]; meas.B = m_T[4 * k + 1]; meas.TX = m_T[4 * k + 2]; meas.TY = m_T[4 * k + 3];
                    ok = m_status[k] != 0;
                    if (!ok) results[s].failed_alignments++;
                }
                SimilarityTransform c;
                if (traj.push(meas, ok, m_w, m_h, c)) {
                    slots.push_back((int32_t)slots.size());      // output i is frame i (lag frames behind)
                    corr.insert(corr.end(), {c.A, c.B, c.TX, c.TY});
                }
            }
            check(vs_clip_warp(m_in, slots.data(), (int)slots.size(), corr.data(), VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0,
                               m_crop, m_dev_out + out_frame * (size_t)q * n_out, (int64_t)out_frame, VS_MEM_DEVICE), "warp");
        }
        check(vs_clip_upload(m_out, 0, G * n_out, m_dev_out, (int64_t)ow * 3, (int64_t)out_frame, VS_MEM_DEVICE), "regroup");
        std::vector<JitterScore> per;
        score(m_out, 0, n_out, ow, oh, G, n_out, &per);
        for (int q = 0; q < G; q++) {
            results[g0 + q].out = per[q];
            results[g0 + q].ratio = in.median_px > 0 ? per[q].median_px / in.median_px : 0.0;
        }
    }
    return results;
}

}  // namespace vstab
