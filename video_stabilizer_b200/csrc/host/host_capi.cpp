// host_capi.cpp — C forwarders over the host classes (include/vstab_host.h).
#include "vstab_host.h"

#include <string.h>

#include <exception>
#include <string>
#include <vector>
#include <algorithm>

#include "aligner_impl.hpp"
#include "clip_stabilizer.hpp"
#include "grid_search.hpp"
#include "multi_gpu.hpp"
#include "partitioned.hpp"
#include "stabilizer.hpp"

namespace {

thread_local std::string g_err;

template <typename F>
int guarded(F f)
{
    try {
        return f();
    } catch (const std::exception& e) {
        g_err = e.what();
    } catch (...) {
        g_err = "unknown C++ exception";
    }
    return -1;
}

SimilarityTransform tf(const double* T)
{
    SimilarityTransform t;
    t.A = T[0]; t.B = T[1]; t.TX = T[2]; t.TY = T[3];
    return t;
}
void put(const SimilarityTransform& t, double* T) { T[0] = t.A; T[1] = t.B; T[2] = t.TX; T[3] = t.TY; }

VideoAlignerParams align_params(const vs_align_params* p)
{
    VideoAlignerParams q;
    if (!p) return q;
    q.phase_correlate = p->phase_correlate != 0;
    q.phase_correlate_threshold = p->phase_correlate_threshold;
    q.threshold = p->threshold;
    q.smallest_fraction = p->smallest_fraction;
    q.max_iters = p->max_iters;
    q.pyramid_min_width = p->pyramid_min_width;
    q.pyramid_min_height = p->pyramid_min_height;
    q.max_displacement = p->max_displacement;
    return q;
}

VideoStabilizerParams stab_params(const vsh_stab_params* p)
{
    VideoStabilizerParams q;
    if (!p) return q;
    q.aligner = align_params(&p->aligner);
    q.lag = p->lag; q.smoother_memory = p->smoother_memory; q.lambda = p->lambda;
    q.enable_smoother = p->enable_smoother != 0; q.crop_pixels = p->crop_pixels;
    q.min_disp = p->min_disp; q.max_disp = p->max_disp; q.min_decay = p->min_decay; q.max_decay = p->max_decay;
    return q;
}

template <typename T>
Halide::Runtime::Buffer<T> wrap2(const T* p, int w, int h) { return Halide::Runtime::Buffer<T>(const_cast<T*>(p), w, h); }
template <typename T>
Halide::Runtime::Buffer<T> wrap3(const T* p, int w, int h, int c)
{
    halide_dimension_t shape[3] = {halide_dimension_t(0, w, 1), halide_dimension_t(0, h, w), halide_dimension_t(0, c, w * h)};
    return Halide::Runtime::Buffer<T>(const_cast<T*>(p), 3, shape);
}

struct AlignerBox : public VideoAligner {
    void set_device(int d) { impl_->device = d; }
};

}  // namespace

extern "C" {

void vsh_stab_params_default(vsh_stab_params* p)
{
    if (!p) return;
    VideoStabilizerParams d;
    vs_align_params_default(&p->aligner);
    p->lag = d.lag; p->smoother_memory = d.smoother_memory; p->lambda = d.lambda;
    p->enable_smoother = d.enable_smoother ? 1 : 0; p->crop_pixels = d.crop_pixels;
    p->min_disp = d.min_disp; p->max_disp = d.max_disp; p->min_decay = d.min_decay; p->max_decay = d.max_decay;
}

const char* vsh_last_error(void) { return g_err.c_str(); }

void vsh_tf_inverse(const double T[4], double out[4]) { put(tf(T).inverse(), out); }
void vsh_tf_compose(const double T1[4], const double T2[4], double out[4]) { put(tf(T1).compose(tf(T2)), out); }
void vsh_tf_warp(const double T[4], double px, double py, double out[2])
{
    Point p = tf(T).warp(Point{px, py});
    out[0] = p.x; out[1] = p.y;
}
void vsh_tf_warp_center(const double T[4], double px, double py, double cx, double cy, double out[2])
{
    Point p = tf(T).warp(Point{px, py}, cx, cy);
    out[0] = p.x; out[1] = p.y;
}
double vsh_tf_max_corner_displacement(const double T[4], double w, double h) { return tf(T).maxCornerDisplacement(w, h); }

int vsh_PyrDown(const uint8_t* in, int iw, int ih, uint8_t* out, int ow, int oh)
{
    return guarded([&] {
        auto a = wrap2(in, iw, ih); auto b = wrap2(out, ow, oh);
        return PyrDown(a, b) ? 1 : 0;
    });
}

int vsh_GradXY(const uint8_t* in, int w, int h, float* gx, float* gy)
{
    return guarded([&] {
        auto a = wrap2(in, w, h); auto x = wrap2(gx, w, h); auto y = wrap2(gy, w, h);
        return GradXY(a, x, y) ? 1 : 0;
    });
}

int vsh_GradArgMax(const float* gx, const float* gy, int w, int h, int* tile, uint16_t* lmx, uint16_t* lmy, int capacity)
{
    return guarded([&] {
        auto x = wrap2(gx, w, h); auto y = wrap2(gy, w, h);
        Halide::Runtime::Buffer<uint16_t> ox, oy;   // callee allocates, as upstream
        int t = 0;
        if (!GradArgMax(x, y, t, ox, oy)) return 0;
        *tile = t;
        const size_t n = ox.number_of_elements();
        if ((size_t)capacity < n) throw std::runtime_error("vsh_GradArgMax: output capacity too small");
        memcpy(lmx, ox.data(), n * sizeof(uint16_t));
        memcpy(lmy, oy.data(), n * sizeof(uint16_t));
        return 1;
    });
}

int vsh_SparseJacobian(const float* gx, const float* gy, int w, int h, const uint16_t* lmx, const uint16_t* lmy,
                       int tw, int th, float* jx, float* jy)
{
    return guarded([&] {
        auto x = wrap2(gx, w, h); auto y = wrap2(gy, w, h);
        auto lx = wrap3(lmx, tw, th, 2); auto ly = wrap3(lmy, tw, th, 2);
        Halide::Runtime::Buffer<float> ox, oy;
        if (!SparseJacobian(x, y, lx, ly, ox, oy)) return 0;
        memcpy(jx, ox.data(), ox.size_in_bytes());
        memcpy(jy, oy.data(), oy.size_in_bytes());
        return 1;
    });
}

int vsh_SparseWarpDiff(const uint8_t* tmpl, const uint8_t* key, int w, int h, const uint16_t* lm, int tw, int th,
                       const double T[4], uint16_t* out)
{
    return guarded([&] {
        auto t = wrap2(tmpl, w, h); auto k = wrap2(key, w, h);
        auto l = wrap3(lm, tw, th, 2);
        Halide::Runtime::Buffer<uint16_t> o;
        if (!SparseWarpDiff(t, k, l, tf(T), o)) return 0;
        memcpy(out, o.data(), o.size_in_bytes());
        return 1;
    });
}

int vsh_SparseICA(const uint8_t* tmpl, const uint8_t* key, int w, int h, const uint16_t* selx, int kx,
                  const uint16_t* sely, int ky, const float* jx, const float* jy, const double T[4], double out[4])
{
    return guarded([&] {
        auto t = wrap2(tmpl, w, h); auto k = wrap2(key, w, h);
        auto sx = wrap2(selx, kx, 2); auto sy = wrap2(sely, ky, 2);
        auto ax = wrap2(jx, kx, 4); auto ay = wrap2(jy, ky, 4);
        Halide::Runtime::Buffer<double> o;
        if (!SparseICA(t, k, sx, sy, ax, ay, tf(T), o)) return 0;
        for (int i = 0; i < 4; i++) out[i] = o(i);
        return 1;
    });
}

int vsh_ImageWarp(const uint8_t* in, int w, int h, const double T[4], float* out, int ow, int oh)
{
    return guarded([&] {
        auto a = wrap2(in, w, h); auto o = wrap2(out, ow, oh);
        return ImageWarp(a, tf(T), o) ? 1 : 0;
    });
}

int vsh_warpBySimilarityTransform(const uint8_t* bgr, int w, int h, int64_t row_stride, const double T[4], uint8_t* out)
{
    return guarded([&] {
        cv::Mat src(h, w, CV_8UC3, (void*)bgr, (size_t)row_stride);
        cv::Mat dst = warpBySimilarityTransform(src, tf(T));
        for (int y = 0; y < h; y++) memcpy(out + (size_t)y * w * 3, dst.ptr(y), (size_t)w * 3);
        return 0;
    });
}

void* vsh_smoother_create(int lag_behind, int lag_ahead, double lambda) { return new L1SmootherCenter(lag_behind, lag_ahead, lambda); }
void vsh_smoother_destroy(void* s) { delete (L1SmootherCenter*)s; }
int vsh_smoother_update(void* s, const double meas[4], double out[4])
{
    SimilarityTransform o;
    bool r = ((L1SmootherCenter*)s)->update(tf(meas), o);
    put(o, out);
    return r ? 1 : 0;
}
void vsh_tvl1_relax(const double* data, int n, double lambda, int iterations, double* out) { vstab::tvl1_relax(data, n, lambda, iterations, out); }

void* vsh_trajectory_create(const vsh_stab_params* p) { return new vstab::StabilizerTrajectory(stab_params(p)); }
void vsh_trajectory_destroy(void* t) { delete (vstab::StabilizerTrajectory*)t; }
int vsh_trajectory_push(void* t, const double meas[4], int success, int w, int h, double correction[4])
{
    SimilarityTransform c;
    bool due = ((vstab::StabilizerTrajectory*)t)->push(tf(meas), success != 0, w, h, c);
    put(c, correction);
    return due ? 1 : 0;
}

void* vsh_aligner_create(int device)
{
    AlignerBox* a = nullptr;
    guarded([&] { a = new AlignerBox(); a->set_device(device); return 0; });
    return a;
}
void vsh_aligner_destroy(void* a) { delete (AlignerBox*)a; }
int vsh_aligner_align(void* a, const uint8_t* bgr, int w, int h, int64_t row_stride, const vs_align_params* params, double T[4])
{
    return guarded([&] {
        cv::Mat frame(h, w, CV_8UC3, (void*)bgr, (size_t)row_stride);
        SimilarityTransform t;
        bool ok = ((AlignerBox*)a)->AlignNextFrame(frame, t, align_params(params));
        put(t, T);
        return ok ? 1 : 0;
    });
}

namespace {
struct StabBox : public VideoStabilizer {
    StabBox(const VideoStabilizerParams& p, int device) : VideoStabilizer(p) { static_cast<AlignerBox&>(aligner).set_device(device); }
};
}  // namespace

void* vsh_stabilizer_create(const vsh_stab_params* p, int device)
{
    StabBox* s = nullptr;
    guarded([&] { s = new StabBox(stab_params(p), device); return 0; });
    return s;
}
void vsh_stabilizer_destroy(void* s) { delete (StabBox*)s; }
int vsh_stabilizer_process(void* s, const uint8_t* bgr, int w, int h, int64_t row_stride, uint8_t* out, int* out_w, int* out_h)
{
    return guarded([&] {
        cv::Mat frame(h, w, CV_8UC3, (void*)bgr, (size_t)row_stride);
        cv::Mat r = ((StabBox*)s)->processFrame(frame);
        if (r.empty()) { *out_w = *out_h = 0; return 0; }
        *out_w = r.cols; *out_h = r.rows;
        for (int y = 0; y < r.rows; y++) memcpy(out + (size_t)y * r.cols * 3, r.ptr(y), (size_t)r.cols * 3);
        return 1;
    });
}

void* vsh_clipstab_create(int device, int width, int height, int chunk_frames, const vsh_stab_params* p)
{
    vstab::ClipStabilizer* c = nullptr;
    guarded([&] { c = new vstab::ClipStabilizer(device, width, height, chunk_frames, stab_params(p)); return 0; });
    return c;
}
void* vsh_clipstab_create_nv12(int device, int width, int height, int chunk_frames, const vsh_stab_params* p)
{
    vstab::ClipStabilizer* c = nullptr;
    guarded([&] { c = new vstab::ClipStabilizer(device, width, height, chunk_frames, stab_params(p), true); return 0; });
    return c;
}
void vsh_clipstab_destroy(void* c) { delete (vstab::ClipStabilizer*)c; }
int vsh_clipstab_reset(void* c) { return guarded([&] { ((vstab::ClipStabilizer*)c)->reset(); return 0; }); }
int vsh_clipstab_set_pipeline_frames(void* c, int frames) { ((vstab::ClipStabilizer*)c)->set_pipeline_frames(frames); return 0; }
int vsh_clipstab_set_solver_lanes(void* c, int lanes) { ((vstab::ClipStabilizer*)c)->set_solver_lanes(lanes); return 0; }
int vsh_clipstab_feed(void* c, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem,
                      uint8_t* out, int64_t out_frame_stride, int out_mem)
{
    return guarded([&] { return ((vstab::ClipStabilizer*)c)->feed(frames, n, row_stride, frame_stride, mem, out, out_frame_stride, out_mem); });
}
int vsh_clipstab_upload_only(void* c, int64_t first_frame, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem)
{
    return guarded([&] { ((vstab::ClipStabilizer*)c)->upload_only((long)first_frame, frames, n, row_stride, frame_stride, mem); return 0; });
}
int vsh_clipstab_feed_resident(void* c, int n, uint8_t* out, int64_t out_frame_stride, int out_mem)
{
    return guarded([&] { return ((vstab::ClipStabilizer*)c)->feed_resident(n, out, out_frame_stride, out_mem); });
}
int vsh_clipstab_last_records(void* c, double* meas, uint8_t* ok, double* corrections)
{
    auto* s = (vstab::ClipStabilizer*)c;
    if (meas) for (size_t i = 0; i < s->measurements().size(); i++) put(s->measurements()[i], meas + 4 * i);
    if (ok) for (size_t i = 0; i < s->successes().size(); i++) ok[i] = s->successes()[i];
    if (corrections) for (size_t i = 0; i < s->corrections().size(); i++) put(s->corrections()[i], corrections + 4 * i);
    return (int)s->corrections().size();
}
int vsh_clipstab_out_size(void* c, int* w, int* h)
{
    auto* s = (vstab::ClipStabilizer*)c;
    *w = s->out_width(); *h = s->out_height();
    return 0;
}
void* vsh_multigpu_create(const int32_t* devices, int n_devices, int width, int height, int max_frames, const vsh_stab_params* p)
{
    vstab::MultiGpuStabilizer* m = nullptr;
    guarded([&] {
        m = new vstab::MultiGpuStabilizer(std::vector<int>(devices, devices + n_devices), width, height, max_frames, stab_params(p));
        return 0;
    });
    return m;
}
void vsh_multigpu_destroy(void* m) { delete (vstab::MultiGpuStabilizer*)m; }
int vsh_multigpu_stabilize(void* m, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, uint8_t* out,
                           int64_t out_frame_stride, double* meas, uint8_t* ok)
{
    return guarded([&] {
        auto* s = (vstab::MultiGpuStabilizer*)m;
        int k = s->stabilize(frames, n, row_stride, frame_stride, out, out_frame_stride);
        if (meas) for (size_t i = 0; i < s->measurements().size(); i++) put(s->measurements()[i], meas + 4 * i);
        if (ok) for (size_t i = 0; i < s->successes().size(); i++) ok[i] = s->successes()[i];
        return k;
    });
}

// ---- quality tooling: jitter score and the batched VideoAlignerParams sweep (grid_search.hpp)
void* vsh_gridsearch_create(int device, int width, int height, int max_frames, int max_combos, int crop_pixels)
{
    vstab::AlignerGridSearch* g = nullptr;
    guarded([&] { g = new vstab::AlignerGridSearch(device, width, height, max_frames, max_combos, crop_pixels); return 0; });
    return g;
}
void vsh_gridsearch_destroy(void* g) { delete (vstab::AlignerGridSearch*)g; }
int vsh_gridsearch_reference_grid(double* combos4, int capacity)
{
    const auto grid = vstab::AlignerGridSearch::reference_grid();
    for (size_t i = 0; i < grid.size() && (int)i < capacity; i++) {
        combos4[4 * i] = grid[i].phase_correlate ? 1.0 : 0.0; combos4[4 * i + 1] = grid[i].threshold;
        combos4[4 * i + 2] = grid[i].smallest_fraction; combos4[4 * i + 3] = grid[i].max_displacement;
    }
    return (int)grid.size();
}
double vsh_flow_median_px(const double T[4], int w, int h) { return vstab::flow_median_px(tf(T), w, h); }
int vsh_gridsearch_jitter(void* g, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, double* out3)
{
    return guarded([&] {
        const vstab::JitterScore s = ((vstab::AlignerGridSearch*)g)->measure_jitter(frames, n, row_stride, frame_stride);
        out3[0] = s.median_px; out3[1] = s.pairs; out3[2] = s.failed;
        return 0;
    });
}
int vsh_gridsearch_run(void* gp, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, const double* combos4,
                       int n_combos, double* input3, double* results5, double* T, int32_t* status)
{
    return guarded([&] {
        auto* g = (vstab::AlignerGridSearch*)gp;
        std::vector<vstab::AlignerGridSearch::Combo> combos(n_combos);
        for (int i = 0; i < n_combos; i++)
            combos[i] = {combos4[4 * i] != 0.0, combos4[4 * i + 1], (float)combos4[4 * i + 2], combos4[4 * i + 3]};
        vstab::JitterScore in;
        const auto res = g->run(frames, n, row_stride, frame_stride, combos, &in);
        if (input3) { input3[0] = in.median_px; input3[1] = in.pairs; input3[2] = in.failed; }
        for (int i = 0; i < n_combos; i++) {
            results5[5 * i] = res[i].out.median_px; results5[5 * i + 1] = res[i].ratio; results5[5 * i + 2] = res[i].out.failed;
            results5[5 * i + 3] = res[i].failed_alignments; results5[5 * i + 4] = res[i].out.pairs;
        }
        if (T) std::copy(g->sweep_transforms().begin(), g->sweep_transforms().end(), T);
        if (status) std::copy(g->sweep_status().begin(), g->sweep_status().end(), status);
        return (int)g->launches();
    });
}

// ---- partitioned video: host half (no GPU) and the full worker
void* vsh_parttraj_create(int rank, int world, int width, int height, int64_t total_frames, int sub_frames, int block,
                          const vsh_stab_params* p, const char* exchange_name, int host_threads)
{
    vstab::PartitionedTrajectory* t = nullptr;
    guarded([&] {
        t = new vstab::PartitionedTrajectory(rank, world, width, height, (long)total_frames, sub_frames, block, stab_params(p),
                                             exchange_name ? exchange_name : "", host_threads);
        return 0;
    });
    return t;
}
void vsh_parttraj_destroy(void* t) { delete (vstab::PartitionedTrajectory*)t; }
int vsh_parttraj_run(void* tp, const double* meas_all, const uint8_t* ok_all, double* corrections, int64_t* frames)
{
    return guarded([&] {
        auto* t = (vstab::PartitionedTrajectory*)tp;
        t->begin_video();
        std::vector<double> T;
        std::vector<int32_t> st;
        for (size_t i = 0; i < t->subs().size(); i++) {
            const auto& s = t->subs()[i];
            T.clear(); st.clear();
            for (long f = std::max(s.a, 1L); f < s.b; f++) {
                T.insert(T.end(), meas_all + 4 * f, meas_all + 4 * f + 4);
                st.push_back(ok_all[f] ? 1 : 0);
            }
            t->submit((int)i, T.data(), st.data());
        }
        t->flush();
        t->end_video();
        for (int k = 0; k < t->output_count(); k++) {
            if (corrections) put(t->correction(k), corrections + 4 * k);
            if (frames) frames[k] = t->local_frame(t->local_of_own(k));
        }
        return t->due();
    });
}
int vsh_parttraj_output_count(void* t) { return ((vstab::PartitionedTrajectory*)t)->output_count(); }

void* vsh_partstab_create(int device, int rank, int world, int width, int height, int64_t total_frames, int sub_frames,
                          int block, const vsh_stab_params* p, const char* exchange_name, int resident, int host_threads, int lanes)
{
    vstab::PartitionedStabilizer* s = nullptr;
    guarded([&] {
        s = new vstab::PartitionedStabilizer(device, rank, world, width, height, (long)total_frames, sub_frames, block, stab_params(p),
                                             exchange_name ? exchange_name : "", resident != 0, host_threads, lanes);
        return 0;
    });
    return s;
}
void* vsh_partstab_create_nv12(int device, int rank, int world, int width, int height, int64_t total_frames, int sub_frames,
                               int block, const vsh_stab_params* p, const char* exchange_name, int resident, int host_threads, int lanes)
{
    vstab::PartitionedStabilizer* s = nullptr;
    guarded([&] {
        s = new vstab::PartitionedStabilizer(device, rank, world, width, height, (long)total_frames, sub_frames, block, stab_params(p),
                                             exchange_name ? exchange_name : "", resident != 0, host_threads, lanes, true);
        return 0;
    });
    return s;
}
void vsh_partstab_destroy(void* s) { delete (vstab::PartitionedStabilizer*)s; }
int vsh_partstab_local_count(void* s) { return ((vstab::PartitionedStabilizer*)s)->trajectory().local_count(); }
int64_t vsh_partstab_local_frame(void* s, int i) { return ((vstab::PartitionedStabilizer*)s)->trajectory().local_frame(i); }
int vsh_partstab_local_is_halo(void* s, int i) { return ((vstab::PartitionedStabilizer*)s)->trajectory().local_is_halo(i) ? 1 : 0; }
int vsh_partstab_output_count(void* s) { return ((vstab::PartitionedStabilizer*)s)->trajectory().output_count(); }
int vsh_partstab_output_frame(void* s, int k)
{
    const auto& t = ((vstab::PartitionedStabilizer*)s)->trajectory();
    return (int)t.local_frame(t.local_of_own(k));
}
int vsh_partstab_upload_resident(void* s, const uint8_t* frames, int64_t row_stride, int64_t frame_stride, int mem)
{
    return guarded([&] { ((vstab::PartitionedStabilizer*)s)->upload_resident(frames, row_stride, frame_stride, mem); return 0; });
}
int vsh_partstab_stabilize(void* s, const uint8_t* frames, int64_t row_stride, int64_t frame_stride, uint8_t* out,
                           int64_t out_frame_stride, int out_mem)
{
    return guarded([&] { return ((vstab::PartitionedStabilizer*)s)->stabilize(frames, row_stride, frame_stride, out, out_frame_stride, out_mem); });
}
int64_t vsh_partstab_records(void* sp, double* corrections, double* meas, uint8_t* ok)
{
    const auto& t = ((vstab::PartitionedStabilizer*)sp)->trajectory();
    if (corrections) for (int k = 0; k < t.output_count(); k++) put(t.correction(k), corrections + 4 * k);
    if (meas) for (size_t i = 0; i < t.measurements().size(); i++) put(t.measurements()[i], meas + 4 * i);
    if (ok) for (size_t i = 0; i < t.successes().size(); i++) ok[i] = t.successes()[i];
    return (int64_t)t.measurements().size();
}
int vsh_partstab_out_size(void* s, int* w, int* h)
{
    *w = ((vstab::PartitionedStabilizer*)s)->out_width(); *h = ((vstab::PartitionedStabilizer*)s)->out_height();
    return 0;
}
int vsh_partstab_set_lanes(void* s, int lanes) { ((vstab::PartitionedStabilizer*)s)->set_lanes(lanes); return 0; }
vs_ctx* vsh_partstab_context(void* s) { return ((vstab::PartitionedStabilizer*)s)->context(); }

vs_ctx* vsh_clipstab_context(void* c) { return ((vstab::ClipStabilizer*)c)->context(); }
vs_clip* vsh_clipstab_clip(void* c) { return ((vstab::ClipStabilizer*)c)->clip(); }

}  // extern "C"
