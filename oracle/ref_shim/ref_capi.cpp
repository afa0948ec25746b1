// ref_capi.cpp — C interface over the reference's OWN classes, compiled unmodified from
// /root/reference (alignment.cpp, imgproc.cpp, smoother.cpp, stabilizer.cpp) against the
// compat headers.  TEST INFRASTRUCTURE ONLY (oracle/_ref/libvs_ref.so): used to validate the
// restated oracle's orchestration (state machine, std::nth_element selection, Hessian,
// Gauss-Newton loop, smoother, stabilizer glue) against the real code, and as the timed
// "reference" CPU baseline.  Only the Halide kernel math and the five OpenCV calls underneath
// are the oracle's restatement (pipelines.cpp, cv_impl.cpp).
#include <string.h>

#include "alignment.hpp"
#include "stabilizer.hpp"

namespace {

struct AlignerTap : public VideoAligner {
    using VideoAligner::CurrFrameIndex;
    using VideoAligner::KeyframeArgMaxX;
    using VideoAligner::KeyframeArgMaxY;
    using VideoAligner::KeyframeJacobianX;
    using VideoAligner::KeyframeJacobianY;
    using VideoAligner::KeyframeTileSize;
    using VideoAligner::PyramidLevels;
    using VideoAligner::ScalePyramid;
    using VideoAligner::SelectedPixelsX;
    using VideoAligner::SelectedPixelsY;
    using VideoAligner::WarpDiffX;
    using VideoAligner::WarpDiffY;
};

struct StabTap : public VideoStabilizer {
    explicit StabTap(const VideoStabilizerParams& p) : VideoStabilizer(p) {}
    using VideoStabilizer::m_accum;
};

struct vr_align_params {
    int phase_correlate;
    double phase_correlate_threshold;
    double threshold;
    float smallest_fraction;
    int max_iters;
    int pyramid_min_width;
    int pyramid_min_height;
    double max_displacement;
};

struct vr_stab_params {
    vr_align_params aligner;
    int lag;
    int smoother_memory;
    double lambda;
    int enable_smoother;
    int crop_pixels;
    double min_disp, max_disp;
    double min_decay, max_decay;
};

VideoAlignerParams to_params(const vr_align_params* p)
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

SimilarityTransform to_tf(const double* T)
{
    SimilarityTransform t;
    t.A = T[0]; t.B = T[1]; t.TX = T[2]; t.TY = T[3];
    return t;
}
void from_tf(const SimilarityTransform& t, double* T) { T[0] = t.A; T[1] = t.B; T[2] = t.TX; T[3] = t.TY; }

}  // namespace

extern "C" {

// ---- transform algebra through the reference's SimilarityTransform (imgproc.cpp:327-437)
void vr_tf_inverse(const double T[4], double out[4]) { from_tf(to_tf(T).inverse(), out); }
void vr_tf_compose(const double T1[4], const double T2[4], double out[4]) { from_tf(to_tf(T1).compose(to_tf(T2)), out); }
void vr_tf_warp(const double T[4], double px, double py, double out[2])
{
    Point p = to_tf(T).warp(Point{px, py});
    out[0] = p.x; out[1] = p.y;
}
void vr_tf_warp_center(const double T[4], double px, double py, double cx, double cy, double out[2])
{
    Point p = to_tf(T).warp(Point{px, py}, cx, cy);
    out[0] = p.x; out[1] = p.y;
}
double vr_tf_max_corner_displacement(const double T[4], double w, double h) { return to_tf(T).maxCornerDisplacement(w, h); }

// ---- VideoAligner (alignment.cpp)
void* vr_aligner_create(void) { return new AlignerTap(); }
void vr_aligner_destroy(void* a) { delete (AlignerTap*)a; }
int vr_aligner_align(void* a, const uint8_t* bgr, int w, int h, const vr_align_params* params, double T[4])
{
    cv::Mat frame(h, w, CV_8UC3, (void*)bgr);
    SimilarityTransform t;
    bool ok = ((AlignerTap*)a)->AlignNextFrame(frame, t, to_params(params));
    from_tf(t, T);
    return ok ? 1 : 0;
}
int vr_aligner_levels(void* a) { return ((AlignerTap*)a)->PyramidLevels; }
int vr_aligner_curr_index(void* a) { return ((AlignerTap*)a)->CurrFrameIndex; }
int vr_aligner_tile_size(void* a, int level) { return ((AlignerTap*)a)->KeyframeTileSize[level]; }
const uint8_t* vr_aligner_pyramid(void* a, int slot, int level, int* w, int* h)
{
    auto& b = ((AlignerTap*)a)->ScalePyramid[slot][level];
    *w = b.width(); *h = b.height();
    return b.data();
}
const uint16_t* vr_aligner_keypoints(void* a, int level, int axis, int* tw, int* th)
{
    auto& b = axis == 0 ? ((AlignerTap*)a)->KeyframeArgMaxX[level] : ((AlignerTap*)a)->KeyframeArgMaxY[level];
    *tw = b.width(); *th = b.height();
    return b.data();
}
const float* vr_aligner_jacobians(void* a, int level, int axis)
{
    auto& b = axis == 0 ? ((AlignerTap*)a)->KeyframeJacobianX[level] : ((AlignerTap*)a)->KeyframeJacobianY[level];
    return b.data();
}
const uint16_t* vr_aligner_warpdiff(void* a, int level, int axis)
{
    auto& b = axis == 0 ? ((AlignerTap*)a)->WarpDiffX[level] : ((AlignerTap*)a)->WarpDiffY[level];
    return b.data();
}
// selected keypoint coordinates in post-nth_element order: planar (k,2) u16
const uint16_t* vr_aligner_selected_pixels(void* a, int level, int axis, int* k)
{
    auto& b = axis == 0 ? ((AlignerTap*)a)->SelectedPixelsX[level] : ((AlignerTap*)a)->SelectedPixelsY[level];
    *k = b.dimensions() == 2 ? b.dim(0).extent() : 0;
    return b.data();
}

// ---- L1SmootherCenter (smoother.cpp)
void* vr_smoother_create(int lag_behind, int lag_ahead, double lambda) { return new L1SmootherCenter(lag_behind, lag_ahead, lambda); }
void vr_smoother_destroy(void* s) { delete (L1SmootherCenter*)s; }
int vr_smoother_update(void* s, const double meas[4], double out[4])
{
    SimilarityTransform o;
    bool r = ((L1SmootherCenter*)s)->update(to_tf(meas), o);
    from_tf(o, out);
    return r ? 1 : 0;
}

// ---- VideoStabilizer (stabilizer.cpp)
void* vr_stabilizer_create(const vr_stab_params* p)
{
    VideoStabilizerParams q;
    if (p) {
        q.aligner = to_params(&p->aligner);
        q.lag = p->lag; q.smoother_memory = p->smoother_memory; q.lambda = p->lambda;
        q.enable_smoother = p->enable_smoother != 0; q.crop_pixels = p->crop_pixels;
        q.min_disp = p->min_disp; q.max_disp = p->max_disp; q.min_decay = p->min_decay; q.max_decay = p->max_decay;
    }
    return new StabTap(q);
}
void vr_stabilizer_destroy(void* s) { delete (StabTap*)s; }
// returns 1 and fills out (dense out_w*out_h*3) when a stabilized frame is produced, else 0;
// accum receives m_accum after the call
int vr_stabilizer_process(void* s, const uint8_t* bgr, int w, int h, uint8_t* out, int* out_w, int* out_h, double accum[4])
{
    cv::Mat frame(h, w, CV_8UC3, (void*)bgr);
    cv::Mat r = ((StabTap*)s)->processFrame(frame);
    from_tf(((StabTap*)s)->m_accum, accum);
    if (r.empty()) { *out_w = *out_h = 0; return 0; }
    *out_w = r.cols; *out_h = r.rows;
    for (int y = 0; y < r.rows; y++) memcpy(out + (size_t)y * r.cols * 3, r.ptr(y), (size_t)r.cols * 3);
    return 1;
}

}  // extern "C"
