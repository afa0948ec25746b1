// alignment.cpp — VideoAligner on the device-resident clip API (vs_clip_*, include/vstab.h).
// Mirrors the frame protocol of reference alignment.cpp:149-235 (size-change reset, two-slot
// alternation, first frame returns false) and :334-704 (keyframe on odd frames, template /
// keyframe roles, inversion when the current frame is not the keyframe); the per-level work
// itself is one kernel launch (vs_clip_align).
#include "aligner_impl.hpp"

#include "frame_io.hpp"

#include <stdexcept>

namespace vstab {

void to_c_params(const VideoAlignerParams& p, vs_align_params* out)
{
    vs_align_params_default(out);
    out->phase_correlate = p.phase_correlate ? 1 : 0;
    out->phase_correlate_threshold = p.phase_correlate_threshold;
    out->threshold = p.threshold;
    out->smallest_fraction = p.smallest_fraction;
    out->max_iters = p.max_iters;
    out->pyramid_min_width = p.pyramid_min_width;
    out->pyramid_min_height = p.pyramid_min_height;
    out->max_displacement = p.max_displacement;
}

}  // namespace vstab

VideoAligner::Impl::~Impl()
{
    destroy_clip();
    if (ctx) vs_ctx_destroy(ctx);
    if (staging) vs_pinned_free(staging);
}

int VideoAligner::Impl::upload(int slot, const uint8_t* data, size_t step, int w, int h)
{
    const size_t row_bytes = (size_t)w * 3;
    if (vs_host_is_pinned(data))
        return vs_clip_upload(clip, slot, 1, data, (int64_t)step, (int64_t)step * h, VS_MEM_HOST);
    if (staging_bytes < row_bytes * h) {
        // the previous frame's DMA out of the old buffer has finished: every call ends with a synchronisation
        if (staging) vs_pinned_free(staging);
        staging = nullptr;
        staging_bytes = 0;
        void* p = nullptr;
        if (vs_pinned_alloc(row_bytes * h, &p) != VS_OK) return VS_ERR_NOMEM;
        staging = (uint8_t*)p;
        staging_bytes = row_bytes * h;
    }
    vstab::copy_rows_parallel(staging, row_bytes, data, step, row_bytes, h);
    return vs_clip_upload(clip, slot, 1, staging, (int64_t)row_bytes, (int64_t)row_bytes * h, VS_MEM_HOST);
}

void VideoAligner::Impl::ensure_context()
{
    if (ctx) return;
    // one context (stream + scratch) per aligner: instances on different host threads are
    // independent, as in the reference's one-stabilizer-per-worker usage
    int dev = device;
    if (dev < 0) {
        const char* e = getenv("VSTAB_DEVICE");
        dev = e ? atoi(e) : 0;
    }
    if (vs_ctx_create(dev, &ctx) != VS_OK)
        throw std::runtime_error(std::string("VideoAligner: cannot create a GPU context: ") + vs_last_error(nullptr));
    device = dev;
}

void VideoAligner::Impl::destroy_clip()
{
    if (clip) vs_clip_destroy(clip);
    clip = nullptr;
}

bool VideoAligner::Impl::ensure_clip(int w, int h, const VideoAlignerParams& params)
{
    vs_align_params cp;
    vstab::to_c_params(params, &cp);
    if (clip && !force_reinit && w == width && h == height) return vs_clip_set_params(clip, &cp) == VS_OK;
    destroy_clip();
    force_reinit = false;
    width = w; height = h;
    frames_since_reset = 0;
    last_slot = -1;
    generation++;
    if (vs_clip_create(ctx, w, h, capacity, 1, &cp, 0, &clip) != VS_OK) {
        std::cerr << "VideoAligner: " << vs_last_error(ctx) << std::endl;
        clip = nullptr;
        width = height = -1;
        return false;
    }
    return true;
}

VideoAligner::VideoAligner() : impl_(new Impl()) {}
VideoAligner::~VideoAligner() = default;
VideoAligner::VideoAligner(VideoAligner&&) noexcept = default;
VideoAligner& VideoAligner::operator=(VideoAligner&&) noexcept = default;

bool VideoAligner::AlignNextFrame(const cv::Mat& frame, SimilarityTransform& transform, const VideoAlignerParams& params)
{
    transform = SimilarityTransform();
    if (frame.empty() || frame.type() != CV_8UC3)
        throw std::runtime_error("VideoAligner::AlignNextFrame: frame must be a non-empty CV_8UC3 (BGR) image");
    Impl& s = *impl_;
    s.ensure_context();
    if (!s.ensure_clip(frame.cols, frame.rows, params)) return false;

    const long n = s.frames_since_reset;
    const int slot = s.slot_of(n);
    const bool is_keyframe = (n & 1) != 0;
    if (s.upload(slot, frame.data, (size_t)frame.step[0], frame.cols, frame.rows) != VS_OK ||
        vs_clip_build_pyramids(s.clip, slot, 1) != VS_OK) {
        std::cerr << "VideoAligner: " << vs_last_error(s.ctx) << std::endl;
        s.last_slot = -1;          // this frame is not in the ring: VideoStabilizer keeps a host copy of it
        s.force_reinit = true;
        return false;
    }
    const int prev = s.last_slot;
    s.last_slot = slot;
    s.frames_since_reset = n + 1;
    if (n == 0) {
        vs_ctx_synchronize(s.ctx);   // the caller may reuse `frame` as soon as we return
        return false;                // no predecessor yet
    }

    if (is_keyframe) {
        const int32_t ks = slot;
        if (vs_clip_build_keyframes(s.clip, &ks, 1) != VS_OK) {
            std::cerr << "VideoAligner: " << vs_last_error(s.ctx) << std::endl;
            s.force_reinit = true;     // force re-initialisation, like upstream's LastWidth = -1 (the ring keeps its true size)
            return false;
        }
    }

    vs_pair pair;
    if (is_keyframe) { pair.template_slot = prev; pair.keyframe_slot = slot; pair.invert = 0; }
    else             { pair.template_slot = slot; pair.keyframe_slot = prev; pair.invert = 1; }
    double T[4] = {0, 0, 0, 0};
    int32_t status = 0;
    if (vs_clip_align(s.clip, &pair, 1, T, &status, nullptr, VS_MEM_HOST) != VS_OK) {
        std::cerr << "VideoAligner: " << vs_last_error(s.ctx) << std::endl;
        return false;
    }
    transform.A = T[0]; transform.B = T[1]; transform.TX = T[2]; transform.TY = T[3];
    return status != 0;
}
