// stabilizer.cpp — VideoStabilizer over the GPU aligner ring.  The control flow and the
// accumulate / decay arithmetic follow reference stabilizer.cpp:9-117 step for step; what
// changes is where frames live: the reference clones every input on the host and re-reads
// it `lag` frames later for cv::warpAffine, here the frame is already resident in the
// aligner's ring and the warp (with the crop fused in) reads it there.
#include "stabilizer.hpp"

#include <algorithm>
#include <stdexcept>

#include "aligner_impl.hpp"
#include "frame_io.hpp"
#include "trajectory.hpp"

namespace vstab {

StabilizerTrajectory::StabilizerTrajectory(const VideoStabilizerParams& params)
    // upstream passes (lag, smoother_memory) as (lagBehind, lagAhead): reference stabilizer.cpp:4
    : m_params(params), m_smoother(params.lag, params.smoother_memory, params.lambda), m_chain(params)
{
}

bool StabilizerTrajectory::push(const SimilarityTransform& measurement, bool success, int frame_width, int frame_height,
                                SimilarityTransform& correction)
{
    // The smoother finalises measurement n - smoother_memory at frame n while the delay line
    // below pops measurement n - lag: the two indices differ by lag - smoother_memory frames
    // upstream, and that is reproduced here, not repaired.
    SimilarityTransform smoothed;
    if (m_params.enable_smoother) m_smoother.update(measurement, smoothed);
    m_measurements.push_back(measurement);
    const long n = m_pushes++;
    SimilarityTransform oldest;
    if (m_measurements.size() > (size_t)std::max(0, m_params.lag)) {
        oldest = m_measurements.front();
        m_measurements.pop_front();
    }
    return m_chain.step(n, success, oldest, smoothed, frame_width, frame_height, correction);
}

}  // namespace vstab

VideoStabilizer::VideoStabilizer(const VideoStabilizerParams& params) : m_params(params), m_trajectory(params)
{
    // frames n-lag .. n must be resident when frame n arrives
    aligner.impl_->capacity = std::max(2, params.lag + 2);
}

VideoStabilizer::~VideoStabilizer() = default;

cv::Mat VideoStabilizer::processFrame(const cv::Mat& inputFrame)
{
    ++m_frameIndex;
    VideoAligner::Impl& ring = *aligner.impl_;

    // A frame-size change (or a forced re-initialisation after a failure) re-creates the ring: rescue the frames still
    // waiting in it first.  ring.width / ring.height are always the true size of the frames in the ring.
    if (ring.clip && ring.width > 0 && ring.height > 0 &&
        (ring.force_reinit || inputFrame.cols != ring.width || inputFrame.rows != ring.height)) {
        for (Pending& p : m_frameBuffer) {
            if (p.generation != ring.generation || !p.host.empty()) continue;
            p.host = cv::Mat(ring.height, ring.width, CV_8UC3);
            if (vs_clip_get_bgr(ring.clip, p.slot, p.host.data) != VS_OK)
                throw std::runtime_error(std::string("VideoStabilizer: ") + vs_last_error(ring.ctx));
        }
    }

    SimilarityTransform measurement;
    const bool success = aligner.AlignNextFrame(inputFrame, measurement, m_params.aligner);

    Pending incoming;
    incoming.slot = ring.last_slot;
    incoming.generation = ring.generation;
    // device error (no ring, or this frame never reached it): keep a host copy so the stream continues
    if (!ring.clip || incoming.slot < 0) incoming.host = inputFrame.clone();
    m_frameBuffer.push_back(incoming);

    SimilarityTransform correction;
    if (!m_trajectory.push(measurement, success, inputFrame.cols, inputFrame.rows, correction)) return cv::Mat();
    if (m_frameBuffer.empty()) return cv::Mat();

    Pending oldest = m_frameBuffer.front();
    m_frameBuffer.pop_front();
    const int crop = std::max(0, m_params.crop_pixels);
    if (oldest.host.empty() && oldest.generation == ring.generation && ring.clip) {
        const int ow = ring.width - 2 * crop, oh = ring.height - 2 * crop;
        if (ow <= 0 || oh <= 0) throw std::runtime_error("VideoStabilizer: crop_pixels removes the whole frame");
#ifdef CV_COMPAT_SHIM
        // the frame the caller receives lives in a recycled page-locked buffer: the warp's output arrives by DMA
        cv::Mat out(oh, ow, CV_8UC3, nullptr, (size_t)ow * 3, vstab::pinned_frame((size_t)ow * oh * 3));
        out.data = (uint8_t*)out.owner().get();
#else
        cv::Mat out(oh, ow, CV_8UC3);       // real OpenCV: see INTEGRATION.md (cv::cuda::HostMem allocator)
#endif
        const int32_t slot = oldest.slot;
        const double T[4] = {correction.A, correction.B, correction.TX, correction.TY};
        if (vs_clip_warp(ring.clip, &slot, 1, T, VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0, crop, out.data,
                         (int64_t)ow * oh * 3, VS_MEM_HOST) != VS_OK)
            throw std::runtime_error(std::string("VideoStabilizer: ") + vs_last_error(ring.ctx));
        return out;
    }
    // frame no longer resident (size changed since it arrived): warp the host copy
    cv::Mat stabilized = warpBySimilarityTransform(oldest.host, correction);
    if (crop > 0) stabilized = stabilized(cv::Rect(crop, crop, stabilized.cols - 2 * crop, stabilized.rows - 2 * crop));
    return stabilized;
}
