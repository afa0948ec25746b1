// frame_io.hpp — host-memory plumbing of the per-frame API (VideoAligner::AlignNextFrame / VideoStabilizer::processFrame,
// reference stabilizer.cpp:9-117).  The reference hands pageable cv::Mat frames in and out; copying pageable memory to or
// from the GPU is staged by the driver through one internal buffer (a few GB/s, and serialised across threads), and a fresh
// 6 MB cv::Mat per output frame is 1500 page faults.  Here
//   * an input frame is copied into a page-locked staging buffer by a few host threads (or not at all when the caller's
//     frame already is page-locked) and uploaded by DMA;
//   * output frames are cv::Mat over page-locked buffers from a recycling pool, so the warped frame is written by DMA
//     straight into the memory the caller receives, and a released frame's buffer is reused instead of re-faulted.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <memory>

namespace vstab {

// memcpy of `rows` rows of `row_bytes` bytes, spread over the process-wide copy threads when they are free
void copy_rows_parallel(uint8_t* dst, size_t dst_stride, const uint8_t* src, size_t src_stride, size_t row_bytes, int rows);

// A page-locked buffer of at least `bytes` bytes; the returned owner gives it back to the pool (which outlives every
// stabilizer: frames may be kept by the caller for as long as it likes).
std::shared_ptr<void> pinned_frame(size_t bytes);

}  // namespace vstab
