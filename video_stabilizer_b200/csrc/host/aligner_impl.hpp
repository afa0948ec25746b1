// aligner_impl.hpp — device state behind VideoAligner (private to the host library).
#pragma once

#include "alignment.hpp"
#include "vstab.h"

struct VideoAligner::Impl {
    vs_ctx* ctx = nullptr;
    vs_clip* clip = nullptr;
    int device = -1;
    int width = -1, height = -1;
    // Frames stay resident in a ring of `capacity` slots: 2 for a bare aligner, lag+2 when a
    // VideoStabilizer owns it and warps delayed frames straight from the ring.
    int capacity = 2;
    long frames_since_reset = 0;   // frame n sits in slot n % capacity; odd n are keyframes
    int last_slot = -1;
    int generation = 0;            // bumped whenever the ring is re-created
    bool force_reinit = false;     // a keyframe / device failure: the next frame re-creates the ring (upstream's LastWidth = -1)
    uint8_t* staging = nullptr;    // page-locked copy of the incoming frame when the caller's memory is pageable (frame_io.hpp)
    size_t staging_bytes = 0;

    ~Impl();
    void ensure_context();
    // (re)create the ring for this frame size; returns false on a device error
    bool ensure_clip(int w, int h, const VideoAlignerParams& params);
    void destroy_clip();
    int slot_of(long frame) const { return (int)(frame % capacity); }
    // enqueue the upload of a host frame into `slot`: DMA from the caller's memory when it is page-locked, else through the
    // page-locked staging buffer; the caller may reuse its frame as soon as this returns
    int upload(int slot, const uint8_t* data, size_t step, int w, int h);
};

namespace vstab {
void to_c_params(const VideoAlignerParams& p, vs_align_params* out);
}
