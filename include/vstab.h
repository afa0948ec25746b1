/*
 * vstab.h — C ABI of libvstab.so: B200 (sm_100a) kernels for the alignment-and-warp
 * hot path of catid/video_stabilizer.
 *
 * This is the boundary a maintainer of the reference binds against: every entry point
 * below replaces one Halide AOT pipeline (generators.cpp), one OpenCV call or one block of
 * host orchestration on the path.  Citations are file:line in the reference tree.
 * INTEGRATION.md shows the reference-side wrappers (imgproc.cpp / alignment.cpp /
 * stabilizer.cpp) rewritten against this header.
 *
 * Conventions
 *  - every function returns 0 (VS_OK) or a negative vs_status; vs_last_error() gives text.
 *  - no exceptions, no C++ types, no torch types cross this boundary.
 *  - `mem` says where the data pointers of that call live: VS_MEM_HOST (the call copies
 *    in, runs, copies out and returns when the result is in host memory) or VS_MEM_DEVICE
 *    (pointers are device pointers, the call only enqueues work on the context's stream).
 *  - images are row-major; `stride` is in ELEMENTS of the image type between rows;
 *    `batch` images are `batch_stride` elements apart.  One launch processes the batch.
 *  - "planar (w,h,c)" arrays are Halide::Runtime::Buffer<T>(w,h,c) dense layout:
 *    index = (c*h + y)*w + x.
 *  - there is no CPU fallback anywhere in this library: without a CUDA device every
 *    compute entry point fails with VS_ERR_CUDA.
 */
#ifndef VSTAB_H
#define VSTAB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 1
#define VS_MAX_LEVELS 12

typedef enum vs_status {
    VS_OK = 0,
    VS_ERR_INVALID = -1,     /* bad argument */
    VS_ERR_CUDA = -2,        /* CUDA runtime / launch failure, or no device */
    VS_ERR_UNSUPPORTED = -3, /* valid request this build does not implement */
    VS_ERR_NOMEM = -4
} vs_status;

enum { VS_MEM_HOST = 0, VS_MEM_DEVICE = 1 };

/* BGR warp interpolation / border (imgproc.cpp:446-484 uses CV_EXACT + CONSTANT0) */
enum { VS_WARP_CV_EXACT_BILINEAR = 0, VS_WARP_FLOAT_BILINEAR = 1, VS_WARP_LANCZOS2 = 2 };
enum { VS_BORDER_CONSTANT0 = 0, VS_BORDER_REPEAT_EDGE = 1 };

typedef struct vs_ctx vs_ctx;
typedef struct vs_clip vs_clip;

typedef struct vs_img {
    void*   data;
    int32_t width, height;
    int64_t stride;        /* elements between rows (bytes for u8; BGR counts bytes too) */
    int32_t batch;         /* >= 1 */
    int64_t batch_stride;  /* elements between images of the batch */
} vs_img;

/* ------------------------------------------------------------------ context */
int vs_abi_version(void);
int vs_device_count(void);
/* One context per GPU and per host thread that drives it.  Owns a stream and scratch. */
int vs_ctx_create(int device, vs_ctx** out);
int vs_ctx_destroy(vs_ctx* ctx);
/* Borrow an external cudaStream_t (e.g. torch's current stream) so that the caller's
 * CUDA events see this library's launches.  NULL restores the context's own stream. */
int vs_ctx_set_stream(vs_ctx* ctx, void* cuda_stream);
int vs_ctx_synchronize(vs_ctx* ctx);
/* text of the last error on this context (or of the last failed vs_ctx_create if NULL) */
const char* vs_last_error(const vs_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
int64_t vs_ctx_launch_count(const vs_ctx* ctx);
/* Per-kernel device timing: when enabled every launch of this context is bracketed by a
 * CUDA event pair on the launching stream.  read() synchronizes the stream and returns the
 * launches and summed milliseconds of one kernel since the last reset. */
enum { VS_KERNEL_BGR2GRAY = 0, VS_KERNEL_PYR_DOWN, VS_KERNEL_GRAD_XY, VS_KERNEL_IMAGE_WARP, VS_KERNEL_BGR_WARP,
       VS_KERNEL_GRAD_ARGMAX, VS_KERNEL_SPARSE_JAC, VS_KERNEL_WARPDIFF, VS_KERNEL_ICA, VS_KERNEL_KEYFRAME,
       VS_KERNEL_SOLVE, VS_KERNEL_INGEST, VS_KERNEL_COUNT };
int vs_ctx_profile_enable(vs_ctx* ctx, int enable);
int vs_ctx_profile_reset(vs_ctx* ctx);
int vs_ctx_profile_read(vs_ctx* ctx, int kernel, int64_t* launches, double* total_ms);
/* Timeline of the launches recorded since the last reset / read, oldest first: out3[i] = {kernel id, start ms, end ms}
 * relative to the first launch's start (the events of launches on the clip's solver lanes included).  Returns how many.
 * Call before vs_ctx_profile_read (which consumes the records). */
int vs_ctx_profile_timeline(vs_ctx* ctx, double* out3, int capacity);
const char* vs_kernel_name(int kernel);
/* device memory helpers for callers without their own allocator (C++ host layer) */
int vs_dev_alloc(vs_ctx* ctx, size_t bytes, void** out);
int vs_dev_free(vs_ctx* ctx, void* p);
int vs_host_alloc_pinned(vs_ctx* ctx, size_t bytes, void** out);
int vs_host_free_pinned(vs_ctx* ctx, void* p);
/* page-locked host memory that is not tied to a context (frame pools that outlive the stabilizer that filled them), and
 * whether a host pointer is page-locked (cudaMallocHost / cudaHostRegister): copies from and to such memory are DMA
 * transfers at the PCIe rate, copies of pageable memory are staged by the driver */
int vs_pinned_alloc(size_t bytes, void** out);
int vs_pinned_free(void* p);
int vs_host_is_pinned(const void* p);
int vs_memcpy_h2d(vs_ctx* ctx, void* dst, const void* src, size_t bytes);  /* async on stream */
int vs_memcpy_d2h(vs_ctx* ctx, void* dst, const void* src, size_t bytes);  /* async on stream */

/* ------------------------------------------------- single operators (1:1) */
/* cv::cvtColor(BGR2GRAY), alignment.cpp:212.  bgr: interleaved, width in pixels,
 * stride in bytes.  gray: u8. */
int vs_bgr2gray_u8(vs_ctx*, const vs_img* bgr, const vs_img* gray, int mem);

/* Fused ingest: cv::cvtColor(BGR2GRAY) (alignment.cpp:212) and the first PyrDown
 * (alignment.cpp:220-223) in ONE pass over the interleaved frame: gray0 = BGR2GRAY(bgr), gray1 = pyr_down(gray0),
 * bit-identical to vs_bgr2gray_u8 followed by vs_pyr_down_u8.  The single fused kernel runs when width % 16 == 0,
 * gray1 is (width/2) x (height/2) and the rows are 16-byte aligned; other geometries run the two kernels. */
int vs_ingest_bgr_u8(vs_ctx*, const vs_img* bgr, const vs_img* gray0, const vs_img* gray1, int mem);

/* pyr_down(), generators.cpp:56-92 via PyrDown(), imgproc.cpp:108-114.  The output
 * extent defines the work; input is read through repeat-edge. */
int vs_pyr_down_u8(vs_ctx*, const vs_img* in, const vs_img* out, int mem);

/* grad_xy(), generators.cpp:202-224 via GradXY(), imgproc.cpp:135-142. */
int vs_grad_xy_u8_f32(vs_ctx*, const vs_img* in, const vs_img* grad_x, const vs_img* grad_y, int mem);

/* tile-size rule of GradArgMax(), imgproc.cpp:151-162 */
int vs_grad_argmax_tile_size(int width, int height);
/* grad_argmax_N(), generators.cpp:260-294 via GradArgMax(), imgproc.cpp:144-202.
 * local_max_x/y: u16 planar (tw,th,2), tw = width/tile, th = height/tile. */
int vs_grad_argmax_f32_u16(vs_ctx*, const vs_img* grad_x, const vs_img* grad_y, int tile,
                           uint16_t* local_max_x, uint16_t* local_max_y, int mem);

/* sparse_jac(), generators.cpp:332-386 via SparseJacobian(), imgproc.cpp:26-44.
 * local_max_*: u16 planar (tw,th,2); out_*: f32 planar (tw,th,4). */
int vs_sparse_jac_f32(vs_ctx*, const vs_img* grad_x, const vs_img* grad_y,
                      const uint16_t* local_max_x, const uint16_t* local_max_y,
                      int tw, int th, float* out_x, float* out_y, int mem);

/* sparse_warpdiff(), generators.cpp:646-700.  A,B,TX,TY are the pipeline's f32
 * upper-left-origin parameters (SparseWarpDiff(), imgproc.cpp:94-104, computes them).
 * local_max: u16 planar (tw,th,2); out: u16 (tw,th). */
int vs_sparse_warpdiff_u8_u16(vs_ctx*, const vs_img* tmpl, const vs_img* keyframe,
                              const uint16_t* local_max, int tw, int th,
                              float A, float B, float TX, float TY, uint16_t* out, int mem);

/* sparse_ica(), generators.cpp:429-596.  selected_*: u16 planar (k,2); jac_*: f32 planar
 * (k,4); out: 4 doubles (SparseICA(), imgproc.cpp:46-78). */
int vs_sparse_ica_f64(vs_ctx*, const vs_img* tmpl, const vs_img* keyframe,
                      const uint16_t* selected_x, int kx, const uint16_t* selected_y, int ky,
                      const float* jac_x, const float* jac_y,
                      float A, float B, float TX, float TY, double* out4, int mem);

/* image_warp(), generators.cpp:126-164 (ImageWarp(), imgproc.cpp:116-133 computes the
 * f32 parameters).  params: 4 floats {A,B,TX,TY} per batch image. */
int vs_image_warp_u8_f32(vs_ctx*, const vs_img* in, const float* params4, const vs_img* out, int mem);

/* The BGR warp of warpBySimilarityTransform(), imgproc.cpp:446-484:
 * cv::warpAffine(src, dst, M, size, INTER_LINEAR, BORDER_CONSTANT, 0) without
 * WARP_INVERSE_MAP.  M: 6 doubles (row-major 2x3 forward matrix) per batch image, always
 * in host memory.  dst may be a cropped window: dst pixel (x,y) is output pixel
 * (x+dst_x0, y+dst_y0) of the full-size warp (stabilizer.cpp:102-109 crop fused). */
int vs_bgr_warp_u8(vs_ctx*, const vs_img* src, const double* M6, const vs_img* dst,
                   int dst_x0, int dst_y0, int mode, int border, int mem);
/* The same warp for the planes of an NV12 frame (no counterpart upstream, whose frames are BGR cv::Mat: the data format
 * a hardware decoder delivers): cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of a 1-channel
 * (Y) or 2-channel (interleaved UV) u8 image.  width in pixels, stride in bytes. */
int vs_plane_warp_u8(vs_ctx*, const vs_img* src, int channels, const double* M6, const vs_img* dst,
                     int dst_x0, int dst_y0, int mem);
/* cv::phaseCorrelate(src1, src2, noArray(), &response) (alignment.cpp:374, align_test.cpp:190,386) on two u8 images of
 * one size taken as CV_32F: out3 = shift x, shift y, response (host memory).  Two-stage f64 DFT on the device. */
int vs_phase_correlate_u8(vs_ctx*, const vs_img* src1, const vs_img* src2, double* out3, int mem);

/* ------------------------------------------- fused, batched, device-resident */
/* VideoAlignerParams, alignment.hpp:5-41 */
typedef struct vs_align_params {
    int32_t phase_correlate;          /* alignment.cpp:369-388: seed the translation from the phase correlation of the level-2 images */
    double  phase_correlate_threshold;
    double  threshold;
    float   smallest_fraction;
    int32_t max_iters;
    int32_t pyramid_min_width;
    int32_t pyramid_min_height;
    double  max_displacement;
} vs_align_params;
void vs_align_params_default(vs_align_params* p);

/* One alignment job: which frame slot is the template, which is the keyframe, and
 * whether the result is inverted (alignment.cpp:396-397, :690-693). */
typedef struct vs_pair {
    int32_t template_slot;
    int32_t keyframe_slot;
    int32_t invert;
} vs_pair;

enum { VS_CLIP_DEBUG_TAPS = 1,    /* keep per-pair warpdiff / selection for inspection */
       VS_CLIP_NV12 = 2 };        /* frames are NV12 instead of BGR (below) */

/* A clip is `capacity` frame slots of one size resident on the GPU: BGR frame, gray
 * pyramid (ComputePyramid, alignment.cpp:149-235) and keyframe features
 * (ComputeKeyFrame, alignment.cpp:237-276).  VideoAligner uses 2 slots, the batched
 * pipeline uses one slot per frame of a chunk.  max_pairs bounds one vs_clip_align call. */
/* VS_CLIP_NV12 (the output of a hardware decoder fed straight in; no counterpart upstream,
 * whose frames are BGR cv::Mat): every frame the clip takes or returns is an NV12 frame: `height` rows of Y, then
 * height/2 rows of interleaved UV, all row_stride bytes apart on input, dense (width bytes per row, 3/2 width height
 * bytes per frame) on output; width, height and crop even.  Every `bgr` / `out` pointer below then means such a frame.
 * The Y plane IS the gray image (cv::cvtColor(COLOR_YUV2GRAY_NV12) copies it): it is uploaded into level 0 of the
 * pyramid, nothing is converted, and the alignment is that of a BGR clip whose three channels equal Y.  Frames are warped
 * plane by plane with the cv-exact bilinear mode and the constant border only: Y by the correction itself, UV as a
 * (width/2) x (height/2) two-channel image by the same similarity with half the translation. */
int vs_clip_create(vs_ctx*, int width, int height, int capacity, int max_pairs,
                   const vs_align_params* params, int flags, vs_clip** out);
int vs_clip_destroy(vs_clip*);
/* Replace the solver parameters (threshold, smallest_fraction, max_iters, max_displacement)
 * for later vs_clip_align calls: AlignNextFrame takes its params per call (alignment.hpp:55-58).
 * pyramid_min_* only act when the pyramid is laid out, i.e. at creation, as upstream
 * (alignment.cpp:155-169). */
int vs_clip_set_params(vs_clip*, const vs_align_params* params);
int vs_clip_levels(const vs_clip*);
int vs_clip_level_info(const vs_clip*, int level, int* w, int* h, int* tile, int* tw, int* th);

/* copy n BGR frames (interleaved u8, row_stride bytes between rows, frame_stride bytes
 * between frames) into slots [slot0, slot0+n) */
int vs_clip_upload(vs_clip*, int slot0, int n, const uint8_t* bgr,
                   int64_t row_stride, int64_t frame_stride, int mem);
/* BGR -> gray -> full pyramid for slots [slot0, slot0+n) */
int vs_clip_build_pyramids(vs_clip*, int slot0, int n);
/* fused grad_xy + grad_argmax + sparse_jac on every level, for the listed slots (host array) */
int vs_clip_build_keyframes(vs_clip*, const int32_t* slots, int n);
/* AlignNextFrame's per-level loop (alignment.cpp:390-693) for n pairs in one launch.
 * pairs: host array.  out_transform: 4 doubles per pair {A,B,TX,TY}; out_status: 1 = ok,
 * 0 = false (non-convergence / over-displacement); out_iters: levels ints per pair
 * (may be NULL).  Outputs live in `mem`. */
int vs_clip_align(vs_clip*, const vs_pair* pairs, int n,
                  double* out_transform, int32_t* out_status, int32_t* out_iters, int mem);
/* One parameter combination of a sweep over VideoAlignerParams (grid_search_align.cpp:134-146 sweeps phase_correlate,
 * threshold, smallest_fraction and max_displacement). */
typedef struct vs_sweep_params {
    double  threshold;
    double  max_displacement;
    float   smallest_fraction;
    int32_t max_iters;
    int32_t phase_correlate;
} vs_sweep_params;
/* The alignment of n_pairs pairs under n_sets parameter combinations in ONE launch (n_pairs * n_sets <= max_pairs):
 * out_transform [n_sets][n_pairs][4], out_status [n_sets][n_pairs], in `mem`.  Each (set, pair) result is the one
 * vs_clip_align gives after vs_clip_set_params with that combination.  phase_correlate_threshold is the clip's. */
int vs_clip_align_sweep(vs_clip*, const vs_pair* pairs, int n_pairs, const vs_sweep_params* sets, int n_sets,
                        double* out_transform, int32_t* out_status, int mem);
#define VS_CLIP_SOLVER_LANES 6
/* The same solve, enqueued on one of the clip's solver streams (lane 0 .. VS_CLIP_SOLVER_LANES-1) behind everything
 * enqueued on the context stream so far, without waiting: the pyramids / keyframe features of the next frames and the warps of frames
 * already decided then run beside it.  The pairs of one call use the per-pair scratch [base, base + n): calls in
 * flight together must not overlap there (base + n <= max_pairs).  vs_clip_align_wait blocks until the call of that
 * lane has finished and copies its transforms / status to host memory. */
int vs_clip_align_async(vs_clip*, const vs_pair* pairs, int n, int base, int lane);
int vs_clip_align_wait(vs_clip*, int lane, double* out_transform, int32_t* out_status);
/* warpBySimilarityTransform (imgproc.cpp:446-484) + crop (stabilizer.cpp:102-109) for the
 * listed slots.  transforms: 4 doubles per frame (centre-based correction), host array.
 * out: (w-2crop)x(h-2crop) BGR frames, dense, out_frame_stride bytes apart, in `mem`. */
int vs_clip_warp(vs_clip*, const int32_t* slots, int n, const double* transforms,
                 int mode, int border, int crop, uint8_t* out, int64_t out_frame_stride, int mem);

/* ---- asynchronous host <-> device transfers for pipelined feeding (ClipStabilizer::feed with
 * host buffers): uploads run on the clip's copy-in stream and downloads on its copy-out stream,
 * so PCIe traffic in both directions overlaps the kernels of neighbouring chunks.  Host buffers
 * should be pinned (vs_host_alloc_pinned / cudaHostRegister) for the copies to be truly
 * asynchronous; pageable memory works but serialises. */
/* like vs_clip_upload(VS_MEM_HOST) on the copy-in stream; ordered after all compute enqueued so far */
int vs_clip_upload_async(vs_clip*, int slot0, int n, const uint8_t* bgr, int64_t row_stride, int64_t frame_stride);
/* compute enqueued after this call waits for every upload issued so far */
int vs_clip_wait_uploads(vs_clip*);
/* like vs_clip_warp(VS_MEM_HOST) but returns as soon as the warp is enqueued; the frames land in
 * `out` on the copy-out stream.  out must stay valid until vs_clip_sync_transfers(). */
int vs_clip_warp_to_host_async(vs_clip*, const int32_t* slots, int n, const double* transforms,
                               int mode, int border, int crop, uint8_t* out, int64_t out_frame_stride);
/* blocks until every asynchronous upload and download issued so far has completed */
int vs_clip_sync_transfers(vs_clip*);

/* inspection taps (host outputs, synchronous) used by the bit-exact parity tests */
int vs_clip_get_bgr(vs_clip*, int slot, uint8_t* out /* w*h*3 dense; NV12 clips: the w*h*3/2 frame */);
int vs_clip_get_gray(vs_clip*, int slot, int level, uint8_t* out /* w*h dense */);
int vs_clip_get_keypoints(vs_clip*, int slot, int level, int axis, uint16_t* out /* planar (tw,th,2) */);
int vs_clip_get_jacobians(vs_clip*, int slot, int level, int axis, float* out /* planar (tw,th,4) */);
/* need VS_CLIP_DEBUG_TAPS; pair = index within the last vs_clip_align call */
int vs_clip_get_warpdiff(vs_clip*, int pair, int level, int axis, uint16_t* out /* (tw,th) */);
int vs_clip_get_selected(vs_clip*, int pair, int level, int axis, uint32_t* out_order, int* out_k);
/* cv::phaseCorrelate result (shift x, shift y, response) of pair `pair` of the last vs_clip_align call made with
 * params.phase_correlate set (alignment.cpp:374) */
int vs_clip_get_phase(vs_clip*, int pair, double* out3);
/* SM cycles pair `pair` of the last vs_clip_align spent per phase, summed over levels:
 * {warp-diff, selection, Hessian sums, SVD beside the first iteration + Gauss-Newton gathers, Gauss-Newton reduce + update,
 *  number of parallel partition rounds of the selection (a count, not cycles)} */
int vs_clip_get_solver_cycles(vs_clip*, int pair, long long* out6);
/* The same per pyramid level: out[level][8] = {warp-diff, selection, Hessian sums, SVD beside the first iteration,
 * Gauss-Newton reduce + update, partition rounds (a count), gathers of the later iterations, 0} */
int vs_clip_get_solver_level_cycles(vs_clip*, int pair, long long* out_levels_x8);
/* Debug tap for the solver's 4x4 conditioning + SVD pseudo-inverse (alignment.cpp:554-583) as it runs on the device:
 * n row-major 4x4 f64 matrices from host memory -> the inverse computed by the lane-parallel form the solver uses
 * (out_quad) and by the serial restatement (out_serial), plus the condition number; the two must agree bit for bit. */
int vs_debug_invert4(vs_ctx*, const double* H, int n, double* out_quad, double* out_serial, double* out_cond);

#ifdef __cplusplus
}
#endif
#endif /* VSTAB_H */
