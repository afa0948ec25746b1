/*
 * vstab_host.h — C view of libvstab_host.so, the C++ host layer that keeps the reference's
 * public classes (VideoAligner, VideoStabilizer, L1SmootherCenter, SimilarityTransform and
 * the imgproc.hpp operators) on top of the CUDA C ABI in vstab.h.
 *
 * C++ callers use the classes directly (video_stabilizer_b200/csrc/host/*.hpp: same names
 * and signatures as the reference's alignment.hpp / imgproc.hpp / stabilizer.hpp).  This
 * header exists for non-C++ callers — the Python tests and bench.py bind it with ctypes —
 * and every function is a thin forwarder to the class method it names.
 *
 * Status: functions returning int give 0 on success, -1 after catching a C++ exception
 * (text in vsh_last_error()), unless documented otherwise.
 */
#ifndef VSTAB_HOST_H
#define VSTAB_HOST_H

#include <stdint.h>

#include "vstab.h"

#ifdef __cplusplus
extern "C" {
#endif

/* VideoStabilizerParams (stabilizer.hpp:13-30) with VideoAlignerParams inline */
typedef struct vsh_stab_params {
    vs_align_params aligner;
    int32_t lag;
    int32_t smoother_memory;
    double  lambda;
    int32_t enable_smoother;
    int32_t crop_pixels;
    double  min_disp, max_disp;
    double  min_decay, max_decay;
} vsh_stab_params;
void vsh_stab_params_default(vsh_stab_params* p);

const char* vsh_last_error(void);

/* ---- SimilarityTransform (imgproc.hpp:40-65); T = {A,B,TX,TY} */
void   vsh_tf_inverse(const double T[4], double out[4]);
void   vsh_tf_compose(const double T1[4], const double T2[4], double out[4]);   /* T1 first */
void   vsh_tf_warp(const double T[4], double px, double py, double out[2]);
void   vsh_tf_warp_center(const double T[4], double px, double py, double cx, double cy, double out[2]);
double vsh_tf_max_corner_displacement(const double T[4], double w, double h);

/* ---- imgproc.hpp operators on dense host arrays, through the C++ wrappers
 * (Halide::Runtime::Buffer marshalling included).  Return 1 = true, 0 = false, -1 = threw. */
int vsh_PyrDown(const uint8_t* in, int iw, int ih, uint8_t* out, int ow, int oh);
int vsh_GradXY(const uint8_t* in, int w, int h, float* gx, float* gy);
/* lmx/lmy sized for the tile size this returns through *tile: planar (w/tile, h/tile, 2) */
int vsh_GradArgMax(const float* gx, const float* gy, int w, int h, int* tile, uint16_t* lmx, uint16_t* lmy, int capacity);
int vsh_SparseJacobian(const float* gx, const float* gy, int w, int h, const uint16_t* lmx, const uint16_t* lmy,
                       int tw, int th, float* jx, float* jy);
int vsh_SparseWarpDiff(const uint8_t* tmpl, const uint8_t* key, int w, int h, const uint16_t* lm, int tw, int th,
                       const double T[4], uint16_t* out);
int vsh_SparseICA(const uint8_t* tmpl, const uint8_t* key, int w, int h, const uint16_t* selx, int kx,
                  const uint16_t* sely, int ky, const float* jx, const float* jy, const double T[4], double out[4]);
int vsh_ImageWarp(const uint8_t* in, int w, int h, const double T[4], float* out, int ow, int oh);
int vsh_warpBySimilarityTransform(const uint8_t* bgr, int w, int h, int64_t row_stride, const double T[4], uint8_t* out);

/* ---- L1SmootherCenter (smoother.hpp) */
void* vsh_smoother_create(int lag_behind, int lag_ahead, double lambda);
void  vsh_smoother_destroy(void*);
int   vsh_smoother_update(void*, const double meas[4], double out[4]);   /* 1 = finalized */
void  vsh_tvl1_relax(const double* data, int n, double lambda, int iterations, double* out);

/* ---- StabilizerTrajectory: the host half of VideoStabilizer (stabilizer.cpp:19-88) */
void* vsh_trajectory_create(const vsh_stab_params* p);
void  vsh_trajectory_destroy(void*);
int   vsh_trajectory_push(void*, const double meas[4], int success, int w, int h, double correction[4]);   /* 1 = frame due */

/* ---- VideoAligner (alignment.hpp:51-58).  device < 0: VSTAB_DEVICE or 0 */
void* vsh_aligner_create(int device);
void  vsh_aligner_destroy(void*);
/* AlignNextFrame: 1 = true, 0 = false, -1 = threw */
int   vsh_aligner_align(void*, const uint8_t* bgr, int w, int h, int64_t row_stride, const vs_align_params* params, double T[4]);

/* ---- VideoStabilizer (stabilizer.hpp:32-39) */
void* vsh_stabilizer_create(const vsh_stab_params* p, int device);
void  vsh_stabilizer_destroy(void*);
/* processFrame: 1 and *out_w/*out_h/out filled when a frame came back, 0 when empty, -1 = threw */
int   vsh_stabilizer_process(void*, const uint8_t* bgr, int w, int h, int64_t row_stride, uint8_t* out, int* out_w, int* out_h);

/* ---- ClipStabilizer: batched VideoStabilizer (clip_stabilizer.hpp) */
void* vsh_clipstab_create(int device, int width, int height, int chunk_frames, const vsh_stab_params* p);
/* the same for NV12 frames in and out (VS_CLIP_NV12 in vstab.h): 3/2 bytes per pixel, even width / height / crop */
void* vsh_clipstab_create_nv12(int device, int width, int height, int chunk_frames, const vsh_stab_params* p);
void  vsh_clipstab_destroy(void*);
int   vsh_clipstab_reset(void*);
/* sub-chunk (frames) of the host-to-host transfer/compute pipeline inside feed(); default 32 */
int   vsh_clipstab_set_pipeline_frames(void*, int frames);
/* pieces a device-resident chunk of >= 128 pairs is cut into, each solved on a stream of its own beside the other
 * stages (1 = every stage back to back on one stream); default 3 */
int   vsh_clipstab_set_solver_lanes(void*, int lanes);
/* returns the number of stabilized frames written to out (>= 0) or -1 */
int   vsh_clipstab_feed(void*, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, int mem,
                        uint8_t* out, int64_t out_frame_stride, int out_mem);
int   vsh_clipstab_upload_only(void*, int64_t first_frame, const uint8_t* frames, int n, int64_t row_stride,
                               int64_t frame_stride, int mem);
int   vsh_clipstab_feed_resident(void*, int n, uint8_t* out, int64_t out_frame_stride, int out_mem);
/* records of the last feed: meas (n x 4), ok (n), corrections (produced x 4); any may be NULL */
int   vsh_clipstab_last_records(void*, double* meas, uint8_t* ok, double* corrections);
int   vsh_clipstab_out_size(void*, int* w, int* h);
vs_ctx*  vsh_clipstab_context(void*);
vs_clip* vsh_clipstab_clip(void*);

/* ---- MultiGpuStabilizer: one video partitioned by frame chunk over several GPUs (multi_gpu.hpp) */
void* vsh_multigpu_create(const int32_t* devices, int n_devices, int width, int height, int max_frames, const vsh_stab_params* p);
void  vsh_multigpu_destroy(void*);
/* returns the number of stabilized frames written to out (>= 0) or -1; meas (n x 4) / ok (n) may be NULL */
int   vsh_multigpu_stabilize(void*, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride,
                             uint8_t* out, int64_t out_frame_stride, double* meas, uint8_t* ok);

/* ---- quality tooling on the batched API (grid_search.hpp; reference eval_jitter.cpp:21-75, grid_search_align.cpp:27-60,
 * 134-210).  A combination is 4 doubles {phase_correlate, threshold, smallest_fraction, max_displacement}. */
void*  vsh_gridsearch_create(int device, int width, int height, int max_frames, int max_combos, int crop_pixels);
void   vsh_gridsearch_destroy(void*);
int    vsh_gridsearch_reference_grid(double* combos4, int capacity);      /* returns 54 */
double vsh_flow_median_px(const double T[4], int w, int h);
/* jitter of n host BGR frames: out3 = {median px, pairs, pairs the aligner could not align} */
int    vsh_gridsearch_jitter(void*, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, double* out3);
/* the sweep: all pairs under all combinations in one solver launch, trajectories (smoother off, lag 1) on the host, output
 * sequences warped and scored in batched launches.  results5 per combination = {output jitter px, ratio to the input's,
 * output pairs not aligned, input pairs not aligned, output pairs}; T [combos][n-1][4] and status [combos][n-1] optional.
 * Returns the number of kernel launches so far (>= 0) or -1. */
int    vsh_gridsearch_run(void*, const uint8_t* frames, int n, int64_t row_stride, int64_t frame_stride, const double* combos4,
                          int n_combos, double* input3, double* results5, double* T, int32_t* status);

/* ---- one video partitioned by frame chunk over several workers (partitioned.hpp): one worker per GPU, processes or
 * threads; the workers share a per-frame table in POSIX shared memory (exchange_name = "/name"; rank 0 creates it).
 * Sub-chunks of sub_frames (even) frames, `block` consecutive sub-chunks per chunk, chunks round-robin over the workers. */
/* host half only (no GPU): runs one video's trajectory from a full measurement table (total x 4, total) the partitioned way;
 * writes this worker's corrections (outputs x 4) and their frame indices; returns the number of outputs or -1 */
void* vsh_parttraj_create(int rank, int world, int width, int height, int64_t total_frames, int sub_frames, int block,
                          const vsh_stab_params* p, const char* exchange_name, int host_threads);
void  vsh_parttraj_destroy(void*);
int   vsh_parttraj_output_count(void*);
int   vsh_parttraj_run(void*, const double* meas_all, const uint8_t* ok_all, double* corrections, int64_t* frames);
/* the full worker.  resident != 0: all local frames live in the ring (upload_resident, then stabilize(frames = NULL));
 * lanes: sub-chunks in flight on the GPU (0 = default) */
void* vsh_partstab_create(int device, int rank, int world, int width, int height, int64_t total_frames, int sub_frames,
                          int block, const vsh_stab_params* p, const char* exchange_name, int resident, int host_threads, int lanes);
/* the same for a video of NV12 frames (VS_CLIP_NV12 in vstab.h) */
void* vsh_partstab_create_nv12(int device, int rank, int world, int width, int height, int64_t total_frames, int sub_frames,
                               int block, const vsh_stab_params* p, const char* exchange_name, int resident, int host_threads, int lanes);
void  vsh_partstab_destroy(void*);
/* local frame list of the worker: own frames in video order, each foreign-preceded run headed by its halo frame */
int     vsh_partstab_local_count(void*);
int64_t vsh_partstab_local_frame(void*, int i);
int     vsh_partstab_local_is_halo(void*, int i);
int     vsh_partstab_output_count(void*);            /* frames stabilize() writes */
int     vsh_partstab_output_frame(void*, int k);     /* video frame index of output k */
int     vsh_partstab_upload_resident(void*, const uint8_t* frames, int64_t row_stride, int64_t frame_stride, int mem);
/* frames: the local frames (host; NULL when resident); out: output_count() dense frames; returns their number or -1 */
int     vsh_partstab_stabilize(void*, const uint8_t* frames, int64_t row_stride, int64_t frame_stride, uint8_t* out,
                               int64_t out_frame_stride, int out_mem);
/* after stabilize: corrections (outputs x 4), the measurements / status of the frames this worker has seen; returns how many */
int64_t vsh_partstab_records(void*, double* corrections, double* meas, uint8_t* ok);
/* sub-chunks in flight on the GPU, one solver lane each (default 3 resident, 2 streamed); 1 = every stage back to back */
int     vsh_partstab_set_lanes(void*, int lanes);
int     vsh_partstab_out_size(void*, int* w, int* h);
vs_ctx* vsh_partstab_context(void*);

#ifdef __cplusplus
}
#endif
#endif /* VSTAB_HOST_H */
