/*
 * vs_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * alignment-and-warp path (catid/video_stabilizer).  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libvstab.so) never links or calls it.
 *
 * Parity status: the reference's own Halide/OpenCV build cannot be produced in
 * this image (no Halide, no OpenCV C++ headers).  The oracle is pinned against
 *   - the reference's transform-algebra known-answer tests (align_test.cpp:261-601),
 *   - the reference's ImageWarp shift test (align_test.cpp:358-400),
 *   - OpenCV 4.13 (cv2) outputs for cvtColor / warpAffine / SVD / invert,
 *     committed as fixtures under tests/golden/,
 *   - the real libstdc++ std::nth_element (called directly, not emulated),
 *   - the reference's own alignment.cpp / smoother.cpp / stabilizer.cpp compiled
 *     from /root/reference against shim headers (oracle/_ref, see oracle/Makefile).
 * ulp-level behaviour of a real Halide/LLVM binary (FMA contraction,
 * reassociation) is "parity unpinned"; see DESIGN.md.
 *
 * All images are dense row-major (stride == width) unless a stride is given.
 * 3-D "Halide planar" arrays (w,h,c) are laid out c-major: idx = (c*h + y)*w + x.
 * Transforms are double[4] = {A, B, TX, TY} (centre-based TX,TY as in imgproc.hpp:34-65).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- element ops (generators.cpp / imgproc.cpp) ---- */
void vo_bgr2gray(const uint8_t* bgr, int w, int h, uint8_t* gray);
void vo_pyr_down(const uint8_t* in, int iw, int ih, uint8_t* out, int ow, int oh);
void vo_grad_xy(const uint8_t* in, int iw, int ih, float* gx, float* gy, int ow, int oh);
int  vo_tile_size(int w, int h);
void vo_grad_argmax(const float* gx, const float* gy, int w, int h, int tile,
                    uint16_t* lmx, uint16_t* lmy);
void vo_sparse_jac(const float* gx, const float* gy, int w, int h,
                   const uint16_t* lmx, const uint16_t* lmy, int tw, int th,
                   float* jx, float* jy);
void vo_sparse_warpdiff(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                        const uint16_t* lm, int tw, int th, const double T[4],
                        uint16_t* out);
void vo_sparse_ica(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                   const uint16_t* selx, int kx, const uint16_t* sely, int ky,
                   const float* jx, const float* jy, const double T[4], double out[4]);
void vo_image_warp(const uint8_t* in, int iw, int ih, const double T[4],
                   float* out, int ow, int oh);

/* kernel-level entry points: the f32 upper-left-origin parameters exactly as the Halide
 * pipelines receive them (used by oracle/ref_shim to stand in for the AOT pipelines) */
void vo_k_sparse_warpdiff(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                          const uint16_t* lm, int tw, int th, float A, float B, float TX, float TY,
                          uint16_t* out);
void vo_k_sparse_ica(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                     const uint16_t* selx, int kx, const uint16_t* sely, int ky,
                     const float* jx, const float* jy, float A, float B, float TX, float TY, double out[4]);
void vo_k_image_warp(const uint8_t* in, int iw, int ih, float A, float B, float TX, float TY,
                     float* out, int ow, int oh);

/* BGR warp, same convention as warpBySimilarityTransform (imgproc.cpp:446-484):
 * dst(M p) = src(p) with M built from the centre-based transform.
 * mode: 0 = OpenCV-exact fixed-point bilinear, 1 = float bilinear, 2 = Lanczos-2
 * border: 0 = constant 0, 1 = repeat edge.  crop: pixels removed on each side. */
void vo_lanczos_table(int16_t* tab);
void vo_warp_bgr_matrix(const uint8_t* src, int w, int h, const double M[6], uint8_t* dst, int ow, int oh, int dx0, int dy0,
                        int mode, int border);
void vo_warp_plane_matrix(const uint8_t* src, int w, int h, int ch, const double M[6], uint8_t* dst, int ow, int oh, int dx0, int dy0);
void vo_warp_nv12(const uint8_t* src, int w, int h, const double T[4], uint8_t* dst, int crop);
void vo_warp_bgr(const uint8_t* src, int w, int h, const double T[4],
                 uint8_t* dst, int mode, int border, int crop);

/* ---- transform algebra (imgproc.cpp:327-437) ---- */
void   vo_tf_inverse(const double T[4], double out[4]);
void   vo_tf_compose(const double T1[4], const double T2[4], double out[4]); /* T1 first, then T2 */
void   vo_tf_warp(const double T[4], double px, double py, double out[2]);
void   vo_tf_warp_center(const double T[4], double px, double py, double cx, double cy, double out[2]);
double vo_tf_max_corner_displacement(const double T[4], double w, double h);

/* ---- 4x4 f64 SVD / inverse, restating OpenCV's Jacobi SVD (alignment.cpp:558,582) ---- */
void vo_svd4(const double H[16], double w[4], double u[16], double vt[16]);
void vo_inv4_svd(const double H[16], double Hinv[16]);

/* ---- std::nth_element selection (alignment.cpp:438-486) ----
 * in: abs_delta[n] in row-major tile order. out: order[k] = tile linear indices
 * in post-nth_element order.  Returns k = (size_t)((float)n * fraction). */
int vo_select_smallest(const uint16_t* abs_delta, int n, float fraction, uint32_t* order);

/* test aid: libstdc++ std::__introselect with an explicit depth limit; order[n] = full permutation */
int vo_introselect_depth(const uint16_t* abs_delta, int n, int nth, int depth, uint32_t* order);

/* ---- aligner (alignment.cpp:149-704) ---- */
/* cv::phaseCorrelate (alignment.cpp:374) on f32 images: out = shift x, shift y, response */
int  vo_optimal_dft_size(int n);
void vo_phase_correlate(const float* img1, const float* img2, int w, int h, int64_t stride, double out[3]);
void vo_phase_correlate_u8(const uint8_t* img1, const uint8_t* img2, int w, int h, double out[3]);

typedef struct vo_align_params {
    int    phase_correlate;          /* must be 0: not restated (default off in the reference) */
    double phase_correlate_threshold;
    double threshold;
    float  smallest_fraction;
    int    max_iters;
    int    pyramid_min_width;
    int    pyramid_min_height;
    double max_displacement;
} vo_align_params;
void vo_align_params_default(vo_align_params* p);

typedef struct vo_aligner vo_aligner;
vo_aligner* vo_aligner_create(void);
void        vo_aligner_destroy(vo_aligner*);
/* returns 1 on success, 0 on false (first frame / non-convergence / over-displacement) */
int  vo_aligner_align(vo_aligner*, const uint8_t* bgr, int w, int h,
                      const vo_align_params* params, double T[4]);
/* debug taps */
void vo_aligner_phase(const vo_aligner*, double out[3]);   /* phaseCorrelate result of the last align */
int  vo_aligner_levels(const vo_aligner*);
int  vo_aligner_curr_index(const vo_aligner*);
void vo_aligner_level_info(const vo_aligner*, int level, int* w, int* h, int* tile, int* tw, int* th);
const uint8_t*  vo_aligner_pyramid(const vo_aligner*, int slot, int level);
const uint16_t* vo_aligner_keypoints(const vo_aligner*, int level, int axis);   /* planar (tw,th,2) */
const float*    vo_aligner_jacobians(const vo_aligner*, int level, int axis);   /* planar (tw,th,4) */
const uint16_t* vo_aligner_warpdiff(const vo_aligner*, int level, int axis);    /* (tw,th) of last align */
/* selected tile indices (post-nth_element order) of the last align call */
int  vo_aligner_selected(const vo_aligner*, int level, int axis, const uint32_t** order);
int  vo_aligner_iterations(const vo_aligner*, int level);  /* iterations of last align at level, 0 if not reached */

/* ---- smoother + stabilizer (smoother.cpp:18-127, stabilizer.cpp:3-117) ---- */
void vo_tvl1_smooth(const double* data, int n, double lambda, int iterations, double* out);

typedef struct vo_smoother vo_smoother;
vo_smoother* vo_smoother_create(int lag_behind, int lag_ahead, double lambda);
void vo_smoother_destroy(vo_smoother*);
int  vo_smoother_update(vo_smoother*, const double meas[4], double out[4]);

typedef struct vo_stab_params {
    vo_align_params aligner;
    int    lag;
    int    smoother_memory;
    double lambda;
    int    enable_smoother;
    int    crop_pixels;
    double min_disp, max_disp;
    double min_decay, max_decay;
} vo_stab_params;
void vo_stab_params_default(vo_stab_params* p);

typedef struct vo_stabilizer vo_stabilizer;
vo_stabilizer* vo_stabilizer_create(const vo_stab_params* p);
void vo_stabilizer_destroy(vo_stabilizer*);
/* returns 1 and fills out (out_w*out_h*3 bytes) when a stabilized frame is produced, else 0.
 * meas_ok/meas receive AlignNextFrame's result for this frame; correction receives the
 * transform passed to warpBySimilarityTransform when a frame is produced. */
int vo_stabilizer_process(vo_stabilizer*, const uint8_t* bgr, int w, int h,
                          uint8_t* out, int* out_w, int* out_h,
                          int* meas_ok, double meas[4], double correction[4]);

#ifdef __cplusplus
}
#endif
