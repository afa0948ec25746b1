// ref_pipelines.h — declarations of the Halide AOT entry points the reference's imgproc.cpp
// calls (signatures as `add_halide_library` generates them from generators.cpp's Input<>/
// Output<> lists, CMakeLists.txt:194-271).  TEST INFRASTRUCTURE: in oracle/_ref these are
// defined by pipelines.cpp on top of the CPU oracle's restated kernel math, so that the
// reference's own host sources run unmodified around them.
#pragma once
#include "HalideRuntime.h"

#ifdef __cplusplus
extern "C" {
#endif
int pyr_down(struct halide_buffer_t* input, struct halide_buffer_t* output);
int image_warp(struct halide_buffer_t* input, float A, float B, float TX, float TY, struct halide_buffer_t* output);
int grad_xy(struct halide_buffer_t* input, struct halide_buffer_t* grad_x, struct halide_buffer_t* grad_y);
#define VS_DECL_ARGMAX(N) \
    int grad_argmax_##N(struct halide_buffer_t* grad_x, struct halide_buffer_t* grad_y, \
                        struct halide_buffer_t* local_max_x, struct halide_buffer_t* local_max_y);
VS_DECL_ARGMAX(2) VS_DECL_ARGMAX(4) VS_DECL_ARGMAX(6) VS_DECL_ARGMAX(8) VS_DECL_ARGMAX(10)
VS_DECL_ARGMAX(12) VS_DECL_ARGMAX(14) VS_DECL_ARGMAX(16) VS_DECL_ARGMAX(18) VS_DECL_ARGMAX(20)
#undef VS_DECL_ARGMAX
int sparse_jac(struct halide_buffer_t* grad_x, struct halide_buffer_t* grad_y, struct halide_buffer_t* local_max_x,
               struct halide_buffer_t* local_max_y, struct halide_buffer_t* output_x, struct halide_buffer_t* output_y);
int sparse_ica(struct halide_buffer_t* input_template, struct halide_buffer_t* input_keyframe,
               struct halide_buffer_t* selected_pixels_x, struct halide_buffer_t* selected_pixels_y,
               struct halide_buffer_t* selected_jacobians_x, struct halide_buffer_t* selected_jacobians_y,
               float A, float B, float TX, float TY, struct halide_buffer_t* output);
int sparse_warpdiff(struct halide_buffer_t* input_template, struct halide_buffer_t* input_keyframe,
                    struct halide_buffer_t* local_max, float A, float B, float TX, float TY,
                    struct halide_buffer_t* output);
#ifdef __cplusplus
}
#endif
