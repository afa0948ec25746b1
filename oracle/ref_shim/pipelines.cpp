// pipelines.cpp — the Halide AOT entry points, implemented with the CPU oracle's restated
// kernel math (vs_oracle.cpp).  TEST INFRASTRUCTURE ONLY.
#include "ref_pipelines.h"

#include <vector>
#include <string.h>

#include "../vs_oracle.h"

namespace {

// dense copy of a 2-D / 3-D planar halide buffer (the reference only ever passes dense ones,
// but strides are honoured)
template <typename T>
struct Dense {
    std::vector<T> tmp;
    const T* p = nullptr;
    int w = 1, h = 1, c = 1;
    explicit Dense(const halide_buffer_t* b)
    {
        w = b->dimensions > 0 ? b->dim[0].extent : 1;
        h = b->dimensions > 1 ? b->dim[1].extent : 1;
        c = b->dimensions > 2 ? b->dim[2].extent : 1;
        bool dense = b->dimensions < 1 || b->dim[0].stride == 1;
        if (b->dimensions > 1) dense = dense && b->dim[1].stride == w;
        if (b->dimensions > 2) dense = dense && b->dim[2].stride == w * h;
        if (dense) { p = (const T*)b->host; return; }
        tmp.resize((size_t)w * h * c);
        for (int k = 0; k < c; k++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    ptrdiff_t o = (ptrdiff_t)x * b->dim[0].stride;
                    if (b->dimensions > 1) o += (ptrdiff_t)y * b->dim[1].stride;
                    if (b->dimensions > 2) o += (ptrdiff_t)k * b->dim[2].stride;
                    tmp[((size_t)k * h + y) * w + x] = ((const T*)b->host)[o];
                }
        p = tmp.data();
    }
};

template <typename T>
struct DenseOut {
    halide_buffer_t* b;
    std::vector<T> tmp;
    T* p;
    int w, h, c;
    explicit DenseOut(halide_buffer_t* buf) : b(buf)
    {
        w = b->dimensions > 0 ? b->dim[0].extent : 1;
        h = b->dimensions > 1 ? b->dim[1].extent : 1;
        c = b->dimensions > 2 ? b->dim[2].extent : 1;
        tmp.resize((size_t)w * h * c);
        p = tmp.data();
    }
    ~DenseOut()
    {
        for (int k = 0; k < c; k++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    ptrdiff_t o = (ptrdiff_t)x * b->dim[0].stride;
                    if (b->dimensions > 1) o += (ptrdiff_t)y * b->dim[1].stride;
                    if (b->dimensions > 2) o += (ptrdiff_t)k * b->dim[2].stride;
                    ((T*)b->host)[o] = tmp[((size_t)k * h + y) * w + x];
                }
    }
};

int argmax_n(int N, halide_buffer_t* gx, halide_buffer_t* gy, halide_buffer_t* lmx, halide_buffer_t* lmy)
{
    Dense<float> x(gx), y(gy);
    DenseOut<uint16_t> ox(lmx), oy(lmy);
    // the pipeline's work domain is the output extent; GradArgMax always asks for (w/N, h/N)
    if (ox.w > x.w / N || ox.h > x.h / N || ox.c != 2 || oy.c != 2) return -1;
    std::vector<uint16_t> fx((size_t)(x.w / N) * (x.h / N) * 2), fy(fx.size());
    vo_grad_argmax(x.p, y.p, x.w, x.h, N, fx.data(), fy.data());
    const int tw = x.w / N, th = x.h / N;
    for (int c = 0; c < 2; c++)
        for (int j = 0; j < ox.h; j++)
            for (int i = 0; i < ox.w; i++) {
                ox.p[((size_t)c * ox.h + j) * ox.w + i] = fx[((size_t)c * th + j) * tw + i];
                oy.p[((size_t)c * oy.h + j) * oy.w + i] = fy[((size_t)c * th + j) * tw + i];
            }
    return 0;
}

}  // namespace

extern "C" {

int pyr_down(halide_buffer_t* input, halide_buffer_t* output)
{
    Dense<uint8_t> in(input);
    DenseOut<uint8_t> out(output);
    vo_pyr_down(in.p, in.w, in.h, out.p, out.w, out.h);
    return 0;
}

int image_warp(halide_buffer_t* input, float A, float B, float TX, float TY, halide_buffer_t* output)
{
    Dense<uint8_t> in(input);
    DenseOut<float> out(output);
    vo_k_image_warp(in.p, in.w, in.h, A, B, TX, TY, out.p, out.w, out.h);
    return 0;
}

int grad_xy(halide_buffer_t* input, halide_buffer_t* grad_x, halide_buffer_t* grad_y)
{
    Dense<uint8_t> in(input);
    DenseOut<float> gx(grad_x), gy(grad_y);
    if (gx.w != gy.w || gx.h != gy.h) return -1;
    vo_grad_xy(in.p, in.w, in.h, gx.p, gy.p, gx.w, gx.h);
    return 0;
}

#define VS_DEF_ARGMAX(N) \
    int grad_argmax_##N(halide_buffer_t* a, halide_buffer_t* b, halide_buffer_t* c, halide_buffer_t* d) { return argmax_n(N, a, b, c, d); }
VS_DEF_ARGMAX(2) VS_DEF_ARGMAX(4) VS_DEF_ARGMAX(6) VS_DEF_ARGMAX(8) VS_DEF_ARGMAX(10)
VS_DEF_ARGMAX(12) VS_DEF_ARGMAX(14) VS_DEF_ARGMAX(16) VS_DEF_ARGMAX(18) VS_DEF_ARGMAX(20)

int sparse_jac(halide_buffer_t* grad_x, halide_buffer_t* grad_y, halide_buffer_t* local_max_x,
               halide_buffer_t* local_max_y, halide_buffer_t* output_x, halide_buffer_t* output_y)
{
    Dense<float> gx(grad_x), gy(grad_y);
    Dense<uint16_t> lx(local_max_x), ly(local_max_y);
    DenseOut<float> ox(output_x), oy(output_y);
    if (ox.w != lx.w || ox.h != lx.h || ox.c != 4 || oy.c != 4) return -1;
    vo_sparse_jac(gx.p, gy.p, gx.w, gx.h, lx.p, ly.p, lx.w, lx.h, ox.p, oy.p);
    return 0;
}

int sparse_ica(halide_buffer_t* input_template, halide_buffer_t* input_keyframe, halide_buffer_t* selected_pixels_x,
               halide_buffer_t* selected_pixels_y, halide_buffer_t* selected_jacobians_x,
               halide_buffer_t* selected_jacobians_y, float A, float B, float TX, float TY, halide_buffer_t* output)
{
    Dense<uint8_t> t(input_template), k(input_keyframe);
    Dense<uint16_t> sx(selected_pixels_x), sy(selected_pixels_y);
    Dense<float> jx(selected_jacobians_x), jy(selected_jacobians_y);
    if (output->dimensions != 1 || output->dim[0].extent != 4) return -1;
    double out[4];
    vo_k_sparse_ica(t.p, k.p, k.w, k.h, sx.p, sx.w, sy.p, sy.w, jx.p, jy.p, A, B, TX, TY, out);
    for (int c = 0; c < 4; c++) ((double*)output->host)[(ptrdiff_t)c * output->dim[0].stride] = out[c];
    return 0;
}

int sparse_warpdiff(halide_buffer_t* input_template, halide_buffer_t* input_keyframe, halide_buffer_t* local_max,
                    float A, float B, float TX, float TY, halide_buffer_t* output)
{
    Dense<uint8_t> t(input_template), k(input_keyframe);
    Dense<uint16_t> lm(local_max);
    DenseOut<uint16_t> out(output);
    if (out.w != lm.w || out.h != lm.h) return -1;
    vo_k_sparse_warpdiff(t.p, k.p, k.w, k.h, lm.p, lm.w, lm.h, A, B, TX, TY, out.p);
    return 0;
}

}  // extern "C"
