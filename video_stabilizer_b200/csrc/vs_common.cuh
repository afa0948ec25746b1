// vs_common.cuh — context, error plumbing and device helpers shared by the sm_100a kernels.
// Compiled with -fmad=false: every f32/f64 operation is individually rounded, in the order
// the reference writes it, so results match the CPU oracle bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "vstab.h"

// kernel ids of the per-kernel timing facility (vs_ctx_profile_*); order = VS_KERNEL_* in vstab.h
enum VsKernelId {
    VSK_BGR2GRAY = 0, VSK_PYR_DOWN, VSK_GRAD_XY, VSK_IMAGE_WARP, VSK_BGR_WARP, VSK_GRAD_ARGMAX,
    VSK_SPARSE_JAC, VSK_WARPDIFF, VSK_ICA, VSK_KEYFRAME, VSK_SOLVE, VSK_INGEST, VSK_COUNT
};

// CUDA-event pair around every launch, on the launching stream, resolved lazily
struct VsProfiler {
    struct Span { cudaEvent_t a, b; int id; };
    std::vector<Span> pending;
    std::vector<cudaEvent_t> pool;
    double ms[VSK_COUNT] = {0};
    int64_t n[VSK_COUNT] = {0};
    int open_id = -1;
    cudaEvent_t open_a = nullptr;
};

struct vs_ctx {
    VsProfiler* prof = nullptr;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    int sm_count = 148;
    // grow-only device scratch used by the VS_MEM_HOST single-operator path
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t scratch_used = 0;
};

int vs_set_error(vs_ctx* ctx, int code, const char* fmt, ...);
void* vs_scratch_alloc(vs_ctx* ctx, size_t bytes);   // nullptr on failure (error set)
void vs_scratch_reset(vs_ctx* ctx);

#define VS_CUDA(ctx, expr)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return vs_set_error((ctx), VS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

void vs_prof_begin(vs_ctx* ctx, int id);
void vs_prof_end(vs_ctx* ctx);
// put directly before a kernel launch that VS_LAUNCH_CHECK follows
#define VS_LAUNCH_BEGIN(ctx, id) do { if ((ctx)->prof) vs_prof_begin((ctx), (id)); } while (0)

#define VS_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        if ((ctx)->prof) vs_prof_end(ctx);                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return vs_set_error((ctx), VS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",       \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define VS_REQUIRE(ctx, cond, msg)                                                            \
    do {                                                                                      \
        if (!(cond)) return vs_set_error((ctx), VS_ERR_INVALID, "%s (%s)", (msg), #cond);     \
    } while (0)

static inline size_t vs_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int vs_cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ int vs_clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// generators.cpp:31-47 — Horner in x^2, zero outside |x| < 2.
__device__ __forceinline__ float vs_lanczos2(float x)
{
    float x2 = __fmul_rn(x, x);
    float v = 0.000858519f;
    v = __fadd_rn(-0.0158853f, __fmul_rn(v, x2));
    v = __fadd_rn(0.128693f, __fmul_rn(v, x2));
    v = __fadd_rn(-0.583468f, __fmul_rn(v, x2));
    v = __fadd_rn(1.52229f, __fmul_rn(v, x2));
    v = __fadd_rn(-2.05238f, __fmul_rn(v, x2));
    v = __fadd_rn(0.999861f, __fmul_rn(v, x2));
    return fabsf(x) >= 2.0f ? 0.0f : v;
}

// generators.cpp:465-497 — Lanczos-2 sample of a repeat-edge u8 image at the warped
// position of (ox,oy).  Tap column/row 0 (offset -2) always has weight exactly 0
// (its argument is <= -2), so adding it leaves the running sums unchanged; it is skipped.
__device__ __forceinline__ float vs_lanczos_sample(const uint8_t* __restrict__ img, int w, int h,
                                                   int pitch, float ox, float oy,
                                                   float A, float B, float TX, float TY)
{
    float onepA = __fadd_rn(1.0f, A);
    float Wx = __fadd_rn(__fsub_rn(__fmul_rn(onepA, ox), __fmul_rn(B, oy)), TX);
    float Wy = __fadd_rn(__fadd_rn(__fmul_rn(B, ox), __fmul_rn(onepA, oy)), TY);
    float fWx = floorf(Wx), fWy = floorf(Wy);
    float rx = __fsub_rn(Wx, fWx), ry = __fsub_rn(Wy, fWy);
    float wx[4], wy[4];
#pragma unroll
    for (int u = 1; u < 5; u++) {
        wx[u - 1] = vs_lanczos2(__fsub_rn((float)(u - 2), rx));
        wy[u - 1] = vs_lanczos2(__fsub_rn((float)(u - 2), ry));
    }
    int ix = (int)fWx, iy = (int)fWy;
    float num = 0.0f, den = 0.0f;
#pragma unroll
    for (int ty = 1; ty < 5; ty++) {
        int sy = vs_clampi(iy + ty - 2, 0, h - 1);
        const uint8_t* row = img + (size_t)sy * pitch;
#pragma unroll
        for (int tx = 1; tx < 5; tx++) {
            int sx = vs_clampi(ix + tx - 2, 0, w - 1);
            float w2 = __fmul_rn(wx[tx - 1], wy[ty - 1]);
            num = __fadd_rn(num, __fmul_rn(w2, (float)__ldg(row + sx)));
            den = __fadd_rn(den, w2);
        }
    }
    return __fdiv_rn(num, den);
}

// The same sample, split so that the loads of the next keypoint can be in flight while the
// current one is evaluated (the solver's loops are latency-bound gathers).  vs_lz_fetch gathers
// the 4x4 window as four 32-bit rows (bytes = columns ix-1 .. ix+2 of rows iy-1 .. iy+2, each
// clamped to the image): two aligned 32-bit loads + a funnel shift per row when the window and
// the second word lie inside the row, byte loads with clamping otherwise.  vs_lz_eval performs
// exactly the operations of vs_lanczos_sample on those values, in the same order.
// Requires a 4-byte aligned image base and pitch.
struct VsLzTaps {
    float rx, ry;
    uint32_t row[4];
};

// position part of a sample: integer origin (ix, iy) = floor(W(p)) and the fractions
__device__ __forceinline__ void vs_lz_pos(float ox, float oy, float A, float B, float TX, float TY, int& ix, int& iy, float& rx, float& ry)
{
    const float onepA = __fadd_rn(1.0f, A);
    const float Wx = __fadd_rn(__fsub_rn(__fmul_rn(onepA, ox), __fmul_rn(B, oy)), TX);
    const float Wy = __fadd_rn(__fadd_rn(__fmul_rn(B, ox), __fmul_rn(onepA, oy)), TY);
    const float fWx = floorf(Wx), fWy = floorf(Wy);
    rx = __fsub_rn(Wx, fWx);
    ry = __fsub_rn(Wy, fWy);
    ix = (int)fWx; iy = (int)fWy;
}

// load part: the window of origin (ix, iy); its bytes are a function of (ix, iy) alone
__device__ __forceinline__ void vs_lz_load(const uint8_t* __restrict__ img, int w, int h, int pitch, int ix, int iy, uint32_t* row)
{
    if (ix >= 1 && ix + 6 < w && iy >= 1 && iy + 2 < h) {
        const int off = (iy - 1) * pitch + ix - 1;
        const uint8_t* p = img + (off & ~3);
        const int sh = (off & 3) * 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(p + j * pitch));
            const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t*>(p + j * pitch + 4));
            row[j] = __funnelshift_r(w0, w1, sh);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint8_t* r = img + (size_t)vs_clampi(iy - 1 + j, 0, h - 1) * pitch;
            row[j] = (uint32_t)__ldg(r + vs_clampi(ix - 1, 0, w - 1)) | ((uint32_t)__ldg(r + vs_clampi(ix, 0, w - 1)) << 8) |
                     ((uint32_t)__ldg(r + vs_clampi(ix + 1, 0, w - 1)) << 16) | ((uint32_t)__ldg(r + vs_clampi(ix + 2, 0, w - 1)) << 24);
        }
    }
}

__device__ __forceinline__ void vs_lz_fetch(const uint8_t* __restrict__ img, int w, int h, int pitch, float ox, float oy,
                                            float A, float B, float TX, float TY, VsLzTaps& t)
{
    int ix, iy;
    vs_lz_pos(ox, oy, A, B, TX, TY, ix, iy, t.rx, t.ry);
    vs_lz_load(img, w, h, pitch, ix, iy, t.row);
}

__device__ __forceinline__ float vs_lz_eval(const VsLzTaps& t)
{
    float wx[4], wy[4];
#pragma unroll
    for (int u = 1; u < 5; u++) {
        wx[u - 1] = vs_lanczos2(__fsub_rn((float)(u - 2), t.rx));
        wy[u - 1] = vs_lanczos2(__fsub_rn((float)(u - 2), t.ry));
    }
    float num = 0.0f, den = 0.0f;
#pragma unroll
    for (int ty = 0; ty < 4; ty++) {
#pragma unroll
        for (int tx = 0; tx < 4; tx++) {
            const float w2 = __fmul_rn(wx[tx], wy[ty]);
            num = __fadd_rn(num, __fmul_rn(w2, (float)((t.row[ty] >> (8 * tx)) & 0xffu)));
            den = __fadd_rn(den, w2);
        }
    }
    return __fdiv_rn(num, den);
}

__device__ __forceinline__ double vs_warp_reduce_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

// ------------------------------------------------ transform algebra (host + device)
#ifdef __CUDACC__
#define VS_HD __host__ __device__ __forceinline__
#else
#define VS_HD inline
#endif

// imgproc.cpp:361-387 — out = T2 o T1 (T1 applied first)
VS_HD void vs_tf_compose(const double* T1, const double* T2, double* out)
{
    double p1 = 1.0 + T1[0], q1 = T1[1];
    double p2 = 1.0 + T2[0], q2 = T2[1];
    double A3 = (p2 * p1 - q2 * q1) - 1.0;
    double B3 = (p2 * q1 + q2 * p1);
    double TX3 = p2 * T1[2] - q2 * T1[3] + T2[2];
    double TY3 = q2 * T1[2] + p2 * T1[3] + T2[3];
    out[0] = A3; out[1] = B3; out[2] = TX3; out[3] = TY3;
}

// imgproc.cpp:333-359
VS_HD void vs_tf_inverse(const double* T, double* out)
{
    double p = 1.0 + T[0], q = T[1];
    double denom = p * p + q * q;
    double a = (p / denom) - 1.0;
    double b = -q / denom;
    double ix = (-p * T[2] - q * T[3]) / denom;
    double iy = (q * T[2] - p * T[3]) / denom;
    out[0] = a; out[1] = b; out[2] = ix; out[3] = iy;
}

// imgproc.cpp:401-411
VS_HD void vs_tf_warp_center(const double* T, double px, double py, double cx, double cy, double* o)
{
    double x = px - cx, y = py - cy;
    o[0] = (1 + T[0]) * x - T[1] * y + cx + T[2];
    o[1] = T[1] * x + (1 + T[0]) * y + cy + T[3];
}

// imgproc.cpp:69-75 / :98-103 — centre-based -> f32 upper-left kernel parameters
VS_HD void vs_ul_params_half(const double* T, int w, int h, float* out)
{
    double hw = (double)((float)w * 0.5f), hh = (double)((float)h * 0.5f);
    out[0] = (float)T[0];
    out[1] = (float)T[1];
    out[2] = (float)(T[2] - T[0] * hw + T[1] * hh);
    out[3] = (float)(T[3] - T[1] * hw - T[0] * hh);
}
