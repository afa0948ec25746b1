/*
 * vs_oracle.cpp — CPU restatement of catid/video_stabilizer's alignment-and-warp path.
 *
 * TEST INFRASTRUCTURE ONLY (see vs_oracle.h).  Not linked into the product.
 *
 * Build: g++ -std=c++17 -O2 -ffp-contract=off -fno-fast-math  (oracle/Makefile).
 * Canonical float rounding: every f32/f64 operation is
 * individually rounded, evaluated in the order written in the reference, with no
 * FMA contraction.  The CUDA kernels use the same order with -fmad=false.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Third-party arithmetic not under /root/reference:
 *   - OpenCV (version unpinned by the reference; arbiter here: cv2 4.13):
 *     cvtColor(BGR2GRAY), warpAffine(INTER_LINEAR,BORDER_CONSTANT), SVD, Mat::inv(DECOMP_SVD)
 *     — restated below from the published algorithms, pinned by tests/golden fixtures.
 *   - Halide 19 (BoundaryConditions::repeat_edge, argmax tie rule, lerp) — restated.
 *   - libstdc++ 13.3 std::nth_element — called directly.
 */
#include "vs_oracle.h"

#include <algorithm>
#include <cfloat>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>
#ifdef VS_ORACLE_FAST
#include <immintrin.h>
#endif

// VS_ORACLE_FAST (the *_fast.so builds, which only ever serve as the TIMED CPU baseline): pyr_down, grad_xy and the cv-exact
// BGR warp are computed by vectorisable / AVX2 forms that produce the same bytes as the restatements below
// (tests/test_oracle_golden.py::test_fast_kernels_equal_the_restatement) — the reference's Halide schedules and OpenCV's
// warpAffine are vectorised too, so a scalar restatement would flatter the GPU arm.
namespace {

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// generators.cpp:31-47 — even degree-12 polynomial in x^2 (Horner), zero for |x| >= 2.
inline float lanczos2(float x)
{
    float x2 = x * x;
    float val = 0.000858519f;
    val = -0.0158853f + val * x2;
    val = 0.128693f + val * x2;
    val = -0.583468f + val * x2;
    val = 1.52229f + val * x2;
    val = -2.05238f + val * x2;
    val = 0.999861f + val * x2;
    return std::fabs(x) >= 2.0f ? 0.0f : val;
}

// generators.cpp:465-497 (identical at :505-537 and :665-697) — 5x5 Lanczos-2 sample of
// `img` (repeat-edge) at the similarity-warped position of (px,py).
// A,B,TX,TY are the f32 upper-left-origin kernel parameters.
inline float lanczos_sample(const uint8_t* img, int w, int h, float ox, float oy,
                            float A, float B, float TX, float TY)
{
    float Wx = (1.0f + A) * ox - B * oy + TX;
    float Wy = B * ox + (1.0f + A) * oy + TY;
    float fWx = std::floor(Wx);
    float fWy = std::floor(Wy);
    float rx = Wx - fWx;
    float ry = Wy - fWy;
    float wx[5], wy[5];
    for (int u = 0; u < 5; u++) {
        wx[u] = lanczos2((float)(u - 2) - rx);
        wy[u] = lanczos2((float)(u - 2) - ry);
    }
    int ix = (int)fWx, iy = (int)fWy;
    float num = 0.0f, den = 0.0f;
    for (int ty = 0; ty < 5; ty++) {
        int sy = clampi(iy + ty - 2, 0, h - 1);
        for (int tx = 0; tx < 5; tx++) {
            int sx = clampi(ix + tx - 2, 0, w - 1);
            float w2 = wx[tx] * wy[ty];
            num = num + w2 * (float)img[(size_t)sy * w + sx];
            den = den + w2;
        }
    }
    return num / den;
}

// imgproc.cpp:69-75 and :98-103 — centre-based transform to the f32 UL-origin kernel params.
// Note (w * 0.5f) is an f32 product promoted to f64.
inline void ul_params_half(const double T[4], int w, int h, float out[4])
{
    double A = T[0], B = T[1], TX = T[2], TY = T[3];
    out[0] = (float)A;
    out[1] = (float)B;
    out[2] = (float)(TX - A * (double)((float)w * 0.5f) + B * (double)((float)h * 0.5f));
    out[3] = (float)(TY - B * (double)((float)w * 0.5f) - A * (double)((float)h * 0.5f));
}

} // namespace

extern "C" {

//------------------------------------------------------------------------------
// cv::cvtColor(BGR2GRAY) as called at alignment.cpp:212.  OpenCV's 8-bit path:
// 15-bit fixed-point coefficients B=3735, G=19235, R=9798 with round-to-nearest.
void vo_bgr2gray(const uint8_t* bgr, int w, int h, uint8_t* gray)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        unsigned b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
    }
}

// generators.cpp:56-92 — separable [1 4 6 4 1]/16 in Y then X on repeat-edge input,
// sampled at (2x,2y), truncating u8 cast.  Evaluated in f32 exactly as written; every
// intermediate is exactly representable so this equals (sum k_j k_i p) >> 8.
void vo_pyr_down(const uint8_t* in, int iw, int ih, uint8_t* out, int ow, int oh)
{
#ifdef VS_ORACLE_FAST
    // every intermediate of the f32 form is an exactly representable dyadic number, so it equals
    // (sum_j sum_i k_j k_i in) >> 8 in integers: vertical [1 4 6 4 1] sums of whole rows, then the horizontal sums
    std::vector<uint16_t> col((size_t)iw);
    for (int y = 0; y < oh; y++) {
        const uint8_t* r0 = in + (size_t)clampi(2 * y - 2, 0, ih - 1) * iw;
        const uint8_t* r1 = in + (size_t)clampi(2 * y - 1, 0, ih - 1) * iw;
        const uint8_t* r2 = in + (size_t)clampi(2 * y, 0, ih - 1) * iw;
        const uint8_t* r3 = in + (size_t)clampi(2 * y + 1, 0, ih - 1) * iw;
        const uint8_t* r4 = in + (size_t)clampi(2 * y + 2, 0, ih - 1) * iw;
        for (int x = 0; x < iw; x++)
            col[x] = (uint16_t)(r0[x] + 4 * r1[x] + 6 * r2[x] + 4 * r3[x] + r4[x]);
        uint8_t* o = out + (size_t)y * ow;
        const int x_lo = 1, x_hi = std::min(ow, (iw - 3) / 2 + 1);      // 2x - 2 >= 0 and 2x + 2 <= iw - 1
        for (int x = 0; x < ow; x++) {
            if (x >= x_lo && x < x_hi) continue;
            const uint32_t v = col[clampi(2 * x - 2, 0, iw - 1)] + 4u * col[clampi(2 * x - 1, 0, iw - 1)] + 6u * col[clampi(2 * x, 0, iw - 1)] +
                               4u * col[clampi(2 * x + 1, 0, iw - 1)] + col[clampi(2 * x + 2, 0, iw - 1)];
            o[x] = (uint8_t)(v >> 8);
        }
        for (int x = x_lo; x < x_hi; x++) {
            const uint16_t* cc = col.data() + 2 * x;
            const uint32_t v = cc[-2] + 4u * cc[-1] + 6u * cc[0] + 4u * cc[1] + cc[2];
            o[x] = (uint8_t)(v >> 8);
        }
    }
    return;
#endif
    const float c[5] = {1.0f / 16, 4.0f / 16, 6.0f / 16, 4.0f / 16, 1.0f / 16};
    for (int y = 0; y < oh; y++) {
        for (int x = 0; x < ow; x++) {
            float by[5];
            for (int i = 0; i < 5; i++) {
                int sx = clampi(2 * x + i - 2, 0, iw - 1);
                float acc = c[0] * (float)in[(size_t)clampi(2 * y - 2, 0, ih - 1) * iw + sx];
                acc = acc + c[1] * (float)in[(size_t)clampi(2 * y - 1, 0, ih - 1) * iw + sx];
                acc = acc + c[2] * (float)in[(size_t)clampi(2 * y, 0, ih - 1) * iw + sx];
                acc = acc + c[3] * (float)in[(size_t)clampi(2 * y + 1, 0, ih - 1) * iw + sx];
                acc = acc + c[4] * (float)in[(size_t)clampi(2 * y + 2, 0, ih - 1) * iw + sx];
                by[i] = acc;
            }
            float v = c[0] * by[0];
            v = v + c[1] * by[1];
            v = v + c[2] * by[2];
            v = v + c[3] * by[3];
            v = v + c[4] * by[4];
            out[(size_t)y * ow + x] = (uint8_t)v;
        }
    }
}

// generators.cpp:202-224 — central differences * 0.5 on repeat-edge input.
void vo_grad_xy(const uint8_t* in, int iw, int ih, float* gx, float* gy, int ow, int oh)
{
#ifdef VS_ORACLE_FAST
    if (ow <= iw && oh <= ih) {      // the same operations, rows with unclamped interiors (vectorisable)
        for (int y = 0; y < oh; y++) {
            const uint8_t* r = in + (size_t)y * iw;
            const uint8_t* rp = in + (size_t)clampi(y + 1, 0, ih - 1) * iw;
            const uint8_t* rm = in + (size_t)clampi(y - 1, 0, ih - 1) * iw;
            float* ox = gx + (size_t)y * ow;
            float* oy = gy + (size_t)y * ow;
            for (int x = 0; x < ow; x++) oy[x] = 0.5f * ((float)rp[x] - (float)rm[x]);
            const int x_hi = std::min(ow, iw - 1);
            if (ow > 0) ox[0] = 0.5f * ((float)r[clampi(1, 0, iw - 1)] - (float)r[0]);
            for (int x = 1; x < x_hi; x++) ox[x] = 0.5f * ((float)r[x + 1] - (float)r[x - 1]);
            for (int x = std::max(1, x_hi); x < ow; x++) ox[x] = 0.5f * ((float)r[clampi(x + 1, 0, iw - 1)] - (float)r[clampi(x - 1, 0, iw - 1)]);
        }
        return;
    }
#endif
    for (int y = 0; y < oh; y++) {
        for (int x = 0; x < ow; x++) {
            int yc = clampi(y, 0, ih - 1), xc = clampi(x, 0, iw - 1);
            float xp = (float)in[(size_t)yc * iw + clampi(x + 1, 0, iw - 1)];
            float xm = (float)in[(size_t)yc * iw + clampi(x - 1, 0, iw - 1)];
            float yp = (float)in[(size_t)clampi(y + 1, 0, ih - 1) * iw + xc];
            float ym = (float)in[(size_t)clampi(y - 1, 0, ih - 1) * iw + xc];
            gx[(size_t)y * ow + x] = 0.5f * (xp - xm);
            gy[(size_t)y * ow + x] = 0.5f * (yp - ym);
        }
    }
}

// imgproc.cpp:151-162 — largest even tile size in [2,20] keeping >= 1000 tiles.
int vo_tile_size(int w, int h)
{
    int tile = 2;
    for (int i = 4; i <= 20; i += 2) {
        int tx = w / i, ty = h / i;
        if (tx * ty < 1000) break;
        tile = i;
    }
    return tile;
}

// generators.cpp:260-294 — per-tile Halide::argmax(abs(g)) over RDom(0,N,0,N):
// r.x is the inner loop, r.y the outer; strict '>' keeps the first maximum; the initial
// best is (0,0) with the lowest float so an all-zero tile returns the tile origin.
void vo_grad_argmax(const float* gx, const float* gy, int w, int h, int tile,
                    uint16_t* lmx, uint16_t* lmy)
{
    int tw = w / tile, th = h / tile;
    for (int axis = 0; axis < 2; axis++) {
        const float* g = axis == 0 ? gx : gy;
        uint16_t* lm = axis == 0 ? lmx : lmy;
        for (int ty = 0; ty < th; ty++) {
            for (int tx = 0; tx < tw; tx++) {
                float best = -INFINITY;
                int bx = 0, by = 0;
                for (int ry = 0; ry < tile; ry++) {
                    for (int rx = 0; rx < tile; rx++) {
                        float v = std::fabs(g[(size_t)(ty * tile + ry) * w + tx * tile + rx]);
                        if (v > best) { best = v; bx = rx; by = ry; }
                    }
                }
                lm[(size_t)(0 * th + ty) * tw + tx] = (uint16_t)(bx + tx * tile);
                lm[(size_t)(1 * th + ty) * tw + tx] = (uint16_t)(by + ty * tile);
            }
        }
    }
}

// generators.cpp:332-386 — 4-vector Jacobian per keypoint, X-only / Y-only split.
void vo_sparse_jac(const float* gx, const float* gy, int w, int h,
                   const uint16_t* lmx, const uint16_t* lmy, int tw, int th,
                   float* jx, float* jy)
{
    float cx = (float)w * 0.5f, cy = (float)h * 0.5f;
    float scale = 1.f / (float)w;
    size_t plane = (size_t)tw * th;
    for (int ty = 0; ty < th; ty++) {
        for (int tx = 0; tx < tw; tx++) {
            size_t t = (size_t)ty * tw + tx;
            int ix0 = std::min<int>(lmx[t], w - 1), iy0 = std::min<int>(lmx[plane + t], h - 1);
            int ix1 = std::min<int>(lmy[t], w - 1), iy1 = std::min<int>(lmy[plane + t], h - 1);
            float u0 = (float)ix0 - cx, v0 = (float)iy0 - cy;
            float u1 = (float)ix1 - cx, v1 = (float)iy1 - cy;
            float g0 = gx[(size_t)iy0 * w + ix0];
            float g1 = gy[(size_t)iy1 * w + ix1];
            jx[0 * plane + t] = 2.f * g0 * u0 * scale;
            jx[1 * plane + t] = 2.f * g0 * (-v0) * scale;
            jx[2 * plane + t] = 2.f * g0;
            jx[3 * plane + t] = 0.f;
            jy[0 * plane + t] = 2.f * g1 * v1 * scale;
            jy[1 * plane + t] = 2.f * g1 * u1 * scale;
            jy[2 * plane + t] = 0.f;
            jy[3 * plane + t] = 2.f * g1;
        }
    }
}

// imgproc.cpp:80-106 + generators.cpp:646-700 — |Lanczos2(keyframe, W(p)) - template(p)|,
// clamped to [0,65535] and truncated to u16, for every tile keypoint.
void vo_k_sparse_warpdiff(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                          const uint16_t* lm, int tw, int th, float A, float B, float TX, float TY,
                          uint16_t* out)
{
    size_t plane = (size_t)tw * th;
    for (size_t t = 0; t < plane; t++) {
        int px = std::min<int>(lm[t], w - 1);
        int py = std::min<int>(lm[plane + t], h - 1);
        float s = lanczos_sample(key, w, h, (float)px, (float)py, A, B, TX, TY);
        float d = std::fabs(s - (float)tmpl[(size_t)py * w + px]);
        d = std::max(std::min(d, 65535.0f), 0.0f);
        out[t] = (uint16_t)d;
    }
}

void vo_sparse_warpdiff(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                        const uint16_t* lm, int tw, int th, const double T[4],
                        uint16_t* out)
{
    float P[4];
    ul_params_half(T, w, h, P);
    vo_k_sparse_warpdiff(tmpl, key, w, h, lm, tw, th, P[0], P[1], P[2], P[3], out);
}

// imgproc.cpp:46-78 + generators.cpp:429-596 — b = 0.5 * (sum_X Jx r + sum_Y Jy r);
// J*r is an f32 product, accumulated serially in f64 per channel.
void vo_k_sparse_ica(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                     const uint16_t* selx, int kx, const uint16_t* sely, int ky,
                     const float* jx, const float* jy, float A, float B, float TX, float TY, double out[4])
{
    double rx[4] = {0, 0, 0, 0}, ry[4] = {0, 0, 0, 0};
    for (int axis = 0; axis < 2; axis++) {
        const uint16_t* sel = axis == 0 ? selx : sely;
        const float* jac = axis == 0 ? jx : jy;
        int k = axis == 0 ? kx : ky;
        double* acc = axis == 0 ? rx : ry;
        for (int i = 0; i < k; i++) {
            int px = sel[i], py = sel[(size_t)k + i];
            float warped = lanczos_sample(key, w, h, (float)px, (float)py, A, B, TX, TY);
            int tx = std::min(px, w - 1), ty = std::min(py, h - 1);
            float residual = (float)tmpl[(size_t)ty * w + tx] - warped;
            for (int c = 0; c < 4; c++) {
                float prod = jac[(size_t)c * k + i] * residual;
                acc[c] += (double)prod;
            }
        }
    }
    for (int c = 0; c < 4; c++) out[c] = (rx[c] + ry[c]) * (double)0.5f;
}

void vo_sparse_ica(const uint8_t* tmpl, const uint8_t* key, int w, int h,
                   const uint16_t* selx, int kx, const uint16_t* sely, int ky,
                   const float* jx, const float* jy, const double T[4], double out[4])
{
    float P[4];
    ul_params_half(T, w, h, P);
    vo_k_sparse_ica(tmpl, key, w, h, selx, kx, sely, ky, jx, jy, P[0], P[1], P[2], P[3], out);
}

// imgproc.cpp:116-133 + generators.cpp:126-164 — pull-mapped bilinear, repeat-edge.
// Halide's float lerp(a,b,t) is a*(1-t) + b*t.
void vo_k_image_warp(const uint8_t* in, int iw, int ih, float A, float B, float TX, float TY,
                     float* out, int ow, int oh)
{
    for (int y = 0; y < oh; y++) {
        for (int x = 0; x < ow; x++) {
            float Wx = (1.0f + A) * (float)x - B * (float)y + TX;
            float Wy = B * (float)x + (1.0f + A) * (float)y + TY;
            int fx = (int)std::floor(Wx), fy = (int)std::floor(Wy);
            float wx = Wx - (float)fx, wy = Wy - (float)fy;
            int x0 = clampi(fx, 0, iw - 1), x1 = clampi(fx + 1, 0, iw - 1);
            int y0 = clampi(fy, 0, ih - 1), y1 = clampi(fy + 1, 0, ih - 1);
            float p00 = (float)in[(size_t)y0 * iw + x0], p10 = (float)in[(size_t)y0 * iw + x1];
            float p01 = (float)in[(size_t)y1 * iw + x0], p11 = (float)in[(size_t)y1 * iw + x1];
            float top = p00 * (1.0f - wx) + p10 * wx;
            float bot = p01 * (1.0f - wx) + p11 * wx;
            out[(size_t)y * ow + x] = top * (1.0f - wy) + bot * wy;
        }
    }
}

void vo_image_warp(const uint8_t* in, int iw, int ih, const double T[4],
                   float* out, int ow, int oh)
{
    double cx = (iw - 1) * 0.5, cy = (ih - 1) * 0.5;
    float A = (float)T[0], B = (float)T[1];
    float TX = (float)(T[2] - T[0] * cx + T[1] * cy);
    float TY = (float)(T[3] - T[1] * cx - T[0] * cy);
    vo_k_image_warp(in, iw, ih, A, B, TX, TY, out, ow, oh);
}

//------------------------------------------------------------------------------
// imgproc.cpp:446-484 — BGR warp.  Mode 0 restates cv::warpAffine(INTER_LINEAR,
// BORDER_CONSTANT 0) without WARP_INVERSE_MAP: OpenCV inverts M in f64, walks the
// destination with 10-bit fixed-point coordinates (AB_BITS=10), rounds to 1/32 px
// (INTER_BITS=5) and blends with 15-bit integer weights.  Modes 1 and 2 have no counterpart
// in the reference (its bgr_image_warp generator is gone): they are defined here, and in
// the kernels alike, for the interpolation sweep (BASELINE.json configs[4]).  Mode 1 is the
// same exact integer bilinear on a finer grid (16 position bits, 1/256-pixel weights); mode
// 2 is a 4 x 4 Lanczos-2 on that grid with tabulated Q11 weights (64 fractions), integer too.
static inline long rint_he(double v) { return std::lrint(v); } // round-half-even (default FE mode)

// Q11 Lanczos-2 weights of the 64 fractions q/64 for the taps at offsets -1, 0, 1, 2: lanczos2(t - f) with the reference's
// polynomial (generators.cpp:31-47) in double, normalised; the rounding residue goes to the largest weight so that every
// row sums to exactly 2048.
void vo_lanczos_table(int16_t* tab /* [64][4] */)
{
    for (int q = 0; q < 64; q++) {
        const double f = q / 64.0;
        double wgt[4], sum = 0.0;
        for (int t = 0; t < 4; t++) {
            const double x = (double)(t - 1) - f, x2 = x * x;
            double v = 0.000858519;
            v = -0.0158853 + v * x2;
            v = 0.128693 + v * x2;
            v = -0.583468 + v * x2;
            v = 1.52229 + v * x2;
            v = -2.05238 + v * x2;
            v = 0.999861 + v * x2;
            wgt[t] = std::fabs(x) >= 2.0 ? 0.0 : v;
            sum += wgt[t];
        }
        int iw[4], isum = 0, big = 0;
        for (int t = 0; t < 4; t++) {
            iw[t] = (int)std::lrint(wgt[t] / sum * 2048.0);
            isum += iw[t];
            if (iw[t] > iw[big]) big = t;
        }
        iw[big] += 2048 - isum;
        for (int t = 0; t < 4; t++) tab[q * 4 + t] = (int16_t)iw[t];
    }
}

#ifdef VS_ORACLE_FAST
// Eight output pixels of one row whose 2 x 2 taps all lie inside the source: the integer arithmetic of the scalar loop in
// AVX2 lanes (four 32-bit gathers of B G R x, per channel four multiply-adds).  Returns false when a pixel is not inside.
__attribute__((target("avx2"))) static inline bool warp8_avx2(const uint8_t* src, int w, int h, const int* ad, const int* bd, int X0, int Y0,
                                                              int P, int W, uint8_t* d)
{
    const __m256i sfx = _mm256_add_epi32(_mm256_set1_epi32(X0), _mm256_loadu_si256((const __m256i*)ad));
    const __m256i sfy = _mm256_add_epi32(_mm256_set1_epi32(Y0), _mm256_loadu_si256((const __m256i*)bd));
    const __m128i cP = _mm_cvtsi32_si128(P), cPW = _mm_cvtsi32_si128(P - W);
    const __m256i sx = _mm256_sra_epi32(sfx, cP), sy = _mm256_sra_epi32(sfy, cP);
    // inside: 0 <= sx, sx + 1 < w, 0 <= sy, sy + 2 < h (the bottom row pair is left to the scalar loop: a 32-bit gather of the
    // frame's last pixel would read one byte past the buffer)
    const __m256i bad = _mm256_or_si256(_mm256_or_si256(_mm256_cmpgt_epi32(_mm256_setzero_si256(), sx), _mm256_cmpgt_epi32(_mm256_add_epi32(sx, _mm256_set1_epi32(2)), _mm256_set1_epi32(w))),
                                        _mm256_or_si256(_mm256_cmpgt_epi32(_mm256_setzero_si256(), sy), _mm256_cmpgt_epi32(_mm256_add_epi32(sy, _mm256_set1_epi32(3)), _mm256_set1_epi32(h))));
    if (!_mm256_testz_si256(bad, bad)) return false;
    const __m256i one = _mm256_set1_epi32(1 << W), msk = _mm256_set1_epi32((1 << W) - 1);
    const __m256i fx = _mm256_and_si256(_mm256_sra_epi32(sfx, cPW), msk), fy = _mm256_and_si256(_mm256_sra_epi32(sfy, cPW), msk);
    const __m256i gx = _mm256_sub_epi32(one, fx), gy = _mm256_sub_epi32(one, fy);
    const __m256i w00 = _mm256_mullo_epi32(gx, gy), w10 = _mm256_mullo_epi32(fx, gy), w01 = _mm256_mullo_epi32(gx, fy), w11 = _mm256_mullo_epi32(fx, fy);
    const __m256i off = _mm256_mullo_epi32(_mm256_add_epi32(_mm256_mullo_epi32(sy, _mm256_set1_epi32(w)), sx), _mm256_set1_epi32(3));
    const int* base = (const int*)src;
    const __m256i t00 = _mm256_i32gather_epi32(base, off, 1);
    const __m256i t10 = _mm256_i32gather_epi32(base, _mm256_add_epi32(off, _mm256_set1_epi32(3)), 1);
    const __m256i t01 = _mm256_i32gather_epi32(base, _mm256_add_epi32(off, _mm256_set1_epi32(3 * w)), 1);
    const __m256i t11 = _mm256_i32gather_epi32(base, _mm256_add_epi32(off, _mm256_set1_epi32(3 * w + 3)), 1);
    const __m256i half = _mm256_set1_epi32(1 << (2 * W - 1)), m8 = _mm256_set1_epi32(255);
    const __m128i c2W = _mm_cvtsi32_si128(2 * W);
    __m256i px = _mm256_setzero_si256();
    for (int c = 0; c < 3; c++) {
        const __m128i sh = _mm_cvtsi32_si128(8 * c);
        __m256i v = _mm256_mullo_epi32(w00, _mm256_and_si256(_mm256_srl_epi32(t00, sh), m8));
        v = _mm256_add_epi32(v, _mm256_mullo_epi32(w10, _mm256_and_si256(_mm256_srl_epi32(t10, sh), m8)));
        v = _mm256_add_epi32(v, _mm256_mullo_epi32(w01, _mm256_and_si256(_mm256_srl_epi32(t01, sh), m8)));
        v = _mm256_add_epi32(v, _mm256_mullo_epi32(w11, _mm256_and_si256(_mm256_srl_epi32(t11, sh), m8)));
        v = _mm256_srl_epi32(_mm256_add_epi32(v, half), c2W);
        px = _mm256_or_si256(px, _mm256_sll_epi32(v, sh));
    }
    // B G R x of four pixels per 128-bit lane -> 12 packed bytes per lane
    const __m256i pk = _mm256_shuffle_epi8(px, _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1,
                                                                0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1));
    alignas(32) uint8_t o[32];
    _mm256_store_si256((__m256i*)o, pk);
    std::memcpy(d, o, 12);
    std::memcpy(d + 12, o + 16, 12);
    return true;
}
#endif

// General form: forward 2x3 matrix M (as cv::warpAffine takes it), source sw x sh, destination window dw x dh whose pixel
// (x, y) is output pixel (x + dx0, y + dy0) of the warp.  vo_warp_bgr below and the synthetic-clip renderer use it.
void vo_warp_bgr_matrix(const uint8_t* src, int w, int h, const double M[6], uint8_t* dst, int ow, int oh, int dx0, int dy0,
                        int mode, int border)
{
    const double m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[3], m11 = M[4], m12 = M[5];
    // cv::warpAffine: invert the 2x3 matrix
    double D = m00 * m11 - m01 * m10;
    D = D != 0 ? 1.0 / D : 0;
    double i00 = m11 * D, i01 = m01 * (-D), i10 = m10 * (-D), i11 = m00 * D;
    double i02 = -i00 * m02 - i01 * m12;
    double i12 = -i10 * m02 - i11 * m12;

    if (mode == 0 || mode == 1) {
        // Both bilinear modes are exact integer arithmetic on a fixed-point grid with P fractional position bits of which
        // the top W are the weight: mode 0 is cv::warpAffine's own (AB_BITS = 10, INTER_BITS = 5), mode 1 keeps 16
        // position bits and 1/256-pixel weights (no counterpart upstream; defined here and in the kernels alike).
        const int P = mode == 0 ? 10 : 16, W = mode == 0 ? 5 : 8;
        const double SCALE = (double)(1 << P);
        const int ROUND = 1 << (P - W - 1), one = 1 << W, half = 1 << (2 * W - 1);
        std::vector<int> adelta(ow), bdelta(ow);
        for (int xo = 0; xo < ow; xo++) {
            adelta[xo] = (int)rint_he(i00 * (xo + dx0) * SCALE);
            bdelta[xo] = (int)rint_he(i10 * (xo + dx0) * SCALE);
        }
        for (int yo = 0; yo < oh; yo++) {
            int y = yo + dy0;
            int X0 = (int)rint_he((i01 * y + i02) * SCALE) + ROUND;
            int Y0 = (int)rint_he((i11 * y + i12) * SCALE) + ROUND;
            for (int xo = 0; xo < ow; xo++) {
#ifdef VS_ORACLE_FAST
                // (mode 1's weights reach 2^16 and its products 2^24: still inside 32 bits)
                static const bool have_avx2 = __builtin_cpu_supports("avx2");
                if (have_avx2 && xo + 8 <= ow &&
                    warp8_avx2(src, w, h, adelta.data() + xo, bdelta.data() + xo, X0, Y0, P, W, dst + ((size_t)yo * ow + xo) * 3)) {
                    xo += 7;
                    continue;
                }
#endif
                const int sfx = X0 + adelta[xo], sfy = Y0 + bdelta[xo];
                const int sx = sfx >> P, sy = sfy >> P;
                const int fx = (sfx >> (P - W)) & (one - 1), fy = (sfy >> (P - W)) & (one - 1);
                // weights sum to 2^(2W); for mode 0 this equals cv's 15-bit weights and (v + 16384) >> 15
                const int w00 = (one - fx) * (one - fy), w10 = fx * (one - fy), w01 = (one - fx) * fy, w11 = fx * fy;
                uint8_t* d = dst + ((size_t)yo * ow + xo) * 3;
                const bool inside = sx >= 0 && sx + 1 < w && sy >= 0 && sy + 1 < h;
                if (inside) {
                    const uint8_t* p0 = src + ((size_t)sy * w + sx) * 3;
                    const uint8_t* p1 = p0 + (size_t)w * 3;
                    for (int c = 0; c < 3; c++) {
                        int v = w00 * p0[c] + w10 * p0[3 + c] + w01 * p1[c] + w11 * p1[3 + c];
                        d[c] = (uint8_t)((v + half) >> (2 * W));
                    }
                    continue;
                }
                for (int c = 0; c < 3; c++) {
                    auto tap = [&](int xx, int yy) -> int {
                        if (border == 1) { xx = clampi(xx, 0, w - 1); yy = clampi(yy, 0, h - 1); }
                        else if (xx < 0 || xx >= w || yy < 0 || yy >= h) return 0;
                        return src[((size_t)yy * w + xx) * 3 + c];
                    };
                    int v = w00 * tap(sx, sy) + w10 * tap(sx + 1, sy) +
                            w01 * tap(sx, sy + 1) + w11 * tap(sx + 1, sy + 1);
                    d[c] = (uint8_t)((v + half) >> (2 * W));
                }
            }
        }
        return;
    }

    // Mode 2: Lanczos-2 (4 x 4 taps) on the same 16.16 grid, weights from a table of 64 fractions in Q11 (the reference's
    // lanczos2 polynomial, generators.cpp:31-47, normalised to sum 2048), exact integer arithmetic: the vertical sums of
    // the four source columns first (Q11), the horizontal sum of those (Q22; |sum| < 1.4e9 fits 32 bits), rounded and clamped.
    {
        const int P = 16, W = 6;
        const double SCALE = (double)(1 << P);
        const int ROUND = 1 << (P - W - 1);
        int16_t tab[64][4];
        vo_lanczos_table(&tab[0][0]);
        for (int yo = 0; yo < oh; yo++) {
            const int y = yo + dy0;
            const int X0 = (int)rint_he((i01 * y + i02) * SCALE) + ROUND;
            const int Y0 = (int)rint_he((i11 * y + i12) * SCALE) + ROUND;
            for (int xo = 0; xo < ow; xo++) {
                const int sfx = X0 + (int)rint_he(i00 * (xo + dx0) * SCALE), sfy = Y0 + (int)rint_he(i10 * (xo + dx0) * SCALE);
                const int ix = sfx >> P, iy = sfy >> P;
                const int16_t* wx = tab[(sfx >> (P - W)) & 63];
                const int16_t* wy = tab[(sfy >> (P - W)) & 63];
                uint8_t* d = dst + ((size_t)yo * ow + xo) * 3;
                for (int c = 0; c < 3; c++) {
                    int hsum = 1 << 21;
                    for (int t = 0; t < 4; t++) {
                        int v = 0;
                        for (int r = 0; r < 4; r++) {
                            int xx = ix - 1 + t, yy = iy - 1 + r, p = 0;
                            if (border == 1) { xx = clampi(xx, 0, w - 1); yy = clampi(yy, 0, h - 1); p = src[((size_t)yy * w + xx) * 3 + c]; }
                            else if (xx >= 0 && xx < w && yy >= 0 && yy < h) p = src[((size_t)yy * w + xx) * 3 + c];
                            v += (int)wy[r] * p;
                        }
                        hsum += (int)wx[t] * v;
                    }
                    d[c] = (uint8_t)clampi(hsum >> 22, 0, 255);
                }
            }
        }
    }
}

// imgproc.cpp:458-466: centre-based similarity -> forward matrix, same-size output cropped by `crop` on every side
void vo_warp_bgr(const uint8_t* src, int w, int h, const double T[4],
                 uint8_t* dst, int mode, int border, int crop)
{
    double A = T[0], B = T[1];
    double cx = (w - 1) * 0.5, cy = (h - 1) * 0.5;
    const double M[6] = {1.0 + A, -B, T[2] - A * cx + B * cy, B, 1.0 + A, T[3] - B * cx - A * cy};
    vo_warp_bgr_matrix(src, w, h, M, dst, w - 2 * crop, h - 2 * crop, crop, crop, mode, border);
}

// The planes of an NV12 frame (no counterpart upstream, whose frames are BGR cv::Mat; the product's VS_CLIP_NV12 clips):
// cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of a u8 image with `ch` interleaved channels (1 = Y, 2 = UV), the same
// fixed-point arithmetic as mode 0 above.  Pinned against cv2.warpAffine on 1- and 2-channel images
// (tests/golden/plane_warp.npz).
void vo_warp_plane_matrix(const uint8_t* src, int w, int h, int ch, const double M[6], uint8_t* dst, int ow, int oh, int dx0, int dy0)
{
    const double m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[3], m11 = M[4], m12 = M[5];
    double D = m00 * m11 - m01 * m10;
    D = D != 0 ? 1.0 / D : 0;
    const double i00 = m11 * D, i01 = m01 * (-D), i10 = m10 * (-D), i11 = m00 * D;
    const double i02 = -i00 * m02 - i01 * m12, i12 = -i10 * m02 - i11 * m12;
    for (int yo = 0; yo < oh; yo++) {
        const int y = yo + dy0;
        const int X0 = (int)rint_he((i01 * y + i02) * 1024.0) + 16, Y0 = (int)rint_he((i11 * y + i12) * 1024.0) + 16;
        for (int xo = 0; xo < ow; xo++) {
            const int sfx = X0 + (int)rint_he(i00 * (xo + dx0) * 1024.0), sfy = Y0 + (int)rint_he(i10 * (xo + dx0) * 1024.0);
            const int sx = sfx >> 10, sy = sfy >> 10, fx = (sfx >> 5) & 31, fy = (sfy >> 5) & 31;
            const int w00 = (32 - fx) * (32 - fy), w10 = fx * (32 - fy), w01 = (32 - fx) * fy, w11 = fx * fy;
            for (int c = 0; c < ch; c++) {
                auto tap = [&](int xx, int yy) -> int {
                    return xx < 0 || xx >= w || yy < 0 || yy >= h ? 0 : src[((size_t)yy * w + xx) * ch + c];
                };
                const int v = w00 * tap(sx, sy) + w10 * tap(sx + 1, sy) + w01 * tap(sx, sy + 1) + w11 * tap(sx + 1, sy + 1);
                dst[((size_t)yo * ow + xo) * ch + c] = (uint8_t)((v + 512) >> 10);
            }
        }
    }
}

// A whole NV12 frame (w x h Y rows, then h/2 rows of interleaved UV; dense) by the centre-based correction T, cropped by
// `crop` (even) on every side: Y as vo_warp_bgr warps a frame, UV as a (w/2) x (h/2) two-channel image by the same
// similarity with half the translation.
void vo_warp_nv12(const uint8_t* src, int w, int h, const double T[4], uint8_t* dst, int crop)
{
    const int ow = w - 2 * crop, oh = h - 2 * crop;
    auto matrix = [](const double t[4], int pw, int ph, double M[6]) {
        const double A = t[0], B = t[1], cx = (pw - 1) * 0.5, cy = (ph - 1) * 0.5;
        M[0] = 1.0 + A; M[1] = -B; M[2] = t[2] - A * cx + B * cy; M[3] = B; M[4] = 1.0 + A; M[5] = t[3] - B * cx - A * cy;
    };
    double M[6];
    matrix(T, w, h, M);
    vo_warp_plane_matrix(src, w, h, 1, M, dst, ow, oh, crop, crop);
    const double Tuv[4] = {T[0], T[1], T[2] * 0.5, T[3] * 0.5};
    matrix(Tuv, w / 2, h / 2, M);
    vo_warp_plane_matrix(src + (size_t)w * h, w / 2, h / 2, 2, M, dst + (size_t)ow * oh, ow / 2, oh / 2, crop / 2, crop / 2);
}

//------------------------------------------------------------------------------
// Transform algebra — imgproc.cpp:333-437.

void vo_tf_inverse(const double T[4], double out[4])   // imgproc.cpp:333-359
{
    double p = 1.0 + T[0], q = T[1];
    double denom = p * p + q * q;
    double a = (p / denom) - 1.0;
    double b = -q / denom;
    double ix = (-p * T[2] - q * T[3]) / denom;
    double iy = (q * T[2] - p * T[3]) / denom;
    out[0] = a; out[1] = b; out[2] = ix; out[3] = iy;
}

void vo_tf_compose(const double T1[4], const double T2[4], double out[4])   // imgproc.cpp:361-387
{
    double p1 = 1.0 + T1[0], q1 = T1[1];
    double p2 = 1.0 + T2[0], q2 = T2[1];
    double A3 = (p2 * p1 - q2 * q1) - 1.0;
    double B3 = (p2 * q1 + q2 * p1);
    double TX3 = p2 * T1[2] - q2 * T1[3] + T2[2];
    double TY3 = q2 * T1[2] + p2 * T1[3] + T2[3];
    out[0] = A3; out[1] = B3; out[2] = TX3; out[3] = TY3;
}

void vo_tf_warp(const double T[4], double px, double py, double out[2])   // imgproc.cpp:389-394
{
    out[0] = (1 + T[0]) * px - T[1] * py + T[2];
    out[1] = T[1] * px + (1 + T[0]) * py + T[3];
}

void vo_tf_warp_center(const double T[4], double px, double py, double cx, double cy, double out[2])
{   // imgproc.cpp:401-411
    double x = px - cx, y = py - cy;
    out[0] = (1 + T[0]) * x - T[1] * y + cx + T[2];
    out[1] = T[1] * x + (1 + T[0]) * y + cy + T[3];
}

static double dist2(const double a[2], const double b[2])   // imgproc.cpp:413-417
{
    double dx = a[0] - b[0], dy = a[1] - b[1];
    return std::sqrt(dx * dx + dy * dy);
}

double vo_tf_max_corner_displacement(const double T[4], double w, double h)   // imgproc.cpp:419-437
{
    double cx = w * 0.5, cy = h * 0.5;
    const double c[4][2] = {{0.0, 0.0}, {w, 0.0}, {0.0, h}, {w, h}};
    double max_d = 0.0;
    for (int i = 0; i < 4; i++) {
        double o[2];
        vo_tf_warp_center(T, c[i][0], c[i][1], cx, cy, o);
        max_d = std::max(max_d, dist2(o, c[i]));
    }
    return max_d;
}

//------------------------------------------------------------------------------
// cv::SVD(H) (alignment.cpp:558) and H.inv(DECOMP_SVD) (alignment.cpp:582) for 4x4 f64.
// Restates OpenCV's one-sided Jacobi (modules/core/src/lapack.cpp, JacobiSVDImpl_<double>):
// rows of At are rotated pairwise until |p| <= eps*sqrt(a*b), eps = 10*DBL_EPSILON,
// at most max(m,30) sweeps; singular values sorted descending.  inv(DECOMP_SVD) is
// SVD::backSubst against the identity with threshold 2*DBL_EPSILON*sum(w).
void vo_svd4(const double H[16], double w[4], double u[16], double vt[16])
{
    const int n = 4;
    const double eps = 2.220446049250313e-16 * 10;
    const double minval = 2.2250738585072014e-308;
    double At[16], Vt[16], W[4];
    // _SVDcompute transposes src into temp_a: At row i = column i of H
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++) At[i * n + k] = H[k * n + i];
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < n; k++) { double t = At[i * n + k]; sd += t * t; }
        W[i] = sd;
        for (int k = 0; k < n; k++) Vt[i * n + k] = 0;
        Vt[i * n + i] = 1;
    }
    for (int iter = 0; iter < 30; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double* Ai = At + i * n; double* Aj = At + j * n;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < n; k++) p += Ai[k] * Aj[k];
                if (std::fabs(p) <= eps * std::sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = std::sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = std::sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < n; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double* Vi = Vt + i * n; double* Vj = Vt + j * n;
                for (int k = 0; k < n; k++) {
                    double t0 = c * Vi[k] + s * Vj[k];
                    double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < n; k++) { double t = At[i * n + k]; sd += t * t; }
        W[i] = std::sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            std::swap(W[i], W[j]);
            for (int k = 0; k < n; k++) std::swap(At[i * n + k], At[j * n + k]);
            for (int k = 0; k < n; k++) std::swap(Vt[i * n + k], Vt[j * n + k]);
        }
    }
    for (int i = 0; i < n; i++) {
        w[i] = W[i];
        // OpenCV substitutes a random orthogonal vector for zero singular values; a
        // zero column is kept here (H is SPD in practice, this branch is not reached).
        double s = W[i] > minval ? 1 / W[i] : 0.;
        for (int k = 0; k < n; k++) At[i * n + k] *= s;
    }
    // u = At^T (columns are left singular vectors), vt = Vt
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++) { u[k * n + i] = At[i * n + k]; vt[i * n + k] = Vt[i * n + k]; }
}

void vo_inv4_svd(const double H[16], double Hinv[16])
{
    double w[4], u[16], vt[16];
    vo_svd4(H, w, u, vt);
    double threshold = 0;
    for (int i = 0; i < 4; i++) threshold += w[i];
    threshold *= 2.220446049250313e-16 * 2;
    for (int i = 0; i < 16; i++) Hinv[i] = 0;
    // x = sum_i v_i (u_i^T b) / w_i with b = I
    for (int i = 0; i < 4; i++) {
        if (std::fabs(w[i]) <= threshold) continue;
        double wi = 1 / w[i];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
                Hinv[r * 4 + c] += vt[i * 4 + r] * (u[c * 4 + i] * wi);
    }
}

//------------------------------------------------------------------------------
// alignment.cpp:438-486 — DeltaPixel array in row-major tile order, keep the
// (size_t)(N * fraction) smallest by std::nth_element on abs_delta (libstdc++ introselect;
// the tie subset is whatever libstdc++ produces, hence the direct call).
namespace {
struct DeltaPixel { uint16_t abs_delta, tile_x, tile_y; };   // alignment.hpp:84-87
}

static size_t select_smallest(const uint16_t* abs_delta, int tw, int th, float fraction,
                              std::vector<DeltaPixel>& dp)
{
    dp.clear();
    for (int j = 0; j < th; j++)
        for (int k = 0; k < tw; k++) {
            DeltaPixel d;
            d.abs_delta = abs_delta[(size_t)j * tw + k];
            d.tile_x = (uint16_t)k;
            d.tile_y = (uint16_t)j;
            dp.push_back(d);
        }
    const size_t count = static_cast<size_t>(dp.size() * fraction);
    std::nth_element(dp.begin(), dp.begin() + count, dp.end(),
                     [](const DeltaPixel& l, const DeltaPixel& r) { return l.abs_delta < r.abs_delta; });
    dp.resize(count);
    return count;
}

int vo_select_smallest(const uint16_t* abs_delta, int n, float fraction, uint32_t* order)
{
    // n tiles in scan order; the linear index rides in (tile_y, tile_x) = (i >> 15, i & 32767), which std::nth_element
    // only carries along (the comparator reads abs_delta)
    std::vector<DeltaPixel> dp(n);
    for (int i = 0; i < n; i++) { dp[i].abs_delta = abs_delta[i]; dp[i].tile_x = (uint16_t)(i & 32767); dp[i].tile_y = (uint16_t)(i >> 15); }
    const size_t k = static_cast<size_t>(dp.size() * fraction);
    std::nth_element(dp.begin(), dp.begin() + k, dp.end(),
                     [](const DeltaPixel& l, const DeltaPixel& r) { return l.abs_delta < r.abs_delta; });
    for (size_t i = 0; i < k; i++) order[i] = ((uint32_t)dp[i].tile_y << 15) | dp[i].tile_x;
    return (int)k;
}

// Test aid: libstdc++'s own std::__introselect with an explicit depth limit, so that the
// heap-select fallback (depth limit exhausted) can be compared against the GPU replay.
int vo_introselect_depth(const uint16_t* abs_delta, int n, int nth, int depth, uint32_t* order)
{
    std::vector<DeltaPixel> dp(n);
    for (int i = 0; i < n; i++) { dp[i].abs_delta = abs_delta[i]; dp[i].tile_x = (uint16_t)i; dp[i].tile_y = 0; }
    auto comp = [](const DeltaPixel& l, const DeltaPixel& r) { return l.abs_delta < r.abs_delta; };
    if (n > 0 && nth < n)
        std::__introselect(dp.begin(), dp.begin() + nth, dp.end(), depth, __gnu_cxx::__ops::__iter_comp_iter(comp));
    for (int i = 0; i < n; i++) order[i] = dp[i].tile_x;
    return n;
}

//------------------------------------------------------------------------------
// cv::phaseCorrelate (alignment.cpp:374; OpenCV imgproc/src/phasecorr.cpp, not under /root/reference) restated
// as a separable two-stage DFT in f64 (pc_dft below) with every sum taken serially in ascending index order and no FMA — the
// canonical order the CUDA kernels reproduce bit for bit.  OpenCV runs its own f32 mixed-radix FFT: its ulps are
// not reproduced, its result is (tests/golden/phase_correlate.npz pins shift and response against cv2 4.13 to 1e-4):
//   M, N = getOptimalDFTSize(rows), (cols); zero padding on the right / bottom (copyMakeBorder);
//   P = F1 conj(F2) (mulSpectrums conjB), C = P |P| / (|P|^2 + FLT_EPSILON) (magSpectrums, divSpectrums),
//   R = unnormalised inverse DFT of C, fftShift, first maximum (minMaxLoc), 5x5 weighted centroid clamped to the
//   array (response = window sum / (M N)), shift = (N/2, M/2) - centroid.
int vo_optimal_dft_size(int n)
{
    for (;; n++) {
        int m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return n;
    }
}

static void pc_twiddles(int n, std::vector<double>& c, std::vector<double>& s)
{
    c.resize(n); s.resize(n);
    for (int j = 0; j < n; j++) {
        const double a = 2.0 * M_PI * (double)j / (double)n;
        c[j] = cos(a); s[j] = -sin(a);
    }
}

// Every 1-D transform of length L is evaluated in two stages (L = L1 L2, L1 the largest divisor of L with L1^2 <= L;
// input index n = L2 n1 + n2, output index k = k1 + L1 k2):
//     A[k1][n2]    = sum_{n1 < L1} x[L2 n1 + n2] W^(L2 n1 k1)        (terms beyond the valid length are skipped)
//     B[k1][n2]    = A[k1][n2] W^(n2 k1)
//     X[k1 + L1 k2] = sum_{n2 < L2} B[k1][n2] W^(L1 n2 k2)
// with W^j = (c[j mod L], s[j mod L]) from the table (the conjugate for an inverse), every sum serial in ascending index
// order, every product and sum individually rounded: L (L1 + L2) multiply-adds instead of L^2, and an order the kernels
// (vs_phasecorr.cu) reproduce bit for bit.
static void pc_split(int L, int& L1, int& L2)
{
    L1 = 1;
    for (int d = 1; d * d <= L; d++)
        if (L % d == 0) L1 = d;
    L2 = L / L1;
}

// one transform: in[n] = (re, im) for n < valid at stride `is` (complex elements), zero beyond; out[k] for k < kout at
// stride `os`; real_in: the input has no imaginary part (in points at doubles, stride `is` doubles); conj: inverse twiddles;
// real_out: only the real parts are produced (out points at doubles)
static void pc_dft(const double* in, size_t is, int valid, bool real_in, int L, const std::vector<double>& c, const std::vector<double>& s,
                   bool conj, double* out, size_t os, int kout, bool real_out, std::vector<double>& B)
{
    int L1, L2;
    pc_split(L, L1, L2);
    B.resize((size_t)L * 2);
    for (int k1 = 0; k1 < L1; k1++)
        for (int n2 = 0; n2 < L2; n2++) {
            double re = 0.0, im = 0.0;
            const int step = (int)(((long long)L2 * k1) % L);
            int j = 0;
            for (int n1 = 0; n1 < L1; n1++) {
                const int n = L2 * n1 + n2;
                if (n < valid) {
                    if (real_in) {
                        const double x = in[(size_t)n * is];
                        re = re + x * c[j];
                        im = conj ? im - x * s[j] : im + x * s[j];
                    } else {
                        const double ar = in[(size_t)n * is * 2], ai = in[(size_t)n * is * 2 + 1];
                        const double t1 = ar * c[j], t2 = ai * s[j], t3 = ar * s[j], t4 = ai * c[j];
                        if (conj) { re = re + (t1 + t2); im = im + (t4 - t3); }
                        else      { re = re + (t1 - t2); im = im + (t3 + t4); }
                    }
                }
                j += step; if (j >= L) j -= L;
            }
            const int jt = (int)(((long long)n2 * k1) % L);
            const double t1 = re * c[jt], t2 = im * s[jt], t3 = re * s[jt], t4 = im * c[jt];
            if (conj) { B[((size_t)k1 * L2 + n2) * 2] = t1 + t2; B[((size_t)k1 * L2 + n2) * 2 + 1] = t4 - t3; }
            else      { B[((size_t)k1 * L2 + n2) * 2] = t1 - t2; B[((size_t)k1 * L2 + n2) * 2 + 1] = t3 + t4; }
        }
    for (int k = 0; k < kout; k++) {
        const int k1 = k % L1, k2 = k / L1;
        double re = 0.0, im = 0.0;
        const int step = (int)(((long long)L1 * k2) % L);
        int j = 0;
        for (int n2 = 0; n2 < L2; n2++) {
            const double ar = B[((size_t)k1 * L2 + n2) * 2], ai = B[((size_t)k1 * L2 + n2) * 2 + 1];
            const double t1 = ar * c[j], t2 = ai * s[j], t3 = ar * s[j], t4 = ai * c[j];
            if (conj) { re = re + (t1 + t2); im = im + (t4 - t3); }
            else      { re = re + (t1 - t2); im = im + (t3 + t4); }
            j += step; if (j >= L) j -= L;
        }
        if (real_out) out[(size_t)k * os] = re;
        else { out[(size_t)k * os * 2] = re; out[(size_t)k * os * 2 + 1] = im; }
    }
}

// forward transform of a real w x h image padded to N x M: G[m][k], k < N/2+1 (interleaved re, im)
static void pc_forward(const float* img, int64_t stride, int w, int h, int M, int N,
                       const std::vector<double>& cN, const std::vector<double>& sN,
                       const std::vector<double>& cM, const std::vector<double>& sM, std::vector<double>& G)
{
    const int Kh = N / 2 + 1;
    std::vector<double> F((size_t)h * Kh * 2), row((size_t)w), B;
    for (int r = 0; r < h; r++) {
        for (int n = 0; n < w; n++) row[n] = (double)img[(size_t)r * stride + n];
        pc_dft(row.data(), 1, w, true, N, cN, sN, false, &F[(size_t)r * Kh * 2], 1, Kh, false, B);
    }
    G.assign((size_t)M * Kh * 2, 0.0);
    for (int k = 0; k < Kh; k++)
        pc_dft(&F[(size_t)k * 2], (size_t)Kh, h, false, M, cM, sM, false, &G[(size_t)k * 2], (size_t)Kh, M, false, B);
}

void vo_phase_correlate(const float* img1, const float* img2, int w, int h, int64_t stride, double out[3])
{
    const int M = vo_optimal_dft_size(h), N = vo_optimal_dft_size(w), Kh = N / 2 + 1;
    std::vector<double> cN, sN, cM, sM, G1, G2, B;
    pc_twiddles(N, cN, sN);
    pc_twiddles(M, cM, sM);
    pc_forward(img1, stride, w, h, M, N, cN, sN, cM, sM, G1);
    pc_forward(img2, stride, w, h, M, N, cN, sN, cM, sM, G2);
    // cross-power spectrum
    std::vector<double> C((size_t)M * Kh * 2), D((size_t)M * Kh * 2);
    for (size_t i = 0; i < (size_t)M * Kh; i++) {
        const double ar = G1[2 * i], ai = G1[2 * i + 1], br = G2[2 * i], bi = G2[2 * i + 1];
        const double pr = ar * br + ai * bi, pi = ai * br - ar * bi;
        const double mag = sqrt(pr * pr + pi * pi);
        const double den = mag * mag + (double)FLT_EPSILON;
        C[2 * i] = (pr * mag) / den; C[2 * i + 1] = (pi * mag) / den;
    }
    // inverse along the columns (conjugate twiddles)
    for (int k = 0; k < Kh; k++)
        pc_dft(&C[(size_t)k * 2], (size_t)Kh, M, false, M, cM, sM, true, &D[(size_t)k * 2], (size_t)Kh, M, false, B);
    // inverse along the rows to a real surface: the row is completed by its Hermitian half, Y[N - k] = conj(Y[k]), and
    // only the real parts of the result are formed; then the first maximum of the shifted surface
    std::vector<double> R((size_t)M * N), Y((size_t)N * 2);
    for (int r = 0; r < M; r++) {
        const double* d = &D[(size_t)r * Kh * 2];
        for (int k = 0; k < N; k++) {
            if (k < Kh) { Y[2 * k] = d[2 * k]; Y[2 * k + 1] = d[2 * k + 1]; }
            else { Y[2 * k] = d[2 * (N - k)]; Y[2 * k + 1] = -d[2 * (N - k) + 1]; }
        }
        pc_dft(Y.data(), 1, N, false, N, cN, sN, true, &R[(size_t)r * N], 1, N, true, B);
    }
    auto shifted = [&](int y, int x) {   // fftShift: element (r, n) moves to ((r + M/2) % M, (n + N/2) % N)
        const int r = (y - M / 2 + M) % M, n = (x - N / 2 + N) % N;
        return R[(size_t)r * N + n];
    };
    int py = 0, px = 0;
    double best = shifted(0, 0);
    for (int y = 0; y < M; y++)
        for (int x = 0; x < N; x++) {
            const double v = shifted(y, x);
            if (v > best) { best = v; py = y; px = x; }
        }
    int minr = py - 2, maxr = py + 2, minc = px - 2, maxc = px + 2;
    if (minr < 0) minr = 0;
    if (minc < 0) minc = 0;
    if (maxr > M - 1) maxr = M - 1;
    if (maxc > N - 1) maxc = N - 1;
    double cx = 0.0, cy = 0.0, sum = 0.0;
    for (int y = minr; y <= maxr; y++)
        for (int x = minc; x <= maxc; x++) {
            const double v = shifted(y, x);
            cx = cx + (double)x * v;
            cy = cy + (double)y * v;
            sum = sum + v;
        }
    const double response = sum / (double)(M * N);
    sum = sum + DBL_EPSILON;
    cx = cx / sum; cy = cy / sum;
    out[0] = (double)N / 2.0 - cx;
    out[1] = (double)M / 2.0 - cy;
    out[2] = response;
}

void vo_phase_correlate_u8(const uint8_t* img1, const uint8_t* img2, int w, int h, double out[3])
{
    std::vector<float> a((size_t)w * h), b((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) { a[i] = (float)img1[i]; b[i] = (float)img2[i]; }
    vo_phase_correlate(a.data(), b.data(), w, h, w, out);
}

//------------------------------------------------------------------------------
// VideoAligner — alignment.cpp:149-704.

void vo_align_params_default(vo_align_params* p)   // alignment.hpp:5-41
{
    p->phase_correlate = 0;
    p->phase_correlate_threshold = 0.5;
    p->threshold = 0.02;
    p->smallest_fraction = 0.8f;
    p->max_iters = 64;
    p->pyramid_min_width = 20;
    p->pyramid_min_height = 20;
    p->max_displacement = 10.0;
}

struct vo_aligner {
    int curr = 0, prev = 1, accumulated = 0;      // alignment.hpp:62-64
    int levels = -1;
    int last_w = -1, last_h = -1;
    struct Level {
        int w = 0, h = 0, tile = 0, tw = 0, th = 0;
        std::vector<uint8_t> img[2];
        std::vector<float> gx, gy;
        std::vector<uint16_t> amx, amy;
        std::vector<float> jx, jy;
        std::vector<uint16_t> wdx, wdy;
        std::vector<uint32_t> ordx, ordy;
        int iters = 0;
    };
    std::vector<Level> lv;
    double phase[3] = {0, 0, 0};                  // last cv::phaseCorrelate shift (x, y) and response
};

vo_aligner* vo_aligner_create(void) { return new vo_aligner(); }
void vo_aligner_destroy(vo_aligner* a) { delete a; }

// alignment.cpp:149-235
static bool compute_pyramid(vo_aligner* a, const uint8_t* bgr, int width, int height,
                            const vo_align_params* params)
{
    if (a->lv.empty() || width != a->last_w || height != a->last_h) {
        a->curr = 0; a->prev = 1; a->accumulated = 0;
        a->last_w = width; a->last_h = height;
        int w = width, h = height;
        a->levels = 0;
        do { a->levels++; w /= 2; h /= 2; }
        while (w >= params->pyramid_min_width && h >= params->pyramid_min_height);
        a->lv.assign(a->levels, vo_aligner::Level());
        w = width; h = height;
        for (int i = 0; i < a->levels; i++) {
            if (i > 0) { w /= 2; h /= 2; }
            auto& L = a->lv[i];
            L.w = w; L.h = h;
            L.img[0].assign((size_t)w * h, 0);
            L.img[1].assign((size_t)w * h, 0);
            L.gx.assign((size_t)w * h, 0.f);
            L.gy.assign((size_t)w * h, 0.f);
        }
    } else {
        a->prev = a->curr;
        a->curr ^= 1;
    }
    vo_bgr2gray(bgr, width, height, a->lv[0].img[a->curr].data());
    for (int i = 1; i < a->levels; i++)
        vo_pyr_down(a->lv[i - 1].img[a->curr].data(), a->lv[i - 1].w, a->lv[i - 1].h,
                    a->lv[i].img[a->curr].data(), a->lv[i].w, a->lv[i].h);
    // alignment.cpp:225-229: PhaseImage = level 2 as f32 — u8 values are exact in f32, vo_phase_correlate_u8 converts on use.
    if (a->accumulated >= 2) return true;
    return ++a->accumulated >= 2;
}

// alignment.cpp:237-276
static void compute_keyframe(vo_aligner* a)
{
    for (int i = 0; i < a->levels; i++) {
        auto& L = a->lv[i];
        vo_grad_xy(L.img[a->curr].data(), L.w, L.h, L.gx.data(), L.gy.data(), L.w, L.h);
        L.tile = vo_tile_size(L.w, L.h);
        L.tw = L.w / L.tile; L.th = L.h / L.tile;
        L.amx.assign((size_t)L.tw * L.th * 2, 0);
        L.amy.assign((size_t)L.tw * L.th * 2, 0);
        vo_grad_argmax(L.gx.data(), L.gy.data(), L.w, L.h, L.tile, L.amx.data(), L.amy.data());
        L.jx.assign((size_t)L.tw * L.th * 4, 0.f);
        L.jy.assign((size_t)L.tw * L.th * 4, 0.f);
        vo_sparse_jac(L.gx.data(), L.gy.data(), L.w, L.h, L.amx.data(), L.amy.data(), L.tw, L.th,
                      L.jx.data(), L.jy.data());
    }
}

// alignment.cpp:334-704
int vo_aligner_align(vo_aligner* a, const uint8_t* bgr, int w, int h,
                     const vo_align_params* params, double Tout[4])
{
    const int KeyframeIndex = 1, NonKeyframeIndex = 0;   // alignment.hpp:65-66
    double T[4] = {0, 0, 0, 0};
    Tout[0] = Tout[1] = Tout[2] = Tout[3] = 0;
    for (auto& L : a->lv) L.iters = 0;

    if (!compute_pyramid(a, bgr, w, h, params)) return 0;
    if (a->curr == KeyframeIndex) compute_keyframe(a);
    // alignment.cpp:369-388 — translation seeded from cv::phaseCorrelate of the level-2 images
    if (params->phase_correlate) {
        const int PhaseLevel = 2;                          // alignment.hpp:69
        if (a->levels <= PhaseLevel) return 0;             // (the reference indexes past its pyramid here)
        auto& PL = a->lv[PhaseLevel];
        double pc[3];
        vo_phase_correlate_u8(PL.img[a->prev].data(), PL.img[a->curr].data(), PL.w, PL.h, pc);
        a->phase[0] = pc[0]; a->phase[1] = pc[1]; a->phase[2] = pc[2];
        if (pc[2] > params->phase_correlate_threshold) {
            const float phase_layer_scale = (1 << PhaseLevel) / float(1 << a->levels);
            T[2] = pc[0] * phase_layer_scale;
            T[3] = pc[1] * phase_layer_scale;
            if (a->curr == KeyframeIndex) { T[2] = -T[2]; T[3] = -T[3]; }
        }
    }

    std::vector<DeltaPixel> dpx, dpy;
    for (int i = a->levels - 1; i >= 0; i--) {
        auto& L = a->lv[i];
        const uint8_t* tmpl = L.img[NonKeyframeIndex].data();
        const uint8_t* key = L.img[KeyframeIndex].data();
        const int iw = L.w, ih = L.h;
        size_t plane = (size_t)L.tw * L.th;

        L.wdx.assign(plane, 0); L.wdy.assign(plane, 0);
        vo_sparse_warpdiff(tmpl, key, iw, ih, L.amx.data(), L.tw, L.th, T, L.wdx.data());
        vo_sparse_warpdiff(tmpl, key, iw, ih, L.amy.data(), L.tw, L.th, T, L.wdy.data());

        size_t kx = select_smallest(L.wdx.data(), L.tw, L.th, params->smallest_fraction, dpx);
        size_t ky = select_smallest(L.wdy.data(), L.tw, L.th, params->smallest_fraction, dpy);

        // alignment.cpp:526-545 — gather coords and Jacobians in post-nth_element order
        std::vector<uint16_t> selx(kx * 2), sely(ky * 2);
        std::vector<float> sjx(kx * 4), sjy(ky * 4);
        L.ordx.resize(kx); L.ordy.resize(ky);
        for (size_t j = 0; j < kx; j++) {
            size_t t = (size_t)dpx[j].tile_y * L.tw + dpx[j].tile_x;
            L.ordx[j] = (uint32_t)t;
            selx[j] = L.amx[t]; selx[kx + j] = L.amx[plane + t];
            for (int c = 0; c < 4; c++) sjx[c * kx + j] = L.jx[c * plane + t];
        }
        for (size_t j = 0; j < ky; j++) {
            size_t t = (size_t)dpy[j].tile_y * L.tw + dpy[j].tile_x;
            L.ordy[j] = (uint32_t)t;
            sely[j] = L.amy[t]; sely[ky + j] = L.amy[plane + t];
            for (int c = 0; c < 4; c++) sjy[c * ky + j] = L.jy[c * plane + t];
        }

        // alignment.cpp:278-332 — H = sum j j^T, upper triangle then mirrored, X rows then Y rows
        double H[16] = {0};
        for (int axis = 0; axis < 2; axis++) {
            const std::vector<float>& J = axis == 0 ? sjx : sjy;
            size_t m = axis == 0 ? kx : ky;
            for (size_t p = 0; p < m; p++) {
                double j[4] = {J[0 * m + p], J[1 * m + p], J[2 * m + p], J[3 * m + p]};
                for (int r = 0; r < 4; r++)
                    for (int c = r; c < 4; c++) H[r * 4 + c] += j[r] * j[c];
            }
        }
        for (int r = 0; r < 4; r++)
            for (int c = r + 1; c < 4; c++) H[c * 4 + r] = H[r * 4 + c];

        // alignment.cpp:554-583 — SVD condition check, Tikhonov, SVD inverse
        {
            double sw[4], su[16], svt[16];
            vo_svd4(H, sw, su, svt);
            double min_sv = sw[3], max_sv = sw[0];
            double cond = max_sv / (min_sv + 1e-10);
            if (cond > 1e6) {
                double lambda = 1e-6 * max_sv;
                for (int d = 0; d < 4; d++) H[d * 4 + d] += lambda;
            }
        }
        double Hinv[16];
        vo_inv4_svd(H, Hinv);

        // alignment.cpp:585-668
        double cxi = iw * 0.5, cyi = ih * 0.5;
        const double corners[4][2] = {{0.0, 0.0}, {(double)(iw - 1.f), 0.0},
                                      {0.0, (double)(ih - 1.f)}, {(double)(iw - 1.f), (double)(ih - 1.f)}};
        double c0[4][2], c1[4][2];
        for (int c = 0; c < 4; c++) {
            vo_tf_warp_center(T, corners[c][0], corners[c][1], cxi, cyi, c0[c]);
            c1[c][0] = c0[c][0]; c1[c][1] = c0[c][1];
        }

        for (int iter = 0; iter < params->max_iters; iter++) {
            L.iters++;
            double b[4];
            vo_sparse_ica(tmpl, key, iw, ih, selx.data(), (int)kx, sely.data(), (int)ky,
                          sjx.data(), sjy.data(), T, b);
            double dt[4];
            for (int r = 0; r < 4; r++) {
                // cv::Mat product of a 4x4 by a 4x1 (alignment.cpp:624): row dot product
                double s = 0;
                for (int c = 0; c < 4; c++) s += Hinv[r * 4 + c] * b[c];
                dt[r] = s;
            }
            double scale = 1.0 / iw;
            double delta[4] = {dt[0] * scale, dt[1] * scale, dt[2], dt[3]};
            double Tn[4];
            vo_tf_compose(delta, T, Tn);   // delta first, then T (alignment.cpp:639)
            T[0] = Tn[0]; T[1] = Tn[1]; T[2] = Tn[2]; T[3] = Tn[3];

            double c2[4][2];
            for (int c = 0; c < 4; c++) vo_tf_warp_center(T, corners[c][0], corners[c][1], cxi, cyi, c2[c]);
            double ud12 = std::max(dist2(c2[0], c1[0]), dist2(c2[1], c1[1]));
            double ld12 = std::max(dist2(c2[2], c1[2]), dist2(c2[3], c1[3]));
            double d12 = std::max(ud12, ld12);
            for (int c = 0; c < 4; c++) { c1[c][0] = c2[c][0]; c1[c][1] = c2[c][1]; }

            if (d12 < params->threshold) break;
            if (iter >= params->max_iters - 1) {
                Tout[0] = T[0]; Tout[1] = T[1]; Tout[2] = T[2]; Tout[3] = T[3];
                return 0;
            }
        }

        double ud01 = std::max(dist2(c0[0], c1[0]), dist2(c0[1], c1[1]));
        double ld01 = std::max(dist2(c0[2], c1[2]), dist2(c0[3], c1[3]));
        double d01 = std::max(ud01, ld01);
        if (d01 > params->max_displacement) {
            Tout[0] = T[0]; Tout[1] = T[1]; Tout[2] = T[2]; Tout[3] = T[3];
            return 0;
        }
        if (i > 0) { T[2] *= 2.0; T[3] *= 2.0; }
    }

    if (a->curr != KeyframeIndex) {
        double Ti[4];
        vo_tf_inverse(T, Ti);
        T[0] = Ti[0]; T[1] = Ti[1]; T[2] = Ti[2]; T[3] = Ti[3];
    }
    Tout[0] = T[0]; Tout[1] = T[1]; Tout[2] = T[2]; Tout[3] = T[3];
    return 1;
}

void vo_aligner_phase(const vo_aligner* a, double out[3]) { out[0] = a->phase[0]; out[1] = a->phase[1]; out[2] = a->phase[2]; }
int vo_aligner_levels(const vo_aligner* a) { return a->levels; }
int vo_aligner_curr_index(const vo_aligner* a) { return a->curr; }
void vo_aligner_level_info(const vo_aligner* a, int level, int* w, int* h, int* tile, int* tw, int* th)
{
    const auto& L = a->lv[level];
    *w = L.w; *h = L.h; *tile = L.tile; *tw = L.tw; *th = L.th;
}
const uint8_t* vo_aligner_pyramid(const vo_aligner* a, int slot, int level) { return a->lv[level].img[slot].data(); }
const uint16_t* vo_aligner_keypoints(const vo_aligner* a, int level, int axis)
{ return axis == 0 ? a->lv[level].amx.data() : a->lv[level].amy.data(); }
const float* vo_aligner_jacobians(const vo_aligner* a, int level, int axis)
{ return axis == 0 ? a->lv[level].jx.data() : a->lv[level].jy.data(); }
const uint16_t* vo_aligner_warpdiff(const vo_aligner* a, int level, int axis)
{ return axis == 0 ? a->lv[level].wdx.data() : a->lv[level].wdy.data(); }
int vo_aligner_selected(const vo_aligner* a, int level, int axis, const uint32_t** order)
{
    const auto& v = axis == 0 ? a->lv[level].ordx : a->lv[level].ordy;
    *order = v.data();
    return (int)v.size();
}
int vo_aligner_iterations(const vo_aligner* a, int level) { return a->lv[level].iters; }

//------------------------------------------------------------------------------
// smoother.cpp:18-64 — 100 iterations of relaxation towards the data followed by a
// sequential in-place pairwise TV shrink (order matters).
void vo_tvl1_smooth(const double* data, int n, double lambda, int iterations, double* x)
{
    for (int i = 0; i < n; i++) x[i] = data[i];
    for (int iter = 0; iter < iterations; ++iter) {
        for (int i = 0; i < n; i++) {
            double alpha = 0.5;
            x[i] = (1.0 - alpha) * x[i] + alpha * data[i];
        }
        for (int i = 0; i + 1 < n; i++) {
            double diff = x[i + 1] - x[i];
            double mag = std::fabs(diff);
            if (mag > lambda) {
                double shrink = (mag - lambda) / mag * 0.5;
                x[i] += diff * shrink;
                x[i + 1] -= diff * shrink;
            } else {
                double mid = 0.5 * (x[i] + x[i + 1]);
                x[i] = mid;
                x[i + 1] = mid;
            }
        }
    }
}

// smoother.cpp:66-127
struct vo_smoother {
    int lag_behind, lag_ahead;
    double lambda;
    int next = 0;
    std::vector<std::array<double, 4>> meas;
};

vo_smoother* vo_smoother_create(int lag_behind, int lag_ahead, double lambda)
{
    auto* s = new vo_smoother();
    s->lag_behind = lag_behind; s->lag_ahead = lag_ahead; s->lambda = lambda;
    return s;
}
void vo_smoother_destroy(vo_smoother* s) { delete s; }

int vo_smoother_update(vo_smoother* s, const double m[4], double out[4])
{
    s->meas.push_back({m[0], m[1], m[2], m[3]});
    const int newest = (int)s->meas.size() - 1;
    if (s->next + s->lag_ahead > newest) return 0;
    int start = std::max(0, s->next - s->lag_behind);
    int end = s->next + s->lag_ahead;
    int n = end - start + 1;
    std::vector<double> v(n), o(n);
    int middle = s->next - start;
    for (int c = 0; c < 4; c++) {
        for (int i = 0; i < n; i++) v[i] = s->meas[start + i][c];
        vo_tvl1_smooth(v.data(), n, s->lambda, 100, o.data());
        out[c] = o[middle];
    }
    s->next++;
    return 1;
}

// stabilizer.hpp:13-30
void vo_stab_params_default(vo_stab_params* p)
{
    vo_align_params_default(&p->aligner);
    p->lag = 10; p->smoother_memory = 5; p->lambda = 4.0;
    p->enable_smoother = 1; p->crop_pixels = 32;
    p->min_disp = 48.0; p->max_disp = 64.0;
    p->min_decay = 0.9; p->max_decay = 0.7;
}

// stabilizer.cpp:3-117
struct vo_stabilizer {
    vo_stab_params params;
    vo_aligner* aligner;
    vo_smoother* smoother;
    int frame_index = 0;
    std::deque<std::array<double, 4>> meas;
    std::deque<std::vector<uint8_t>> frames;
    std::deque<std::pair<int, int>> frame_sizes;   // each buffered frame keeps its own size (stabilizer.cpp:15,91-99)
    double accum[4] = {0, 0, 0, 0};
};

vo_stabilizer* vo_stabilizer_create(const vo_stab_params* p)
{
    auto* s = new vo_stabilizer();
    s->params = *p;
    s->aligner = vo_aligner_create();
    s->smoother = vo_smoother_create(p->lag, p->smoother_memory, p->lambda);   // stabilizer.cpp:4
    return s;
}
void vo_stabilizer_destroy(vo_stabilizer* s)
{
    vo_aligner_destroy(s->aligner);
    vo_smoother_destroy(s->smoother);
    delete s;
}

int vo_stabilizer_process(vo_stabilizer* s, const uint8_t* bgr, int w, int h,
                          uint8_t* out, int* out_w, int* out_h,
                          int* meas_ok, double meas_out[4], double correction_out[4])
{
    const vo_stab_params& P = s->params;
    ++s->frame_index;
    s->frames.emplace_back(bgr, bgr + (size_t)w * h * 3);
    s->frame_sizes.emplace_back(w, h);

    double cur[4];
    int success = vo_aligner_align(s->aligner, bgr, w, h, &P.aligner, cur);
    if (meas_ok) *meas_ok = success;
    if (meas_out) for (int c = 0; c < 4; c++) meas_out[c] = cur[c];

    double smoothed[4] = {0, 0, 0, 0};
    if (P.enable_smoother) vo_smoother_update(s->smoother, cur, smoothed);
    if (!success) s->accum[0] = s->accum[1] = s->accum[2] = s->accum[3] = 0;
    s->meas.push_back({cur[0], cur[1], cur[2], cur[3]});

    if (s->meas.size() <= (size_t)P.lag) return 0;

    std::array<double, 4> earliest = s->meas.front();
    s->meas.pop_front();
    double jitter[4];
    if (P.enable_smoother) {
        double inv[4];
        vo_tf_inverse(smoothed, inv);
        vo_tf_compose(earliest.data(), inv, jitter);
    } else {
        for (int c = 0; c < 4; c++) jitter[c] = earliest[c];
    }
    double na[4];
    vo_tf_compose(s->accum, jitter, na);
    double displacement = vo_tf_max_corner_displacement(na, w, h);
    double decay = 1.0;
    if (displacement > P.max_disp) {
        decay = P.max_decay;
    } else if (displacement > P.min_disp) {
        double f = (displacement - P.min_disp) / (P.max_disp - P.min_disp);
        f = std::max(0.0, std::min(1.0, f));
        decay = P.min_decay * (1.0 - f) + P.max_decay * f;
    } else {
        decay = P.min_decay;
    }
    na[2] *= decay; na[3] *= decay; na[0] *= decay; na[1] *= decay;
    for (int c = 0; c < 4; c++) s->accum[c] = na[c];

    if (s->frames.empty()) return 0;
    std::vector<uint8_t> frame = std::move(s->frames.front());
    s->frames.pop_front();
    const int fw = s->frame_sizes.front().first, fh = s->frame_sizes.front().second;
    s->frame_sizes.pop_front();
    double corr[4];
    vo_tf_inverse(na, corr);
    if (correction_out) for (int c = 0; c < 4; c++) correction_out[c] = corr[c];
    int crop = P.crop_pixels > 0 ? P.crop_pixels : 0;
    vo_warp_bgr(frame.data(), fw, fh, corr, out, 0, 0, crop);   // warped at the buffered frame's own size
    *out_w = fw - 2 * crop;
    *out_h = fh - 2 * crop;
    return 1;
}

} // extern "C"
