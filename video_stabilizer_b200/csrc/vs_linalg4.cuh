// vs_linalg4.cuh — 4x4 f64 singular values and SVD pseudo-inverse (host + device).
//
// alignment.cpp:558 builds cv::SVD(H) for the condition number and :582 inverts with
// H.inv(cv::DECOMP_SVD).  OpenCV is not part of this build; the functions below restate its
// published one-sided Jacobi SVD for small matrices (rows of H^T rotated pairwise until
// |p| <= 10*eps*sqrt(a*b), at most 30 sweeps, singular values sorted descending) and
// SVD::backSubst against the identity (singular values <= 2*eps*sum(w) are dropped).
// On the device this runs on one thread per frame pair, once per pyramid level.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define VS_LA_HD __host__ __device__ inline
#else
#define VS_LA_HD inline
#endif

// H: row-major 4x4 (symmetric in practice).  w: 4 singular values, descending.
// u (columns = left singular vectors) and vt (rows = right singular vectors), row-major.
VS_LA_HD void vs_svd4(const double* H, double* w, double* u, double* vt)
{
    const double eps = 2.220446049250313e-16 * 10;
    const double tiny = 2.2250738585072014e-308;
    double At[16], V[16], W[4];
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++) { At[i * 4 + k] = H[k * 4 + i]; V[i * 4 + k] = (i == k) ? 1.0 : 0.0; }
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i * 4 + k] * At[i * 4 + k];
        W[i] = sd;
    }
    for (int sweep = 0; sweep < 30; sweep++) {
        bool changed = false;
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 4; j++) {
                double* Ai = At + i * 4; double* Aj = At + j * 4;
                double a = W[i], b = W[j], p = 0;
                for (int k = 0; k < 4; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < 4; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double* Vi = V + i * 4; double* Vj = V + j * 4;
                for (int k = 0; k < 4; k++) {
                    double t0 = c * Vi[k] + s * Vj[k];
                    double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i * 4 + k] * At[i * 4 + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < 3; i++) {
        int j = i;
        for (int k = i + 1; k < 4; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < 4; k++) {
                t = At[i * 4 + k]; At[i * 4 + k] = At[j * 4 + k]; At[j * 4 + k] = t;
                t = V[i * 4 + k]; V[i * 4 + k] = V[j * 4 + k]; V[j * 4 + k] = t;
            }
        }
    }
    for (int i = 0; i < 4; i++) {
        w[i] = W[i];
        double s = W[i] > tiny ? 1 / W[i] : 0.;
        for (int k = 0; k < 4; k++) { u[k * 4 + i] = At[i * 4 + k] * s; vt[i * 4 + k] = V[i * 4 + k]; }
    }
}

// Hinv = sum_i v_i u_i^T / w_i over singular values above 2*eps*sum(w)
VS_LA_HD void vs_inv4_from_svd(const double* w, const double* u, const double* vt, double* Hinv)
{
    double threshold = 0;
    for (int i = 0; i < 4; i++) threshold += w[i];
    threshold *= 2.220446049250313e-16 * 2;
    for (int i = 0; i < 16; i++) Hinv[i] = 0;
    for (int i = 0; i < 4; i++) {
        if (fabs(w[i]) <= threshold) continue;
        double wi = 1 / w[i];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) Hinv[r * 4 + c] += vt[i * 4 + r] * (u[c * 4 + i] * wi);
    }
}

// alignment.cpp:554-583: condition check, optional Tikhonov, SVD inverse.  H is modified
// when regularised.  Returns the condition number.
VS_LA_HD double vs_condition_and_invert(double* H, double* Hinv)
{
    double w[4], u[16], vt[16];
    vs_svd4(H, w, u, vt);
    double cond = w[0] / (w[3] + 1e-10);
    if (cond > 1e6) {
        double lambda = 1e-6 * w[0];
        for (int d = 0; d < 4; d++) H[d * 4 + d] += lambda;
    }
    // H.inv(DECOMP_SVD) decomposes H again.  When H was not regularised that second
    // decomposition is the same deterministic computation on the same input, so its result is
    // reused instead of recomputed (identical bits, half the serial work).
    if (cond > 1e6) vs_svd4(H, w, u, vt);
    vs_inv4_from_svd(w, u, vt, Hinv);
    return cond;
}

#ifdef __CUDACC__
// ---- the same computation spread over four lanes (device only) ---------------------------------------
// vs_svd4 on one thread keeps At, V and W in local memory (dynamic row indices) and takes ~20 us per
// call, which sits on the solver's critical path once per pyramid level.  Here lane k of a quad owns
// column k of At and of V in registers (the (i, j) loops are unrolled, so every index is static); the
// three dot products of a rotation are summed across the quad in the serial order k = 0, 1, 2, 3
// (sum4_ordered), and everything that depends only on those sums is computed redundantly by the four
// lanes.  Every floating-point operation and its operands are those of vs_svd4 / vs_inv4_from_svd /
// vs_condition_and_invert: the results are bit-identical.  All 32 lanes of the calling warp must
// execute it (lanes 4.. mirror a quad of their own).
__device__ __forceinline__ double vs_sum4_ordered(double t)
{
    const double t0 = __shfl_sync(0xffffffffu, t, 0, 4), t1 = __shfl_sync(0xffffffffu, t, 1, 4),
                 t2 = __shfl_sync(0xffffffffu, t, 2, 4), t3 = __shfl_sync(0xffffffffu, t, 3, 4);
    return ((t0 + t1) + t2) + t3;   // 0 + t0 == t0
}

// hrow: row k of H on lane k.  w: singular values (all lanes).  ucol[i] = u[k][i], vcol[i] = vt[i][k] on lane k.
__device__ __forceinline__ void vs_svd4_quad(const double* hrow, double* w, double* ucol, double* vcol)
{
    const int k = threadIdx.x & 3;
    const double eps = 2.220446049250313e-16 * 10;
    const double tiny = 2.2250738585072014e-308;
    double a[4], v[4], W[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = hrow[i]; v[i] = (i == k) ? 1.0 : 0.0; }
#pragma unroll
    for (int i = 0; i < 4; i++) W[i] = vs_sum4_ordered(a[i] * a[i]);
    for (int sweep = 0; sweep < 30; sweep++) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i + 1; j < 4; j++) {
                double aa = W[i], bb = W[j];
                double p = vs_sum4_ordered(a[i] * a[j]);
                if (fabs(p) <= eps * sqrt(aa * bb)) continue;
                p *= 2;
                double beta = aa - bb, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                const double t0 = c * a[i] + s * a[j];
                const double t1 = -s * a[i] + c * a[j];
                a[i] = t0; a[j] = t1;
                W[i] = vs_sum4_ordered(t0 * t0); W[j] = vs_sum4_ordered(t1 * t1);
                changed = true;
                const double v0 = c * v[i] + s * v[j];
                const double v1 = -s * v[i] + c * v[j];
                v[i] = v0; v[j] = v1;
            }
        if (!changed) break;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) W[i] = sqrt(vs_sum4_ordered(a[i] * a[i]));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int j = i;
        double wj = W[i];
#pragma unroll
        for (int q = i + 1; q < 4; q++)
            if (wj < W[q]) { wj = W[q]; j = q; }
#pragma unroll
        for (int q = i + 1; q < 4; q++)
            if (j == q) {
                double t = W[i]; W[i] = W[q]; W[q] = t;
                t = a[i]; a[i] = a[q]; a[q] = t;
                t = v[i]; v[i] = v[q]; v[q] = t;
            }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[i] = W[i];
        const double s = W[i] > tiny ? 1 / W[i] : 0.;
        ucol[i] = a[i] * s;
        vcol[i] = v[i];
    }
}

// hrow: row k of H on lane k (modified when regularised).  hinv_row: row k of the inverse on lane k.
__device__ __forceinline__ double vs_condition_and_invert_quad(double* hrow, double* hinv_row)
{
    const int k = threadIdx.x & 3;
    double w[4], u[4], v[4];
    vs_svd4_quad(hrow, w, u, v);
    const double cond = w[0] / (w[3] + 1e-10);
    if (cond > 1e6) {
        const double lambda = 1e-6 * w[0];
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (d == k) hrow[d] += lambda;
        vs_svd4_quad(hrow, w, u, v);
    }
    double threshold = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) threshold += w[i];
    threshold *= 2.220446049250313e-16 * 2;
#pragma unroll
    for (int c = 0; c < 4; c++) hinv_row[c] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool keep = !(fabs(w[i]) <= threshold);   // uniform over the quad
        const double wi = 1 / w[i];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double uc = __shfl_sync(0xffffffffu, u[i], c, 4);   // u[c][i]
            if (keep) hinv_row[c] += v[i] * (uc * wi);               // vt[i][k] * (u[c][i] / w[i])
        }
    }
    return cond;
}
#endif  // __CUDACC__
