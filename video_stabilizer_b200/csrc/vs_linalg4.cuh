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
