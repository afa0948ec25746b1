// smoother.cpp — see smoother.hpp.  Arithmetic order follows reference smoother.cpp:18-127
// exactly (the pairwise pass is in-place and order dependent), so trajectories match the
// reference bit for bit.
#include "smoother.hpp"

#include <immintrin.h>
#include <math.h>

#include <algorithm>

namespace vstab {

void tvl1_relax(const double* data, int n, double lambda, int iterations, double* x)
{
    for (int i = 0; i < n; i++) x[i] = data[i];
    const double alpha = 0.5;
    for (int it = 0; it < iterations; it++) {
        for (int i = 0; i < n; i++) x[i] = (1.0 - alpha) * x[i] + alpha * data[i];
        for (int i = 0; i + 1 < n; i++) {
            const double diff = x[i + 1] - x[i];
            const double mag = fabs(diff);
            if (mag > lambda) {
                const double shrink = (mag - lambda) / mag * 0.5;
                x[i] += diff * shrink;
                x[i + 1] -= diff * shrink;
            } else {
                const double mid = 0.5 * (x[i] + x[i + 1]);
                x[i] = mid;
                x[i + 1] = mid;
            }
        }
    }
}

namespace {

// The four transform parameters are smoothed independently with identical control flow
// apart from the per-pair branch, so they run as the four lanes of one vector: every lane
// performs exactly the scalar operations above (both branch results are formed and the
// lane's own one is kept), which keeps the trajectory bit-identical to the scalar code.
typedef double v4d __attribute__((vector_size(32)));
typedef long long v4i __attribute__((vector_size(32)));

inline v4d v4_abs(v4d a)
{
    v4i bits = (v4i)a & (v4i){0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
    return (v4d)bits;
}

inline v4d v4_select(v4i mask, v4d a, v4d b) { return (v4d)(((v4i)a & mask) | ((v4i)b & ~mask)); }

// data, x: n vectors (A, B, TX, TY).
// The scalar loop nest is a chain of iterations x (n-1) dependent pair updates, each with a
// divide: pure latency.  Pair i of iteration t only needs pair i-1 of iteration t and pair
// i+1 of iteration t-1, so all pairs with the same 2t+i are independent; walking those
// wavefronts in place performs exactly the same operations on the same values in a
// dependency-respecting order (bit-identical result) with up to (n-1)/2 divides in flight.
void tvl1_relax4(const v4d* data, int n, double lambda, int iterations, v4d* x)
{
    const v4d half = {0.5, 0.5, 0.5, 0.5};
    const v4d lam = {lambda, lambda, lambda, lambda};
    for (int i = 0; i < n; i++) x[i] = data[i];
    if (n < 2 || iterations <= 0) return;   // a lone sample relaxes onto itself: 0.5 d + 0.5 d == d
    const int pairs = n - 1;
    const int waves = 2 * (iterations - 1) + pairs;
    for (int w = 0; w < waves; w++) {
        // iterations t with 0 <= w - 2t <= pairs - 1
        int t_lo = (w - (pairs - 1) + 1) / 2;
        if (t_lo < 0) t_lo = 0;
        int t_hi = w / 2;
        if (t_hi > iterations - 1) t_hi = iterations - 1;
        for (int t = t_lo; t <= t_hi; t++) {
            const int i = w - 2 * t;
            // the relaxation step of iteration t, applied where each value is first used:
            // x[i+1] always, x[i] only for the first pair (later pairs inherit it from pair i-1)
            const v4d a = i == 0 ? half * x[0] + half * data[0] : x[i];
            const v4d b = half * x[i + 1] + half * data[i + 1];
            const v4d diff = b - a;
            const v4d mag = v4_abs(diff);
            const v4i shrink_lane = mag > lam;
            const v4d mid = half * (a + b);
            if (!_mm256_movemask_pd((__m256d)shrink_lane)) {
                // no lane exceeds lambda (the common case for sub-lambda jitter): skip the divide
                x[i] = mid;
                x[i + 1] = mid;
                continue;
            }
            const v4d shrink = (mag - lam) / mag * half;
            const v4d step = diff * shrink;
            x[i] = v4_select(shrink_lane, a + step, mid);
            x[i + 1] = v4_select(shrink_lane, b - step, mid);
        }
    }
}

}  // namespace

}  // namespace vstab

L1SmootherCenter::L1SmootherCenter(int lagBehind, int lagAhead, double lambda)
    : m_lagBehind(lagBehind), m_lagAhead(lagAhead), m_lambda(lambda), m_nextToFinalize(0)
{
}

namespace vstab {

SimilarityTransform smoother_finalize(const SimilarityTransform* meas, long s, int lagBehind, int lagAhead, double lambda)
{
    const long first = std::max(0L, s - lagBehind);
    const long last = s + lagAhead;
    const int n = (int)(last - first + 1);
    // one window, the four parameters as the four lanes of a vector
    v4d stack_window[64], stack_smooth[64];
    std::vector<v4d> heap;
    v4d* window = stack_window;
    v4d* smooth = stack_smooth;
    if (n > 64) {
        heap.resize((size_t)2 * n);
        window = heap.data();
        smooth = heap.data() + n;
    }
    for (int i = 0; i < n; i++) {
        const SimilarityTransform& m = meas[first + i];
        window[i] = (v4d){m.A, m.B, m.TX, m.TY};
    }
    tvl1_relax4(window, n, lambda, 100, smooth);
    const int mid = (int)(s - first);
    SimilarityTransform out;
    out.A = smooth[mid][0];
    out.B = smooth[mid][1];
    out.TX = smooth[mid][2];
    out.TY = smooth[mid][3];
    return out;
}

}  // namespace vstab

bool L1SmootherCenter::update(const SimilarityTransform& meas, SimilarityTransform& outFinalized)
{
    m_measurements.push_back(meas);
    const int newest = (int)m_measurements.size() - 1;
    if (m_nextToFinalize + m_lagAhead > newest) return false;
    outFinalized = vstab::smoother_finalize(m_measurements.data(), m_nextToFinalize, m_lagBehind, m_lagAhead, m_lambda);
    m_nextToFinalize++;
    return true;
}
