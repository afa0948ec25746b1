// smoother.cpp — see smoother.hpp.  Arithmetic order follows reference smoother.cpp:18-127
// exactly (the pairwise pass is in-place and order dependent), so trajectories match the
// reference bit for bit.
#include "smoother.hpp"

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

}  // namespace vstab

L1SmootherCenter::L1SmootherCenter(int lagBehind, int lagAhead, double lambda)
    : m_lagBehind(lagBehind), m_lagAhead(lagAhead), m_lambda(lambda), m_nextToFinalize(0)
{
}

bool L1SmootherCenter::update(const SimilarityTransform& meas, SimilarityTransform& outFinalized)
{
    m_measurements.push_back(meas);
    const int newest = (int)m_measurements.size() - 1;
    if (m_nextToFinalize + m_lagAhead > newest) return false;

    const int first = std::max(0, m_nextToFinalize - m_lagBehind);
    const int last = m_nextToFinalize + m_lagAhead;
    const int n = last - first + 1;
    // one window per parameter, smoothed independently
    std::vector<double> window(4 * (size_t)n), smooth(4 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const SimilarityTransform& m = m_measurements[first + i];
        window[0 * n + i] = m.A;
        window[1 * n + i] = m.B;
        window[2 * n + i] = m.TX;
        window[3 * n + i] = m.TY;
    }
    for (int c = 0; c < 4; c++) vstab::tvl1_relax(&window[(size_t)c * n], n, m_lambda, 100, &smooth[(size_t)c * n]);
    const int mid = m_nextToFinalize - first;
    outFinalized.A = smooth[0 * n + mid];
    outFinalized.B = smooth[1 * n + mid];
    outFinalized.TX = smooth[2 * n + mid];
    outFinalized.TY = smooth[3 * n + mid];
    m_nextToFinalize++;
    return true;
}
