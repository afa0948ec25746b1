// smoother.hpp — sequential L1 trajectory smoother; stays host C++ by design (north_star):
// it is a 100-iteration relaxation over a 16-sample window, microseconds per frame.
// Public interface as the reference's smoother.hpp:10-29.
#pragma once

#include <vector>

#include "imgproc.hpp"

class L1SmootherCenter {
public:
    L1SmootherCenter(int lagBehind, int lagAhead, double lambda = 1.0);

    // Append the measurement of the newest frame.  Returns true and writes the smoothed
    // value of the oldest not-yet-finalised measurement once `lagAhead` later measurements
    // exist; returns false (outFinalized untouched) before that.
    bool update(const SimilarityTransform& meas, SimilarityTransform& outFinalized);

private:
    int m_lagBehind, m_lagAhead;
    double m_lambda;
    int m_nextToFinalize;
    std::vector<SimilarityTransform> m_measurements;   // grows for the life of the stream, as upstream
};

namespace vstab {
// The relaxation itself (reference smoother.cpp:18-64), exposed for tests: `iterations`
// sweeps of (a) a half step of every sample back towards its datum and (b) a sequential
// pass over neighbouring pairs that shrinks |x[i+1]-x[i]| by lambda or merges the pair.
void tvl1_relax(const double* data, int n, double lambda, int iterations, double* x);
// The value L1SmootherCenter::update finalises for measurement s (smoother.cpp:84-125 upstream): the relaxation of the raw
// window [max(0, s - lagBehind), s + lagAhead], read from `meas` (measurement 0 first).  A pure function of the raw
// measurements: values of different s are independent of each other.
SimilarityTransform smoother_finalize(const SimilarityTransform* meas, long s, int lagBehind, int lagAhead, double lambda);
}  // namespace vstab
