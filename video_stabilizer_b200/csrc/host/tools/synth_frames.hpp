// synth_frames.hpp — procedural test footage for align_test / video_test (the reference's
// fixtures input.png, template.png and recordings/*.mp4 are not part of its repository).
// A multi-octave value-noise texture on an oversized canvas; frame t is the canvas seen through
// a small random-walk similarity, rendered with warpBySimilarityTransform itself.
#pragma once

#include <math.h>
#include <stdint.h>

#include <random>
#include <vector>

#include "imgproc.hpp"

namespace synth {

inline uint32_t hash2(uint32_t x, uint32_t y, uint32_t seed)
{
    uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (seed * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

inline float value_noise(float x, float y, uint32_t seed)
{
    const int xi = (int)floorf(x), yi = (int)floorf(y);
    const float fx = x - xi, fy = y - yi;
    const float sx = fx * fx * (3 - 2 * fx), sy = fy * fy * (3 - 2 * fy);
    auto v = [&](int i, int j) { return (hash2((uint32_t)(xi + i), (uint32_t)(yi + j), seed) & 0xffff) / 65535.0f; };
    const float a = v(0, 0) + (v(1, 0) - v(0, 0)) * sx, b = v(0, 1) + (v(1, 1) - v(0, 1)) * sx;
    return a + (b - a) * sy;
}

// BGR canvas (gray replicated) with detail at 3, 9 and 30 pixel scales
inline cv::Mat make_canvas(int width, int height, uint32_t seed)
{
    cv::Mat m(height, width, CV_8UC3);
    for (int y = 0; y < height; y++) {
        uint8_t* row = m.ptr<uint8_t>(y);
        for (int x = 0; x < width; x++) {
            float v = 0.2f * value_noise(x / 3.0f, y / 3.0f, seed) + 0.3f * value_noise(x / 9.0f, y / 9.0f, seed + 1) +
                      0.5f * value_noise(x / 30.0f, y / 30.0f, seed + 2);
            int g = (int)(v * 255.0f + 0.5f);
            g = g < 0 ? 0 : (g > 255 ? 255 : g);
            row[3 * x] = row[3 * x + 1] = row[3 * x + 2] = (uint8_t)g;
        }
    }
    return m;
}

inline cv::Mat to_gray(const cv::Mat& bgr)
{
    cv::Mat g(bgr.rows, bgr.cols, CV_8UC1);
    for (int y = 0; y < bgr.rows; y++)
        for (int x = 0; x < bgr.cols; x++) g.ptr<uint8_t>(y)[x] = bgr.ptr<uint8_t>(y)[3 * x + 1];
    return g;
}

struct Jitter {
    std::mt19937 rng;
    double tx = 0, ty = 0;
    explicit Jitter(uint32_t seed) : rng(seed) {}
    SimilarityTransform next()
    {
        std::uniform_real_distribution<double> step(-2.0, 2.0), ab(-0.001, 0.001);
        tx = std::max(-16.0, std::min(16.0, tx + step(rng)));
        ty = std::max(-16.0, std::min(16.0, ty + step(rng)));
        SimilarityTransform t;
        t.A = ab(rng); t.B = ab(rng); t.TX = tx; t.TY = ty;
        return t;
    }
};

// the view of `canvas` under camera pose `pose`: warpBySimilarityTransform pushes content by the
// transform, so the frame shows canvas content displaced by pose
inline cv::Mat render(const cv::Mat& canvas, const SimilarityTransform& pose) { return warpBySimilarityTransform(canvas, pose); }

}  // namespace synth
