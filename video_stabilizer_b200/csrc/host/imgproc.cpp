// imgproc.cpp — host side of the operator API: marshals Halide::Runtime::Buffer / cv::Mat
// arguments into vs_img descriptors and calls the C ABI (include/vstab.h).  Allocation rules,
// centre->upper-left parameter conversions and error behaviour follow the reference wrappers
// (reference imgproc.cpp:26-202, :204-325, :327-437, :446-484).
#include "imgproc.hpp"

#include <math.h>
#include <stdlib.h>

#include <sstream>
#include <stdexcept>

#include "vstab.h"

namespace vstab {

namespace {
struct ThreadCtx {
    vs_ctx* ctx = nullptr;
    int device = -1;
    ~ThreadCtx() { if (ctx) vs_ctx_destroy(ctx); }
};
thread_local ThreadCtx t_ctx;

int default_device()
{
    const char* e = getenv("VSTAB_DEVICE");
    return e ? atoi(e) : 0;
}
}  // namespace

vs_ctx* thread_context()
{
    if (!t_ctx.ctx) {
        int dev = t_ctx.device >= 0 ? t_ctx.device : default_device();
        if (vs_ctx_create(dev, &t_ctx.ctx) != VS_OK)
            throw std::runtime_error(std::string("vstab: cannot create a GPU context: ") + vs_last_error(nullptr));
        t_ctx.device = dev;
    }
    return t_ctx.ctx;
}

void set_thread_device(int device)
{
    if (t_ctx.ctx && t_ctx.device == device) return;
    if (t_ctx.ctx) { vs_ctx_destroy(t_ctx.ctx); t_ctx.ctx = nullptr; }
    t_ctx.device = device;
}

}  // namespace vstab

namespace {

template <typename T>
vs_img img2d(const Halide::Runtime::Buffer<T>& b)
{
    if (b.dimensions() < 2) throw std::runtime_error("vstab: expected a 2-D buffer");
    if (b.dim(0).stride() != 1) throw std::runtime_error("vstab: buffers must be dense along x");
    vs_img d;
    d.data = (void*)b.data();
    d.width = b.dim(0).extent();
    d.height = b.dim(1).extent();
    d.stride = b.dim(1).stride();
    d.batch = 1;
    d.batch_stride = 0;
    return d;
}

// planar (w,h,c) buffers are handed to the C ABI as one dense block
template <typename T>
bool dense_planar(const Halide::Runtime::Buffer<T>& b, int dims)
{
    if (b.dimensions() != dims) return false;
    int stride = 1;
    for (int i = 0; i < dims; i++) {
        if (b.dim(i).stride() != stride) return false;
        stride *= b.dim(i).extent();
    }
    return true;
}

bool report(int rc, const char* what)
{
    if (rc == VS_OK) return true;
    std::cerr << what << " failed: " << vs_last_error(vstab::thread_context()) << std::endl;
    return false;
}

// reference imgproc.cpp:69-75 / :98-103: (width*0.5f) is an f32 product promoted to f64
void ul_params_half(const SimilarityTransform& t, int w, int h, float out[4])
{
    out[0] = static_cast<float>(t.A);
    out[1] = static_cast<float>(t.B);
    out[2] = static_cast<float>(t.TX - t.A * (w * 0.5f) + t.B * (h * 0.5f));
    out[3] = static_cast<float>(t.TY - t.B * (w * 0.5f) - t.A * (h * 0.5f));
}

}  // namespace

bool SparseJacobian(Halide::Runtime::Buffer<float>& grad_x, Halide::Runtime::Buffer<float>& grad_y,
                    Halide::Runtime::Buffer<uint16_t>& local_max_x, Halide::Runtime::Buffer<uint16_t>& local_max_y,
                    Halide::Runtime::Buffer<float>& output_x, Halide::Runtime::Buffer<float>& output_y)
{
    const int tw = local_max_x.dim(0).extent(), th = local_max_x.dim(1).extent();
    if (output_x.dimensions() != 3 || output_x.dim(0).extent() != tw || output_x.dim(1).extent() != th) {
        output_x = Halide::Runtime::Buffer<float>(tw, th, 4);
        output_y = Halide::Runtime::Buffer<float>(tw, th, 4);
    }
    if (!dense_planar(local_max_x, 3) || !dense_planar(local_max_y, 3) || !dense_planar(output_x, 3) || !dense_planar(output_y, 3)) {
        std::cerr << "SparseJacobian: keypoint / Jacobian buffers must be dense planar" << std::endl;
        return false;
    }
    vs_img gx = img2d(grad_x), gy = img2d(grad_y);
    return report(vs_sparse_jac_f32(vstab::thread_context(), &gx, &gy, local_max_x.data(), local_max_y.data(), tw, th,
                                    output_x.data(), output_y.data(), VS_MEM_HOST), "SparseJacobian");
}

bool SparseICA(Halide::Runtime::Buffer<uint8_t>& input_template, Halide::Runtime::Buffer<uint8_t>& input_keyframe,
               Halide::Runtime::Buffer<uint16_t>& selected_pixels_x, Halide::Runtime::Buffer<uint16_t>& selected_pixels_y,
               Halide::Runtime::Buffer<float>& selected_jacobians_x, Halide::Runtime::Buffer<float>& selected_jacobians_y,
               const SimilarityTransform& transform, Halide::Runtime::Buffer<double>& output)
{
    if (output.dimensions() != 1 || output.dim(0).extent() != 4) output = Halide::Runtime::Buffer<double>(4);
    if (!dense_planar(selected_pixels_x, 2) || !dense_planar(selected_pixels_y, 2) ||
        !dense_planar(selected_jacobians_x, 2) || !dense_planar(selected_jacobians_y, 2)) {
        std::cerr << "SparseICA: selection buffers must be dense planar" << std::endl;
        return false;
    }
    float P[4];
    ul_params_half(transform, input_template.width(), input_template.height(), P);
    vs_img t = img2d(input_template), k = img2d(input_keyframe);
    return report(vs_sparse_ica_f64(vstab::thread_context(), &t, &k, selected_pixels_x.data(), selected_pixels_x.dim(0).extent(),
                                    selected_pixels_y.data(), selected_pixels_y.dim(0).extent(), selected_jacobians_x.data(),
                                    selected_jacobians_y.data(), P[0], P[1], P[2], P[3], output.data(), VS_MEM_HOST), "SparseICA");
}

bool SparseWarpDiff(Halide::Runtime::Buffer<uint8_t>& input_template, Halide::Runtime::Buffer<uint8_t>& input_keyframe,
                    Halide::Runtime::Buffer<uint16_t>& local_max, const SimilarityTransform& transform,
                    Halide::Runtime::Buffer<uint16_t>& output)
{
    const int tw = local_max.dim(0).extent(), th = local_max.dim(1).extent();
    if (output.dimensions() != 2 || output.dim(0).extent() != tw || output.dim(1).extent() != th)
        output = Halide::Runtime::Buffer<uint16_t>(tw, th);
    if (!dense_planar(local_max, 3) || !dense_planar(output, 2)) {
        std::cerr << "SparseWarpDiff: keypoint / output buffers must be dense planar" << std::endl;
        return false;
    }
    float P[4];
    ul_params_half(transform, input_template.width(), input_template.height(), P);
    vs_img t = img2d(input_template), k = img2d(input_keyframe);
    return report(vs_sparse_warpdiff_u8_u16(vstab::thread_context(), &t, &k, local_max.data(), tw, th, P[0], P[1], P[2], P[3],
                                            output.data(), VS_MEM_HOST), "SparseWarpDiff");
}

bool PyrDown(Halide::Runtime::Buffer<uint8_t>& input, Halide::Runtime::Buffer<uint8_t>& output)
{
    vs_img in = img2d(input), out = img2d(output);
    return report(vs_pyr_down_u8(vstab::thread_context(), &in, &out, VS_MEM_HOST), "PyrDown");
}

bool ImageWarp(Halide::Runtime::Buffer<uint8_t>& input, const SimilarityTransform& transform,
               Halide::Runtime::Buffer<float>& output)
{
    // reference imgproc.cpp:125-131: this operator's centre is ((w-1)/2, (h-1)/2) and the
    // double parameters narrow to f32 at the pipeline boundary
    const double cx = (input.width() - 1) * 0.5, cy = (input.height() - 1) * 0.5;
    const float P[4] = {(float)transform.A, (float)transform.B,
                        (float)(transform.TX - transform.A * cx + transform.B * cy),
                        (float)(transform.TY - transform.B * cx - transform.A * cy)};
    vs_img in = img2d(input), out = img2d(output);
    return report(vs_image_warp_u8_f32(vstab::thread_context(), &in, P, &out, VS_MEM_HOST), "ImageWarp");
}

bool GradXY(Halide::Runtime::Buffer<uint8_t>& input, Halide::Runtime::Buffer<float>& output_x,
            Halide::Runtime::Buffer<float>& output_y)
{
    vs_img in = img2d(input), gx = img2d(output_x), gy = img2d(output_y);
    return report(vs_grad_xy_u8_f32(vstab::thread_context(), &in, &gx, &gy, VS_MEM_HOST), "GradXY");
}

bool GradArgMax(Halide::Runtime::Buffer<float>& grad_x, Halide::Runtime::Buffer<float>& grad_y, int& tile_size,
                Halide::Runtime::Buffer<uint16_t>& local_max_x, Halide::Runtime::Buffer<uint16_t>& local_max_y)
{
    tile_size = vs_grad_argmax_tile_size(grad_x.width(), grad_y.height());
    const int tw = grad_x.width() / tile_size, th = grad_y.height() / tile_size;
    if (local_max_x.dimensions() != 3 || local_max_x.dim(0).extent() != tw || local_max_x.dim(1).extent() != th) {
        local_max_x = Halide::Runtime::Buffer<uint16_t>(tw, th, 2);
        local_max_y = Halide::Runtime::Buffer<uint16_t>(tw, th, 2);
    }
    if (!dense_planar(local_max_x, 3) || !dense_planar(local_max_y, 3)) {
        std::cerr << "GradArgMax: output buffers must be dense planar" << std::endl;
        return false;
    }
    vs_img gx = img2d(grad_x), gy = img2d(grad_y);
    return report(vs_grad_argmax_f32_u16(vstab::thread_context(), &gx, &gy, tile_size, local_max_x.data(), local_max_y.data(),
                                         VS_MEM_HOST), "GradArgMax");
}

// ------------------------------------------------------------------ converters
Halide::Runtime::Buffer<uint8_t> mat_to_halide_buffer_u8(const cv::Mat& mat)
{
    if (mat.type() != CV_8UC1) throw std::runtime_error("Input cv::Mat must be an 8-bit single-channel (grayscale) image.");
    // A non-continuous Mat is wrapped with its row step (the C ABI takes strides), so the
    // returned buffer always aliases `mat` — the reference's version dangles in that case
    // (reference imgproc.cpp:212-219).
    halide_dimension_t shape[2] = {
        halide_dimension_t(0, mat.cols, 1),
        halide_dimension_t(0, mat.rows, (int32_t)mat.step1(0)),
    };
    return Halide::Runtime::Buffer<uint8_t>(mat.data, 2, shape);
}

Halide::Runtime::Buffer<uint8_t> bgr_mat_to_halide_buffer_u8(const cv::Mat& mat)
{
    if (mat.type() != CV_8UC3) throw std::runtime_error("Input cv::Mat must be an 8-bit 3-channel (BGR) image.");
    // interleaved view (x stride 3, channel stride 1); the reference's strides here are
    // wrong and the function is dead code upstream (reference imgproc.cpp:236-267)
    halide_dimension_t shape[3] = {
        halide_dimension_t(0, mat.cols, 3),
        halide_dimension_t(0, mat.rows, (int32_t)mat.step1(0)),
        halide_dimension_t(0, 3, 1),
    };
    return Halide::Runtime::Buffer<uint8_t>(mat.data, 3, shape);
}

cv::Mat halide_buffer_to_mat(const Halide::Runtime::Buffer<uint8_t>& buffer)
{
    if (buffer.dimensions() != 2) throw std::runtime_error("Only 2-dimensional Halide buffers can be converted to cv::Mat.");
    return cv::Mat(buffer.height(), buffer.width(), CV_8UC1, (void*)buffer.data(), (size_t)buffer.stride(1) * sizeof(uint8_t));
}

cv::Mat halide_buffer_to_mat(const Halide::Runtime::Buffer<float>& buffer)
{
    if (buffer.dimensions() != 2) throw std::runtime_error("Only 2-dimensional Halide buffers can be converted to cv::Mat.");
    return cv::Mat(buffer.height(), buffer.width(), CV_32FC1, (void*)buffer.data(), (size_t)buffer.stride(1) * sizeof(float));
}

cv::Mat halide_vec4_to_mat(const Halide::Runtime::Buffer<double>& vec4)
{
    if (vec4.dimensions() != 1 || vec4.width() != 4) throw std::runtime_error("Expected a 1D Halide buffer of length 4");
    cv::Mat v(4, 1, CV_64F);
    for (int i = 0; i < 4; i++) v.at<double>(i, 0) = vec4(i);
    return v;
}

// ------------------------------------------------------------------ transform algebra
std::string SimilarityTransform::toString() const
{
    std::stringstream ss;
    ss << "A=" << A << ", B=" << B << ", TX=" << TX << ", TY=" << TY;
    return ss.str();
}

// With p = 1+A, q = B the linear part is [[p,-q],[q,p]]; its inverse is the transpose
// divided by p^2+q^2 (reference imgproc.cpp:333-359).
SimilarityTransform SimilarityTransform::inverse() const
{
    const double p = 1.0 + A, q = B;
    const double denom = p * p + q * q;
    SimilarityTransform r;
    r.A = (p / denom) - 1.0;
    r.B = -q / denom;
    r.TX = (-p * TX - q * TY) / denom;
    r.TY = (q * TX - p * TY) / denom;
    return r;
}

// reference imgproc.cpp:361-387
SimilarityTransform SimilarityTransform::compose(const SimilarityTransform& w2) const
{
    const double p1 = 1.0 + A, q1 = B, p2 = 1.0 + w2.A, q2 = w2.B;
    SimilarityTransform r;
    r.A = (p2 * p1 - q2 * q1) - 1.0;
    r.B = (p2 * q1 + q2 * p1);
    r.TX = p2 * TX - q2 * TY + w2.TX;
    r.TY = q2 * TX + p2 * TY + w2.TY;
    return r;
}

Point SimilarityTransform::warp(Point p) const
{
    Point r;
    r.x = (1 + A) * p.x - B * p.y + TX;
    r.y = B * p.x + (1 + A) * p.y + TY;
    return r;
}

Point SimilarityTransform::warp(Point p, double cx, double cy) const
{
    const double x = p.x - cx, y = p.y - cy;
    Point r;
    r.x = (1 + A) * x - B * y + cx + TX;
    r.y = B * x + (1 + A) * y + cy + TY;
    return r;
}

double Point::distance(const Point& p) const
{
    const double dx = x - p.x, dy = y - p.y;
    return std::sqrt(dx * dx + dy * dy);
}

// reference imgproc.cpp:419-437: corners (0,0) (w,0) (0,h) (w,h) about (w/2, h/2)
double SimilarityTransform::maxCornerDisplacement(double width, double height) const
{
    const double cx = width * 0.5, cy = height * 0.5;
    const Point corners[4] = {Point{0.0, 0.0}, Point{width, 0.0}, Point{0.0, height}, Point{width, height}};
    double worst = 0.0;
    for (const Point& c : corners) worst = std::max(worst, warp(c, cx, cy).distance(c));
    return worst;
}

// ------------------------------------------------------------------ BGR warp
cv::Mat warpBySimilarityTransform(const cv::Mat& src, const SimilarityTransform& transform)
{
    if (src.type() != CV_8UC3) throw std::runtime_error("warpBySimilarityTransform: the GPU path takes CV_8UC3 frames");
    // forward matrix of reference imgproc.cpp:458-469 (centre ((cols-1)/2, (rows-1)/2))
    const double cx = (src.cols - 1) * 0.5, cy = (src.rows - 1) * 0.5;
    const double M[6] = {1.0 + transform.A, -transform.B, transform.TX - transform.A * cx + transform.B * cy,
                         transform.B, 1.0 + transform.A, transform.TY - transform.B * cx - transform.A * cy};
    cv::Mat dst(src.rows, src.cols, CV_8UC3);
    vs_img s{src.data, src.cols, src.rows, (int64_t)src.step[0], 1, 0};
    vs_img d{dst.data, dst.cols, dst.rows, (int64_t)dst.step[0], 1, 0};
    vs_ctx* ctx = vstab::thread_context();
    if (vs_bgr_warp_u8(ctx, &s, M, &d, 0, 0, VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0, VS_MEM_HOST) != VS_OK)
        throw std::runtime_error(std::string("warpBySimilarityTransform: ") + vs_last_error(ctx));
    return dst;
}
