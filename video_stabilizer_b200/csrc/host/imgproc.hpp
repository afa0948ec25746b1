// imgproc.hpp — operator API of the alignment-and-warp path, drop-in for the reference's
// imgproc.hpp (same names, argument order, ownership and bool/throw error behaviour;
// reference imgproc.hpp:8-97).  Every operator runs on the GPU through the C ABI in
// include/vstab.h (VS_MEM_HOST: host buffers in, host buffers out); there is no CPU path.
//
// <Halide.h> and <opencv2/opencv.hpp> resolve to the real libraries when they are installed
// and to include/compat otherwise (the build adds include/compat last on the include path).
#pragma once

#include <Halide.h>
#include <opencv2/opencv.hpp>

#include <string>

// Shorthand for the buffer types of the signatures below (aliases: the functions take exactly
// Halide::Runtime::Buffer<T>&, as the reference's do).
namespace vstab {
using BufU8 = Halide::Runtime::Buffer<uint8_t>;
using BufU16 = Halide::Runtime::Buffer<uint16_t>;
using BufF32 = Halide::Runtime::Buffer<float>;
using BufF64 = Halide::Runtime::Buffer<double>;
}  // namespace vstab

// ---- keyframe feature operators -------------------------------------------------------
// sparse_jac (reference imgproc.hpp:8-14, imgproc.cpp:26-44): outputs are (re)allocated to
// (tw, th, 4) when their shape does not match local_max_x.
bool SparseJacobian(vstab::BufF32& grad_x, vstab::BufF32& grad_y,
                    vstab::BufU16& local_max_x, vstab::BufU16& local_max_y,
                    vstab::BufF32& output_x, vstab::BufF32& output_y);

// pyr_down (imgproc.hpp:16-18): caller allocates `output`; its extent is the work domain.
bool PyrDown(vstab::BufU8& input, vstab::BufU8& output);

// grad_xy (imgproc.hpp:20-23): caller allocates both outputs.
bool GradXY(vstab::BufU8& input, vstab::BufF32& output_x,
            vstab::BufF32& output_y);

// grad_argmax (imgproc.hpp:25-32): picks tile_size (largest even value in [2,20] that keeps
// at least 1000 tiles), (re)allocates the outputs to (w/tile, h/tile, 2) and writes, per tile,
// the pixel position of the first maximum of |grad|.
bool GradArgMax(vstab::BufF32& grad_x, vstab::BufF32& grad_y, int& tile_size,
                vstab::BufU16& local_max_x, vstab::BufU16& local_max_y);

// ---- transform algebra (imgproc.hpp:34-65) ----------------------------------------------
struct Point {
    double x = 0.0, y = 0.0;
    double distance(const Point& p) const;
};

// W(x,y) = ((1+A) x - B y + TX,  B x + (1+A) y + TY); all zero is the identity.  Pixels,
// +x right, +y down.  TX,TY are centre-based wherever a frame size is involved.
struct SimilarityTransform {
    double A = 0.0, B = 0.0, TX = 0.0, TY = 0.0;

    std::string toString() const;
    SimilarityTransform inverse() const;
    Point warp(Point p) const;                        // about the origin
    Point warp(Point p, double cx, double cy) const;  // about (cx, cy)
    double maxCornerDisplacement(double width, double height) const;
    // result(p) = w2(this(p)): this transform is applied first
    SimilarityTransform compose(const SimilarityTransform& w2) const;
};

// image_warp (imgproc.hpp:67-70): out(x,y) = bilinear(input, W(x,y)), repeat-edge, f32 output
// allocated by the caller.
bool ImageWarp(vstab::BufU8& input, const SimilarityTransform& transform,
               vstab::BufF32& output);

// ---- cv::Mat <-> Buffer converters (imgproc.hpp:72-76); throw std::runtime_error on misuse
vstab::BufU8 mat_to_halide_buffer_u8(const cv::Mat& mat);
vstab::BufU8 bgr_mat_to_halide_buffer_u8(const cv::Mat& mat);
cv::Mat halide_buffer_to_mat(const vstab::BufU8& buffer);
cv::Mat halide_buffer_to_mat(const vstab::BufF32& buffer);
cv::Mat halide_vec4_to_mat(const vstab::BufF64& vec4);

// ---- sparse solver operators ------------------------------------------------------------
// sparse_ica (imgproc.hpp:78-88): output(4) = 0.5 * sum_i J_i (template(p_i) - keyframe(W(p_i)))
bool SparseICA(vstab::BufU8& input_template, vstab::BufU8& input_keyframe,
               vstab::BufU16& selected_pixels_x, vstab::BufU16& selected_pixels_y,
               vstab::BufF32& selected_jacobians_x, vstab::BufF32& selected_jacobians_y,
               const SimilarityTransform& transform, vstab::BufF64& output);

// sparse_warpdiff (imgproc.hpp:90-95): output(tw,th) = trunc |keyframe(W(p)) - template(p)|
bool SparseWarpDiff(vstab::BufU8& input_template, vstab::BufU8& input_keyframe,
                    vstab::BufU16& local_max, const SimilarityTransform& transform,
                    vstab::BufU16& output);

// imgproc.hpp:97 — the BGR warp of the stabilizer: bit-exact with
// cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT) of the forward matrix built from `transform`.
cv::Mat warpBySimilarityTransform(const cv::Mat& src, const SimilarityTransform& transform);

// ---- additions of this implementation (not in the reference) ------------------------------
struct vs_ctx;
namespace vstab {
// The calling thread's GPU context (created on first use on device VSTAB_DEVICE, default 0).
// Throws std::runtime_error when libvstab.so finds no CUDA device: there is no CPU fallback.
vs_ctx* thread_context();
// Bind the calling thread's operators to another device (drops the previous context).
void set_thread_device(int device);
}  // namespace vstab
