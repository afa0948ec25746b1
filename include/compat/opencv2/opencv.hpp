// compat/opencv2/opencv.hpp — minimal cv:: surface for builds without OpenCV.
//
// The reference's public interfaces take and return cv::Mat (alignment.hpp:55-58,
// stabilizer.hpp:39, imgproc.hpp:72-76,97).  This header provides a reference-counted
// cv::Mat with the members those interfaces and their callers touch (rows, cols, data, step,
// type, clone, ROI, at<>, zeros, convertTo) plus the small value types.  Everything is
// header-only EXCEPT the functions at the bottom (cvtColor, warpAffine, SVD, Mat::inv,
// operator*, phaseCorrelate): those are only DECLARED here.  The product never calls them —
// its colour conversion, warps and 4x4 algebra run on the GPU — so libvstab_host.so has no
// definition for them; the CPU oracle defines them (oracle/ref_shim/cv_impl.cpp) so that the
// reference's own alignment.cpp / imgproc.cpp / stabilizer.cpp compile unmodified against
// this header into oracle/_ref.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <iostream>
#include <memory>
#include <random>      // the reference's align_test.cpp uses std::mt19937 with only OpenCV / Halide headers included
#include <stdexcept>
#include <string>
#include <vector>

#define CV_COMPAT_SHIM 1

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    double operator[](int i) const { return val[i]; }
};

enum { INTER_NEAREST = 0, INTER_LINEAR = 1, WARP_INVERSE_MAP = 16 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };
enum { COLOR_BGR2GRAY = 6 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { NORM_MINMAX = 32 };
enum { FILLED = -1 };
enum { CAP_PROP_FRAME_WIDTH = 3, CAP_PROP_FRAME_HEIGHT = 4, CAP_PROP_FPS = 5, CAP_PROP_FRAME_COUNT = 7 };
enum { VIDEOWRITER_PROP_QUALITY = 1 };

inline size_t depth_bytes(int depth)
{
    switch (depth) {
    case CV_8U: return 1;
    case CV_16U: return 2;
    case CV_32S: case CV_32F: return 4;
    case CV_64F: return 8;
    default: throw std::runtime_error("compat cv::Mat: unsupported depth");
    }
}

class Mat;
class MatExpr;

// cv::MatStep: converts to the row step in bytes, indexable
struct MatStep {
    size_t p[2];
    MatStep() { p[0] = p[1] = 0; }
    operator size_t() const { return p[0]; }
    size_t operator[](int i) const { return p[i]; }
    size_t& operator[](int i) { return p[i]; }
};

class Mat {
public:
    int flags = 0;   // holds the type
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    MatStep step;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    // non-owning wrap of external memory
    Mat(int r, int c, int type, void* ext, size_t step_bytes = 0)
    {
        flags = type; rows = r; cols = c; data = (uint8_t*)ext;
        step[1] = elemSize();
        step[0] = step_bytes ? step_bytes : (size_t)c * elemSize();
    }
    // wrap of external memory that the Mat (and its copies / ROIs) keeps alive through `owner` — what a custom
    // cv::MatAllocator (e.g. cv::cuda::HostMem::getAllocator() for page-locked memory) does with real OpenCV
    Mat(int r, int c, int type, void* ext, size_t step_bytes, std::shared_ptr<void> owner)
    {
        flags = type; rows = r; cols = c; data = (uint8_t*)ext;
        step[1] = elemSize();
        step[0] = step_bytes ? step_bytes : (size_t)c * elemSize();
        owner_ = std::move(owner);
    }

    void create(int r, int c, int type)
    {
        if (data && owner_ && rows == r && cols == c && flags == type && isContinuous()) return;
        flags = type; rows = r; cols = c;
        step[1] = elemSize();
        step[0] = (size_t)c * elemSize();
        auto buf = std::make_shared<std::vector<uint8_t>>((size_t)r * step[0] + 64);
        data = buf->data();
        owner_ = buf;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; }

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.data) memset(m.data, 0, (size_t)r * m.step[0]); return m; }
    static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
    static Mat eye(int r, int c, int type)
    {
        Mat m = zeros(r, c, type);
        for (int i = 0; i < std::min(r, c); i++) {
            if (type == CV_64F) m.at<double>(i, i) = 1.0;
            else if (type == CV_32F) m.at<float>(i, i) = 1.0f;
            else throw std::runtime_error("compat cv::Mat::eye: unsupported type");
        }
        return m;
    }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { return depth_bytes(depth()); }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    size_t step1(int i = 0) const { return step[i] / elemSize1(); }
    size_t total() const { return (size_t)rows * (size_t)cols; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step[0] == (size_t)cols * elemSize(); }

    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step[0]); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step[0]); }
    uint8_t* ptr(int r = 0) { return data + (size_t)r * step[0]; }
    const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step[0]; }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step[0]))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step[0]))[c]; }
    // single index: element i of a vector-shaped (or continuous) matrix
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : (cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols)); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : (cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols)); }

    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step[0], data + (size_t)r * step[0], rb);
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }

    // ROI view sharing the parent's storage
    Mat operator()(const Rect& roi) const
    {
        if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > cols || roi.y + roi.height > rows)
            throw std::runtime_error("compat cv::Mat: ROI out of range");
        Mat m;
        m.flags = flags; m.rows = roi.height; m.cols = roi.width;
        m.step = step;
        m.data = data + (size_t)roi.y * step[0] + (size_t)roi.x * elemSize();
        m.owner_ = owner_;
        return m;
    }

    Mat& setTo(const Scalar& s)
    {
        const int cn = channels();
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < cn; k++) {
                    const double v = s.val[k < 4 ? k : 3];
                    switch (depth()) {
                    case CV_8U: ptr<uint8_t>(r)[c * cn + k] = (uint8_t)v; break;
                    case CV_16U: ptr<uint16_t>(r)[c * cn + k] = (uint16_t)v; break;
                    case CV_32S: ptr<int32_t>(r)[c * cn + k] = (int32_t)v; break;
                    case CV_32F: ptr<float>(r)[c * cn + k] = (float)v; break;
                    case CV_64F: ptr<double>(r)[c * cn + k] = v; break;
                    }
                }
        return *this;
    }

    // u8/f32/f64 -> f32/f64/u8 (saturating round for u8), used for the phase-layer conversion
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const
    {
        const int cn = channels();
        Mat out(rows, cols, CV_MAKETYPE(CV_MAT_DEPTH(rtype), cn));
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols * cn; c++) {
                double v;
                switch (depth()) {
                case CV_8U: v = ptr<uint8_t>(r)[c]; break;
                case CV_16U: v = ptr<uint16_t>(r)[c]; break;
                case CV_32F: v = ptr<float>(r)[c]; break;
                case CV_64F: v = ptr<double>(r)[c]; break;
                default: throw std::runtime_error("compat cv::Mat::convertTo: unsupported source depth");
                }
                v = v * alpha + beta;
                switch (out.depth()) {
                case CV_8U: out.ptr<uint8_t>(r)[c] = (uint8_t)std::min(255.0, std::max(0.0, nearbyint(v))); break;
                case CV_32F: out.ptr<float>(r)[c] = (float)v; break;
                case CV_64F: out.ptr<double>(r)[c] = v; break;
                default: throw std::runtime_error("compat cv::Mat::convertTo: unsupported destination depth");
                }
            }
        dst = out;
    }

    // declared only (see header comment): defined by the oracle's shim, never by the product
    Mat inv(int method = DECOMP_LU) const;

    bool ownsData() const { return (bool)owner_; }
    const std::shared_ptr<void>& owner() const { return owner_; }

private:
    std::shared_ptr<void> owner_;
};

// cv::Mat_<T>(r,c) << a, b, c ... ; only what imgproc.cpp:467-469 needs
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m, T first);
    MatCommaInitializer_& operator,(T v);
    operator Mat() const;
    operator Mat_<T>() const;
private:
    Mat_<T>* m_;
    int idx_;
};

template <typename T> struct DepthOf;
template <> struct DepthOf<uint8_t> { enum { value = CV_8U }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) { if (data) memset(data, 0, (size_t)r * step[0]); }
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
    MatCommaInitializer_<T> operator<<(T v) { return MatCommaInitializer_<T>(this, v); }
};

template <typename T> MatCommaInitializer_<T>::MatCommaInitializer_(Mat_<T>* m, T first) : m_(m), idx_(0) { (*this), first; }
template <typename T> MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(T v)
{
    if (idx_ >= m_->rows * m_->cols) throw std::runtime_error("compat cv::Mat_: too many initializers");
    m_->template at<T>(idx_ / m_->cols, idx_ % m_->cols) = v;
    idx_++;
    return *this;
}
template <typename T> MatCommaInitializer_<T>::operator Mat() const { return *m_; }
template <typename T> MatCommaInitializer_<T>::operator Mat_<T>() const { return *m_; }

// cv::InputArray / noArray() as far as phaseCorrelate's signature needs them
struct _NoArray {};
inline _NoArray noArray() { return _NoArray(); }

class SVD {
public:
    Mat u, w, vt;
    SVD() {}
    explicit SVD(const Mat& src, int flags = 0);   // declared only
};

// ---- small host utilities the reference's drivers use (align_test.cpp:24,361): plain loops, nothing on the hot path
// cv::normalize(src, dst, alpha, beta, NORM_MINMAX): affine map of [min, max] onto [alpha, beta], same depth
inline void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type)
{
    if (norm_type != NORM_MINMAX || src.channels() != 1 || (src.depth() != CV_32F && src.depth() != CV_8U))
        throw std::runtime_error("compat cv::normalize: NORM_MINMAX on single-channel 8U / 32F only");
    double lo = 1e300, hi = -1e300;
    for (int r = 0; r < src.rows; r++)
        for (int c = 0; c < src.cols; c++) {
            const double v = src.depth() == CV_32F ? (double)src.ptr<float>(r)[c] : (double)src.ptr<uint8_t>(r)[c];
            lo = std::min(lo, v); hi = std::max(hi, v);
        }
    const double scale = hi > lo ? (beta - alpha) / (hi - lo) : 0.0;
    Mat out(src.rows, src.cols, src.type());
    for (int r = 0; r < src.rows; r++)
        for (int c = 0; c < src.cols; c++) {
            if (src.depth() == CV_32F) out.ptr<float>(r)[c] = (float)((src.ptr<float>(r)[c] - lo) * scale + alpha);
            else out.ptr<uint8_t>(r)[c] = (uint8_t)std::min(255.0, std::max(0.0, nearbyint((src.ptr<uint8_t>(r)[c] - lo) * scale + alpha)));
        }
    dst = out;
}

// cv::rectangle(img, rect, color, FILLED) on 8-bit images
inline void rectangle(Mat& img, Rect rc, const Scalar& color, int thickness = 1)
{
    if (img.depth() != CV_8U) throw std::runtime_error("compat cv::rectangle: 8-bit images only");
    const int cn = img.channels();
    for (int y = std::max(0, rc.y); y < std::min(img.rows, rc.y + rc.height); y++)
        for (int x = std::max(0, rc.x); x < std::min(img.cols, rc.x + rc.width); x++) {
            const bool edge = y == rc.y || y == rc.y + rc.height - 1 || x == rc.x || x == rc.x + rc.width - 1;
            if (thickness != FILLED && !edge) continue;
            for (int k = 0; k < cn; k++) img.ptr<uint8_t>(y)[x * cn + k] = (uint8_t)color.val[k < 4 ? k : 3];
        }
}

// ---- image / video I/O of the reference's drivers (align_test.cpp:45,632,682; video_test.cpp:62-117): declared only.
// The product never reads or writes files; a caller links OpenCV's own, the tests link oracle/ref_shim/cv_io.cpp.
Mat imread(const std::string& filename, int flags = IMREAD_COLOR);
bool imwrite(const std::string& filename, const Mat& img);

class VideoCapture {
public:
    VideoCapture();
    explicit VideoCapture(const std::string& filename);
    ~VideoCapture();
    bool open(const std::string& filename);
    bool isOpened() const;
    double get(int prop) const;
    bool read(Mat& frame);
    void release();
private:
    struct Impl;
    std::shared_ptr<Impl> impl_;
};

class VideoWriter {
public:
    VideoWriter();
    ~VideoWriter();
    static int fourcc(char a, char b, char c, char d) { return (a & 255) | ((b & 255) << 8) | ((c & 255) << 16) | ((d & 255) << 24); }
    bool open(const std::string& filename, int fourcc, double fps, Size frameSize, bool isColor = true);
    bool isOpened() const;
    bool set(int prop, double value);
    void write(const Mat& frame);
    void release();
private:
    struct Impl;
    std::shared_ptr<Impl> impl_;
};

// ---- declared only: CPU definitions live in oracle/ref_shim/cv_impl.cpp (test infrastructure)
void cvtColor(const Mat& src, Mat& dst, int code);
void warpAffine(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags = INTER_LINEAR,
                int borderMode = BORDER_CONSTANT, const Scalar& borderValue = Scalar());
Point2d phaseCorrelate(const Mat& src1, const Mat& src2, _NoArray window = _NoArray(), double* response = nullptr);
Mat operator*(const Mat& a, const Mat& b);

}  // namespace cv
