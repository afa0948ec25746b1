// compat/HalideBuffer.h — minimal Halide::Runtime::Buffer<T> for builds without Halide.
//
// The reference's operator API (imgproc.hpp:8-97) passes host images as
// Halide::Runtime::Buffer<T>&.  This header provides the surface those signatures and their
// callers (alignment.cpp, align_test.cpp) use: dense planar allocation, shared ownership on
// copy, non-owning wrap of external memory, dim(i).extent()/min()/stride(), element access
// and the implicit conversion to halide_buffer_t* that AOT pipelines take.
#pragma once
#include <memory>
#include <stdexcept>
#include <string.h>
#include <type_traits>
#include <vector>

#include "HalideRuntime.h"

namespace Halide {
namespace Runtime {

template <typename T>
struct BufferTypeOf;
template <> struct BufferTypeOf<uint8_t>  { static halide_type_t get() { return {halide_type_uint, 8, 1}; } };
template <> struct BufferTypeOf<uint16_t> { static halide_type_t get() { return {halide_type_uint, 16, 1}; } };
template <> struct BufferTypeOf<uint32_t> { static halide_type_t get() { return {halide_type_uint, 32, 1}; } };
template <> struct BufferTypeOf<int32_t>  { static halide_type_t get() { return {halide_type_int, 32, 1}; } };
template <> struct BufferTypeOf<float>    { static halide_type_t get() { return {halide_type_float, 32, 1}; } };
template <> struct BufferTypeOf<double>   { static halide_type_t get() { return {halide_type_float, 64, 1}; } };

template <typename T>
class Buffer {
public:
    static constexpr int kMaxDims = 4;
    using ElemT = typename std::remove_const<T>::type;

    class Dimension {
    public:
        explicit Dimension(const halide_dimension_t& d) : d_(d) {}
        int min() const { return d_.min; }
        int extent() const { return d_.extent; }
        int stride() const { return d_.stride; }
        int max() const { return d_.min + d_.extent - 1; }
    private:
        halide_dimension_t d_;
    };

    Buffer() { init_header(0); }
    explicit Buffer(int e0) { int e[1] = {e0}; allocate(1, e); }
    Buffer(int e0, int e1) { int e[2] = {e0, e1}; allocate(2, e); }
    Buffer(int e0, int e1, int e2) { int e[3] = {e0, e1, e2}; allocate(3, e); }
    Buffer(int e0, int e1, int e2, int e3) { int e[4] = {e0, e1, e2, e3}; allocate(4, e); }
    // non-owning wrap (imgproc.cpp:227-231)
    Buffer(T* data, int dims, const halide_dimension_t* shape)
    {
        init_header(dims);
        for (int i = 0; i < dims; i++) dims_[i] = shape[i];
        buf_.host = (uint8_t*)const_cast<ElemT*>(data);
    }
    Buffer(T* data, int e0, int e1)
    {
        init_header(2);
        dims_[0] = halide_dimension_t(0, e0, 1);
        dims_[1] = halide_dimension_t(0, e1, e0);
        buf_.host = (uint8_t*)const_cast<ElemT*>(data);
    }
    Buffer(const Buffer& o) { copy_from(o); }
    Buffer& operator=(const Buffer& o) { if (this != &o) copy_from(o); return *this; }

    int dimensions() const { return buf_.dimensions; }
    Dimension dim(int i) const { return Dimension(dims_[i]); }
    int min(int i) const { return dims_[i].min; }
    int extent(int i) const { return dims_[i].extent; }
    int stride(int i) const { return dims_[i].stride; }
    int width() const { return buf_.dimensions > 0 ? dims_[0].extent : 1; }
    int height() const { return buf_.dimensions > 1 ? dims_[1].extent : 1; }
    int channels() const { return buf_.dimensions > 2 ? dims_[2].extent : 1; }
    size_t number_of_elements() const
    {
        size_t n = 1;
        for (int i = 0; i < buf_.dimensions; i++) n *= (size_t)dims_[i].extent;
        return n;
    }
    size_t size_in_bytes() const { return number_of_elements() * sizeof(ElemT); }
    T* data() const { return (T*)buf_.host; }
    T* begin() const { return data(); }
    bool owns_host_memory() const { return (bool)alloc_; }

    T& operator()(int x) const { return data()[off(0, x)]; }
    T& operator()(int x, int y) const { return data()[off(0, x) + off(1, y)]; }
    T& operator()(int x, int y, int c) const { return data()[off(0, x) + off(1, y) + off(2, c)]; }

    void fill(ElemT v) const
    {
        // dense allocations only (everything this project fills is one)
        ElemT* p = (ElemT*)buf_.host;
        for (size_t i = 0, n = number_of_elements(); i < n; i++) p[i] = v;
    }

    halide_buffer_t* raw_buffer() { return &buf_; }
    const halide_buffer_t* raw_buffer() const { return &buf_; }
    operator halide_buffer_t*() { return &buf_; }
    // no-ops kept so host code written for GPU-enabled Halide builds compiles
    void set_host_dirty(bool = true) {}
    void copy_to_host() {}

private:
    halide_buffer_t buf_;
    halide_dimension_t dims_[kMaxDims];
    std::shared_ptr<std::vector<ElemT>> alloc_;

    ptrdiff_t off(int d, int i) const { return (ptrdiff_t)(i - dims_[d].min) * dims_[d].stride; }
    void init_header(int dims)
    {
        if (dims > kMaxDims) throw std::runtime_error("compat Halide::Runtime::Buffer: too many dimensions");
        memset(&buf_, 0, sizeof(buf_));
        buf_.type = BufferTypeOf<ElemT>::get();
        buf_.dimensions = dims;
        buf_.dim = dims_;
        for (int i = 0; i < kMaxDims; i++) dims_[i] = halide_dimension_t();
    }
    void allocate(int dims, const int* extents)
    {
        init_header(dims);
        size_t n = 1;
        for (int i = 0; i < dims; i++) {
            dims_[i] = halide_dimension_t(0, extents[i], (int32_t)n);
            n *= (size_t)(extents[i] < 0 ? 0 : extents[i]);
        }
        alloc_ = std::make_shared<std::vector<ElemT>>(n);   // zero-initialised
        buf_.host = (uint8_t*)alloc_->data();
    }
    void copy_from(const Buffer& o)
    {
        buf_ = o.buf_;
        for (int i = 0; i < kMaxDims; i++) dims_[i] = o.dims_[i];
        buf_.dim = dims_;
        alloc_ = o.alloc_;
    }
};

}  // namespace Runtime
}  // namespace Halide
