// compat/HalideRuntime.h — the few C types of Halide's runtime ABI that appear in the public
// signatures of the video_stabilizer operator API (imgproc.hpp).  Used only when the real
// Halide headers are absent; layout-compatible with Halide 15-19's halide_buffer_t so that a
// build against the real headers can swap in without touching callers.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum halide_type_code_t { halide_type_int = 0, halide_type_uint = 1, halide_type_float = 2, halide_type_handle = 3 } halide_type_code_t;

struct halide_type_t {
    uint8_t code;
    uint8_t bits;
    uint16_t lanes;
};

typedef struct halide_dimension_t {
    int32_t min, extent, stride;
    uint32_t flags;
#ifdef __cplusplus
    halide_dimension_t() : min(0), extent(0), stride(0), flags(0) {}
    halide_dimension_t(int32_t m, int32_t e, int32_t s, uint32_t f = 0) : min(m), extent(e), stride(s), flags(f) {}
#endif
} halide_dimension_t;

struct halide_device_interface_t;

typedef struct halide_buffer_t {
    uint64_t device;
    const struct halide_device_interface_t* device_interface;
    uint8_t* host;
    uint64_t flags;
    struct halide_type_t type;
    int32_t dimensions;
    halide_dimension_t* dim;
    void* padding;
} halide_buffer_t;

#ifdef __cplusplus
}
#endif
