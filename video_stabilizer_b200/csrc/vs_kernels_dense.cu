// vs_kernels_dense.cu — full-frame kernels: BGR->gray, pyr_down, grad_xy, image_warp, BGR warp.
// All are HBM-bound byte/integer work: coalesced vector loads, packed 16-bit SIMD-in-register
// arithmetic where the math allows it, no tensor cores (nothing here is a contraction).
#include "vs_internal.h"

#include <cuda.h>   // CUtensorMap; the encoder is resolved at run time (vs_tensor_map_encoder)
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace {

// ------------------------------------------------------------------ BGR -> gray
// cv::cvtColor(BGR2GRAY) (alignment.cpp:212): (3735 B + 19235 G + 9798 R + 16384) >> 15.
// One thread converts 16 pixels: 3 x 16-byte loads, 1 x 16-byte store.
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t* v, int k)
{
    return (v[k >> 2] >> (8 * (k & 3))) & 0xffu;
}

__global__ void __launch_bounds__(128)
k_bgr2gray(const uint8_t* __restrict__ bgr, int64_t in_stride, int64_t in_bs,
           uint8_t* __restrict__ gray, int64_t out_stride, int64_t out_bs, int w, int h, int vec_ok)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y;
    if (x0 >= w) return;
    const uint8_t* src = bgr + (size_t)blockIdx.z * in_bs + (size_t)y * in_stride + (size_t)x0 * 3;
    uint8_t* dst = gray + (size_t)blockIdx.z * out_bs + (size_t)y * out_stride + x0;
    if (vec_ok && x0 + 16 <= w) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
        uint32_t v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int p = q * 4 + i;
                acc |= gray_of(byte_of(v, 3 * p), byte_of(v, 3 * p + 1), byte_of(v, 3 * p + 2)) << (8 * i);
            }
            o[q] = acc;
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        int n = min(16, w - x0);
        for (int i = 0; i < n; i++)
            dst[i] = (uint8_t)gray_of(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
    }
}

// ------------------------------------------------------------------ pyr_down
// generators.cpp:56-92: out(x,y) = (sum_j sum_i k_j k_i in(clamp(2x+i), clamp(2y+j))) >> 8,
// k = [1 4 6 4 1].  Streaming, register-only: a thread owns 4 adjacent output columns and
// walks down PD_ROWS output rows.  Per input row it loads the 12 bytes it needs as one
// 8-byte word pair plus the two/one halo bytes on either side and forms the four horizontal
// sums with IDP.4A against constant weight words; the vertical pass runs on a sliding window
// of five rows, two outputs per register in 16-bit lanes (horizontal sums <= 4080, vertical
// sums <= 65280: no carry between lanes).  The truncating >> 8 and the packing of the four
// output bytes are one PRMT.  No shared memory, no barriers; the 3-row halo of a strip is
// re-read from L1/L2.
constexpr int PD_ROWS_DEFAULT = 8;

// horizontal sums of outputs x..x+3 (x = first output column) from
//   a16 = in[2x-2], in[2x-1] (low two bytes)   b = in[2x .. 2x+7]   c8 = in[2x+8]
__device__ __forceinline__ void pd_hsum(uint32_t a16, uint2 b, uint32_t c8, uint32_t& h01, uint32_t& h23)
{
    const uint32_t h0 = __dp4a(b.x, 0x00010406u, __dp4a(a16, 0x00000401u, 0u));
    const uint32_t h1 = __dp4a(b.y, 0x00000001u, __dp4a(b.x, 0x04060401u, 0u));
    const uint32_t h2 = __dp4a(b.y, 0x00010406u, __dp4a(b.x, 0x04010000u, 0u));
    const uint32_t h3 = __dp4a(b.y, 0x04060401u, c8);
    h01 = h0 | (h1 << 16);
    h23 = h2 | (h3 << 16);
}

template <bool FAST>
__device__ __forceinline__ void pd_load_row(const uint8_t* __restrict__ src, int64_t stride, int iw, int ih, int row, int x4,
                                            uint32_t& h01, uint32_t& h23)
{
    const uint8_t* r = src + (size_t)vs_clampi(row, 0, ih - 1) * stride;
    uint32_t a16, c8;
    uint2 b;
    if (FAST) {
        // 2*x4 .. 2*x4+7 inside the row, 8-byte aligned
        b = __ldg(reinterpret_cast<const uint2*>(r + 2 * x4));
        a16 = x4 > 0 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(r + 2 * x4 - 2)) : (b.x & 0xffu) * 0x0101u;
        c8 = __ldg(r + min(2 * x4 + 8, iw - 1));
    } else {
        uint32_t v[11];
#pragma unroll
        for (int i = 0; i < 11; i++) v[i] = __ldg(r + vs_clampi(2 * x4 - 2 + i, 0, iw - 1));
        a16 = v[0] | (v[1] << 8);
        b.x = v[2] | (v[3] << 8) | (v[4] << 16) | (v[5] << 24);
        b.y = v[6] | (v[7] << 8) | (v[8] << 16) | (v[9] << 24);
        c8 = v[10];
    }
    pd_hsum(a16, b, c8, h01, h23);
}

template <bool FAST, int PD_ROWS, bool FULL>
__device__ __forceinline__ void pd_strip(const uint8_t* __restrict__ src, int64_t in_stride, int iw, int ih,
                                         uint8_t* __restrict__ dst, int64_t out_stride, int ow, int oh, int x4, int y0, bool store_word)
{
    // FULL strips (all PD_ROWS outputs exist) have no exit inside the unrolled loop, so the loads of
    // all 2*PD_ROWS+3 input rows can be scheduled ahead of the arithmetic: that memory-level
    // parallelism is what a latency-bound streaming kernel lives on
    uint32_t h01[2 * PD_ROWS + 3], h23[2 * PD_ROWS + 3];
    if (FULL) {
#pragma unroll
        for (int j = 0; j < 2 * PD_ROWS + 3; j++) pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y0 - 2 + j, x4, h01[j], h23[j]);
#pragma unroll
        for (int i = 0; i < PD_ROWS; i++) {
            const uint32_t v01 = h01[2 * i] + h01[2 * i + 4] + 4u * (h01[2 * i + 1] + h01[2 * i + 3]) + 6u * h01[2 * i + 2];
            const uint32_t v23 = h23[2 * i] + h23[2 * i + 4] + 4u * (h23[2 * i + 1] + h23[2 * i + 3]) + 6u * h23[2 * i + 2];
            const uint32_t packed = __byte_perm(v01, v23, 0x7531);   // (v >> 8) of the four 16-bit lanes
            uint8_t* o = dst + (size_t)(y0 + i) * out_stride + x4;
            if (store_word) {
                *reinterpret_cast<uint32_t*>(o) = packed;
            } else {
                for (int k = 0; k < 4 && x4 + k < ow; k++) o[k] = (uint8_t)(packed >> (8 * k));
            }
        }
        return;
    }
    uint32_t a01, a23, b01, b23, c01, c23;
    pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y0 - 2, x4, a01, a23);
    pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y0 - 1, x4, b01, b23);
    pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y0, x4, c01, c23);
    for (int i = 0; i < PD_ROWS; i++) {
        const int y = y0 + i;
        if (y >= oh) break;
        uint32_t d01, d23, e01, e23;
        pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y + 1, x4, d01, d23);
        pd_load_row<FAST>(src, in_stride, iw, ih, 2 * y + 2, x4, e01, e23);
        const uint32_t v01 = a01 + e01 + 4u * (b01 + d01) + 6u * c01;
        const uint32_t v23 = a23 + e23 + 4u * (b23 + d23) + 6u * c23;
        const uint32_t packed = __byte_perm(v01, v23, 0x7531);
        uint8_t* o = dst + (size_t)y * out_stride + x4;
        if (store_word) {
            *reinterpret_cast<uint32_t*>(o) = packed;
        } else {
            for (int k = 0; k < 4 && x4 + k < ow; k++) o[k] = (uint8_t)(packed >> (8 * k));
        }
        a01 = c01; a23 = c23; b01 = d01; b23 = d23; c01 = e01; c23 = e23;
    }
}

// Wide variant: 8 adjacent outputs per thread from one 16-byte load per input row (+ the same
// 2 + 1 halo bytes): twice the bytes in flight per load instruction.
//   a16 = in[2x-2], in[2x-1]   b = in[2x .. 2x+15]   c8 = in[2x+16]
__device__ __forceinline__ void pd_hsum8(uint32_t a16, uint4 b, uint32_t c8, uint32_t h[4])
{
    const uint32_t h0 = __dp4a(b.x, 0x00010406u, __dp4a(a16, 0x00000401u, 0u));
    const uint32_t h1 = __dp4a(b.y, 0x00000001u, __dp4a(b.x, 0x04060401u, 0u));
    const uint32_t h2 = __dp4a(b.y, 0x00010406u, __dp4a(b.x, 0x04010000u, 0u));
    const uint32_t h3 = __dp4a(b.z, 0x00000001u, __dp4a(b.y, 0x04060401u, 0u));
    const uint32_t h4 = __dp4a(b.z, 0x00010406u, __dp4a(b.y, 0x04010000u, 0u));
    const uint32_t h5 = __dp4a(b.w, 0x00000001u, __dp4a(b.z, 0x04060401u, 0u));
    const uint32_t h6 = __dp4a(b.w, 0x00010406u, __dp4a(b.z, 0x04010000u, 0u));
    const uint32_t h7 = __dp4a(b.w, 0x04060401u, c8);
    h[0] = h0 | (h1 << 16); h[1] = h2 | (h3 << 16); h[2] = h4 | (h5 << 16); h[3] = h6 | (h7 << 16);
}

template <int PD_ROWS>
__global__ void __launch_bounds__(128)
k_pyr_down_wide(const uint8_t* __restrict__ in, int64_t in_stride, int64_t in_bs, int iw, int ih,
                uint8_t* __restrict__ out, int64_t out_stride, int64_t out_bs, int ow, int oh)
{
    // launched only when rows are 16-byte aligned, ow % 8 == 0 and 2*ow <= iw (every 16-byte load inside the row)
    const int x8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * PD_ROWS;
    if (x8 >= ow || y0 >= oh) return;
    const uint8_t* src = in + (size_t)blockIdx.z * in_bs;
    uint8_t* dst = out + (size_t)blockIdx.z * out_bs;
    uint32_t h[2 * PD_ROWS + 3][4];
#pragma unroll
    for (int j = 0; j < 2 * PD_ROWS + 3; j++) {
        const uint8_t* r = src + (size_t)vs_clampi(2 * y0 - 2 + j, 0, ih - 1) * in_stride;
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(r + 2 * x8));
        const uint32_t a16 = x8 > 0 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(r + 2 * x8 - 2)) : (b.x & 0xffu) * 0x0101u;
        const uint32_t c8 = __ldg(r + min(2 * x8 + 16, iw - 1));
        pd_hsum8(a16, b, c8, h[j]);
    }
#pragma unroll
    for (int i = 0; i < PD_ROWS; i++) {
        if (y0 + i >= oh) break;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            v[k] = h[2 * i][k] + h[2 * i + 4][k] + 4u * (h[2 * i + 1][k] + h[2 * i + 3][k]) + 6u * h[2 * i + 2][k];
        *reinterpret_cast<uint2*>(dst + (size_t)(y0 + i) * out_stride + x8) =
            make_uint2(__byte_perm(v[0], v[1], 0x7531), __byte_perm(v[2], v[3], 0x7531));
    }
}

template <int PD_ROWS, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_pyr_down(const uint8_t* __restrict__ in, int64_t in_stride, int64_t in_bs, int iw, int ih,
           uint8_t* __restrict__ out, int64_t out_stride, int64_t out_bs, int ow, int oh, int in_al8, int out_al4)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * PD_ROWS;
    if (x4 >= ow || y0 >= oh) return;
    const uint8_t* src = in + (size_t)blockIdx.z * in_bs;
    uint8_t* dst = out + (size_t)blockIdx.z * out_bs;
    const bool store_word = out_al4 && x4 + 4 <= ow;
    const bool fast = in_al8 && 2 * x4 + 8 <= iw;
    if (fast && y0 + PD_ROWS <= oh)
        pd_strip<true, PD_ROWS, true>(src, in_stride, iw, ih, dst, out_stride, ow, oh, x4, y0, store_word);
    else if (fast)
        pd_strip<true, PD_ROWS, false>(src, in_stride, iw, ih, dst, out_stride, ow, oh, x4, y0, store_word);
    else
        pd_strip<false, PD_ROWS, false>(src, in_stride, iw, ih, dst, out_stride, ow, oh, x4, y0, store_word);
}

// ------------------------------------------------------------------ fused ingest: BGR -> gray L0 + pyramid L1
// cv::cvtColor(BGR2GRAY) (alignment.cpp:212) and the first PyrDown (alignment.cpp:220-223, generators.cpp:56-92) in one
// pass over the interleaved frame: the gray level is written once and never read back (unfused, 1080p: 2.07 MB written by
// BGR->gray and read again by pyr_down, per frame).  Register streaming, no shared memory, no barriers: a thread owns 16
// level-0 columns (48 BGR bytes = three 16-byte loads per row, one 16-byte gray store) = 8 level-1 columns (one 8-byte
// store per level-1 row) and walks down a strip of IG_L1_ROWS level-1 rows with a five-row window of horizontal
// [1 4 6 4 1] sums (IDP.4A, two 16-bit sums per register, exactly as k_pyr_down_wide).  The three gray pixels a thread
// needs from its neighbours (two on the left, one on the right) come by warp shuffle; only the first / last lane of a
// warp fetches them from the BGR frame itself.  The loads of the next two rows are issued before the current two are
// converted.  Gray: t = 2 (3735 B + 19235 G + 9798 R) + 32768 as two IDP.4A (weights split into high and low bytes) and one
// multiply-add per pixel; the gray value is byte 2 of t, so four pixels are packed by three PRMTs.
constexpr int IG_L1_ROWS = 36;          // level-1 rows per strip: 72 level-0 rows + 3 halo rows (4 % re-read)

__device__ __forceinline__ uint32_t ig_t(uint32_t p)
{
    return __dp4a(p, 0x004C961Du, 0u) * 256u + __dp4a(p, 0x008C462Eu, 32768u);
}

// four gray pixels from 12 interleaved bytes (three words)
__device__ __forceinline__ uint32_t ig_gray4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    const uint32_t t0 = ig_t(w0), t1 = ig_t(__byte_perm(w0, w1, 0x0543)), t2 = ig_t(__byte_perm(w1, w2, 0x0432)), t3 = ig_t(w2 >> 8);
    return __byte_perm(__byte_perm(t0, t1, 0x0062), __byte_perm(t2, t3, 0x0062), 0x5410);
}

struct IgRow {
    uint4 a, b, c;      // the 48 BGR bytes of this thread's 16 pixels
    uint2 l;            // the 8 bytes before them (first lane of a warp, not at the image edge)
    uint32_t r;         // the 4 bytes after them (last lane of a warp, not at the image edge)
};

__device__ __forceinline__ void ig_load(const uint8_t* __restrict__ src, int64_t stride, int h, int row, int xb,
                                        bool need_l, bool need_r, IgRow& o)
{
    const uint8_t* p = src + (size_t)vs_clampi(row, 0, h - 1) * stride + xb;
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
    o.a = __ldg(p4); o.b = __ldg(p4 + 1); o.c = __ldg(p4 + 2);
    o.l = need_l ? __ldg(reinterpret_cast<const uint2*>(p - 8)) : make_uint2(0u, 0u);
    o.r = need_r ? __ldg(reinterpret_cast<const uint32_t*>(p + 48)) : 0u;
}

// gray of the 16 pixels (g) and the horizontal sums of the 8 level-1 outputs (hs)
__device__ __forceinline__ void ig_convert(const IgRow& in, bool first_col, bool last_col, bool need_l, bool need_r,
                                           uint4& g, uint32_t hs[4])
{
    g.x = ig_gray4(in.a.x, in.a.y, in.a.z);
    g.y = ig_gray4(in.a.w, in.b.x, in.b.y);
    g.z = ig_gray4(in.b.z, in.b.w, in.c.x);
    g.w = ig_gray4(in.c.y, in.c.z, in.c.w);
    uint32_t a16 = __shfl_up_sync(0xffffffffu, g.w, 1) >> 16;        // pixels x0-2, x0-1 of the lane to the left
    uint32_t c8 = __shfl_down_sync(0xffffffffu, g.x, 1) & 0xffu;     // pixel x0+16 of the lane to the right
    if (need_l) {
        const uint32_t t0 = ig_t(__byte_perm(in.l.x, in.l.y, 0x0432)), t1 = ig_t(in.l.y >> 8);
        a16 = __byte_perm(t0, t1, 0x0062) & 0xffffu;
    }
    if (need_r) c8 = ig_t(in.r) >> 16;
    if (first_col) a16 = (g.x & 0xffu) * 0x0101u;                      // repeat-edge
    if (last_col) c8 = g.w >> 24;
    pd_hsum8(a16, g, c8, hs);
}

__global__ void __launch_bounds__(128)
k_ingest_bgr_gray_l1(const uint8_t* __restrict__ bgr, int64_t bgr_stride, int64_t bgr_bs,
                     uint8_t* __restrict__ g0, int64_t g0_stride, int64_t g0_bs,
                     uint8_t* __restrict__ g1, int64_t g1_stride, int64_t g1_bs, int w, int h, int ow, int oh)
{
    // launched only when w % 16 == 0, ow == w / 2, oh == h / 2 and all rows are 16-byte (level 1: 8-byte) aligned
    const int lane = threadIdx.x;
    const int col = blockIdx.x * 32 + lane;                      // 16-pixel column group
    const int ncols = w >> 4;
    const bool live = col < ncols;
    const int x0 = min(col, ncols - 1) << 4;                    // lanes past the row shadow the last group (shuffles need them)
    const int strip = blockIdx.y * blockDim.y + threadIdx.y;
    const int Y0 = strip * IG_L1_ROWS;
    if (Y0 >= oh) return;                                        // a whole warp leaves together (one strip per warp)
    const int Y1 = min(Y0 + IG_L1_ROWS, oh);
    const bool last_strip = Y1 >= oh;
    const uint8_t* src = bgr + (size_t)blockIdx.z * bgr_bs;
    uint8_t* d0 = g0 + (size_t)blockIdx.z * g0_bs + x0;
    uint8_t* d1 = g1 + (size_t)blockIdx.z * g1_bs + (x0 >> 1);
    const bool first_col = x0 == 0, last_col = x0 + 16 >= w;
    const bool need_l = lane == 0 && !first_col, need_r = lane == 31 && !last_col;
    const int xb = x0 * 3;

    // level-0 row r is written by the strip whose range [2 Y0, 2 Y1) holds it; the last strip also owns the rows below
    // 2 oh (one row when h is odd)
    auto store0 = [&](int r, const uint4& g) {
        if (live && r >= 2 * Y0 && r < h && (r < 2 * Y1 || last_strip))
            *reinterpret_cast<uint4*>(d0 + (size_t)r * g0_stride) = g;
    };

    uint32_t hA[4], hB[4], hC[4];
    {
        IgRow ra, rb, rc;
        ig_load(src, bgr_stride, h, 2 * Y0 - 2, xb, need_l, need_r, ra);
        ig_load(src, bgr_stride, h, 2 * Y0 - 1, xb, need_l, need_r, rb);
        ig_load(src, bgr_stride, h, 2 * Y0, xb, need_l, need_r, rc);
        uint4 g;
        ig_convert(ra, first_col, last_col, need_l, need_r, g, hA);
        ig_convert(rb, first_col, last_col, need_l, need_r, g, hB);
        ig_convert(rc, first_col, last_col, need_l, need_r, g, hC);
        store0(2 * Y0, g);
    }
    IgRow nd, ne;
    ig_load(src, bgr_stride, h, 2 * Y0 + 1, xb, need_l, need_r, nd);
    ig_load(src, bgr_stride, h, 2 * Y0 + 2, xb, need_l, need_r, ne);
#pragma unroll 1
    for (int y = Y0; y < Y1; y++) {
        const IgRow rd = nd, re = ne;
        if (y + 1 < Y1) {                                        // rows of the next iteration: in flight during this one
            ig_load(src, bgr_stride, h, 2 * y + 3, xb, need_l, need_r, nd);
            ig_load(src, bgr_stride, h, 2 * y + 4, xb, need_l, need_r, ne);
        }
        uint32_t hD[4], hE[4];
        uint4 gd, ge;
        ig_convert(rd, first_col, last_col, need_l, need_r, gd, hD);
        ig_convert(re, first_col, last_col, need_l, need_r, ge, hE);
        store0(2 * y + 1, gd);
        store0(2 * y + 2, ge);
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = hA[k] + hE[k] + 4u * (hB[k] + hD[k]) + 6u * hC[k];
        if (live)
            *reinterpret_cast<uint2*>(d1 + (size_t)y * g1_stride) =
                make_uint2(__byte_perm(v[0], v[1], 0x7531), __byte_perm(v[2], v[3], 0x7531));
#pragma unroll
        for (int k = 0; k < 4; k++) { hA[k] = hC[k]; hB[k] = hD[k]; hC[k] = hE[k]; }
    }
}

// ------------------------------------------------------------------ grad_xy
// generators.cpp:202-224: 0.5*(I(x+1,y)-I(x-1,y)), 0.5*(I(x,y+1)-I(x,y-1)), repeat-edge.
__global__ void __launch_bounds__(128)
k_grad_xy(const uint8_t* __restrict__ in, int64_t in_stride, int64_t in_bs, int iw, int ih,
          float* __restrict__ gx, int64_t gx_stride, int64_t gx_bs,
          float* __restrict__ gy, int64_t gy_stride, int64_t gy_bs, int ow, int oh, int vec_ok)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= ow) return;
    const uint8_t* src = in + (size_t)blockIdx.z * in_bs;
    const uint8_t* rc = src + (size_t)vs_clampi(y, 0, ih - 1) * in_stride;
    const uint8_t* rp = src + (size_t)vs_clampi(y + 1, 0, ih - 1) * in_stride;
    const uint8_t* rm = src + (size_t)vs_clampi(y - 1, 0, ih - 1) * in_stride;
    float ax[4], ay[4];
    float c[6];
#pragma unroll
    for (int i = 0; i < 6; i++) c[i] = (float)__ldg(rc + vs_clampi(x0 - 1 + i, 0, iw - 1));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xc = vs_clampi(x0 + i, 0, iw - 1);
        ax[i] = __fmul_rn(0.5f, __fsub_rn(c[i + 2], c[i]));
        ay[i] = __fmul_rn(0.5f, __fsub_rn((float)__ldg(rp + xc), (float)__ldg(rm + xc)));
    }
    float* ox = gx + (size_t)blockIdx.z * gx_bs + (size_t)y * gx_stride + x0;
    float* oy = gy + (size_t)blockIdx.z * gy_bs + (size_t)y * gy_stride + x0;
    if (vec_ok && x0 + 4 <= ow) {
        *reinterpret_cast<float4*>(ox) = make_float4(ax[0], ax[1], ax[2], ax[3]);
        *reinterpret_cast<float4*>(oy) = make_float4(ay[0], ay[1], ay[2], ay[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < ow; i++) { ox[i] = ax[i]; oy[i] = ay[i]; }
    }
}

// ------------------------------------------------------------------ image_warp
// generators.cpp:126-164: pull-mapped bilinear with repeat-edge; Halide's float lerp is
// a*(1-t) + b*t.  params: {A,B,TX,TY} f32 per batch image (UL origin, imgproc.cpp:125-131).
__global__ void __launch_bounds__(128)
k_image_warp(const uint8_t* __restrict__ in, int64_t in_stride, int64_t in_bs, int iw, int ih,
             const float* __restrict__ params, float* __restrict__ out, int64_t out_stride,
             int64_t out_bs, int ow, int oh, int vec_ok)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= ow) return;
    const uint8_t* src = in + (size_t)blockIdx.z * in_bs;
    const float A = params[blockIdx.z * 4 + 0], B = params[blockIdx.z * 4 + 1];
    const float TX = params[blockIdx.z * 4 + 2], TY = params[blockIdx.z * 4 + 3];
    const float onepA = __fadd_rn(1.0f, A);
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float fx = (float)(x0 + i), fy = (float)y;
        float Wx = __fadd_rn(__fsub_rn(__fmul_rn(onepA, fx), __fmul_rn(B, fy)), TX);
        float Wy = __fadd_rn(__fadd_rn(__fmul_rn(B, fx), __fmul_rn(onepA, fy)), TY);
        int flx = (int)floorf(Wx), fly = (int)floorf(Wy);
        float wx = __fsub_rn(Wx, (float)flx), wy = __fsub_rn(Wy, (float)fly);
        int xa = vs_clampi(flx, 0, iw - 1), xb = vs_clampi(flx + 1, 0, iw - 1);
        int ya = vs_clampi(fly, 0, ih - 1), yb = vs_clampi(fly + 1, 0, ih - 1);
        float p00 = (float)__ldg(src + (size_t)ya * in_stride + xa);
        float p10 = (float)__ldg(src + (size_t)ya * in_stride + xb);
        float p01 = (float)__ldg(src + (size_t)yb * in_stride + xa);
        float p11 = (float)__ldg(src + (size_t)yb * in_stride + xb);
        float omx = __fsub_rn(1.0f, wx), omy = __fsub_rn(1.0f, wy);
        float top = __fadd_rn(__fmul_rn(p00, omx), __fmul_rn(p10, wx));
        float bot = __fadd_rn(__fmul_rn(p01, omx), __fmul_rn(p11, wx));
        r[i] = __fadd_rn(__fmul_rn(top, omy), __fmul_rn(bot, wy));
    }
    float* o = out + (size_t)blockIdx.z * out_bs + (size_t)y * out_stride + x0;
    if (vec_ok && x0 + 4 <= ow) {
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < ow; i++) o[i] = r[i];
    }
}

// ------------------------------------------------------------------ the two bilinear modes: one arithmetic, two grids
// Both bilinear modes are exact integer arithmetic on a fixed-point grid:
//   position of output pixel (x, y):  sfx = rint(i00 x 2^P) + rint((i01 y + i02) 2^P) + 2^(P-W-1)   (likewise sfy),
//   source pixel (sfx >> P, sfy >> P), weight fractions fx, fy = the top W bits below the point,
//   out = (sum of (2^W - fx | fx)(2^W - fy | fy) taps + 2^(2W-1)) >> 2W per channel.
// VS_WARP_CV_EXACT_BILINEAR is cv::warpAffine's grid (AB_BITS = 10, INTER_BITS = 5: P = 10, W = 5; imgproc.cpp:446-484).
// VS_WARP_FLOAT_BILINEAR (no counterpart upstream) keeps 16 fractional position bits and 1/256-pixel weights (P = 16,
// W = 8): positions are good to 2^-16 px on frames up to 16384 pixels (an f32 coordinate is good to 2^-11 px at 8K), the
// weights are as fine as the 8-bit output can show.  Same kernels, same instruction count: the weights of one pixel sum to
// 2^(2W), and with the cv weights scaled by 64 both grids blend to byte 2 of a 32-bit accumulator.
template <int MODE> struct WarpGrid;
template <> struct WarpGrid<VS_WARP_CV_EXACT_BILINEAR> { static constexpr int P = 10, W = 5; };
template <> struct WarpGrid<VS_WARP_FLOAT_BILINEAR> { static constexpr int P = 16, W = 8; };
// Lanczos-2 (below): the fine grid with 64 tabulated weight fractions
template <> struct WarpGrid<VS_WARP_LANCZOS2> { static constexpr int P = 16, W = 6; };

// one output pixel from four BGRX taps, any fractions: returns B | G << 8 | R << 16
template <int MODE>
__device__ __forceinline__ uint32_t warp_blend_taps(uint32_t t00, uint32_t t10, uint32_t t01, uint32_t t11, int sfx, int sfy)
{
    constexpr int P = WarpGrid<MODE>::P, W = WarpGrid<MODE>::W;
    const uint32_t fx = ((uint32_t)sfx >> (P - W)) & ((1u << W) - 1u), fy = ((uint32_t)sfy >> (P - W)) & ((1u << W) - 1u);
    const uint32_t one = 1u << W, half = 1u << (2 * W - 1);
    const uint32_t w00 = (one - fx) * (one - fy), w10 = fx * (one - fy), w01 = (one - fx) * fy, w11 = fx * fy;
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const uint32_t v = w00 * ((t00 >> (8 * c)) & 0xffu) + w10 * ((t10 >> (8 * c)) & 0xffu) +
                           w01 * ((t01 >> (8 * c)) & 0xffu) + w11 * ((t11 >> (8 * c)) & 0xffu);
        out |= ((v + half) >> (2 * W)) << (8 * c);
    }
    return out;
}

// ------------------------------------------------------------------ BGR warp, cv-exact, tiled
// Mode 0 (the warp VideoStabilizer runs on every output frame), first tiled generation; k_bgr_warp_cv_rows below is
// the form used whenever the source can be described by a tensor map.
// A CTA of 128 threads produces a 120 x 16 pixel output tile (120 px = 360 B = 45 8-byte
// stores per row; 1920 and 3840 are multiples of 120; the source box of a tile is then at
// most 32 four-pixel granules wide, one per lane):
//   1. thread t owns output column t: its two column terms adelta/bdelta (the only f64 work)
//      are computed once and reused for the 16 rows; the 16 row terms X0/Y0 go to shared memory.
//   2. the source bounding box of the tile follows from its four corners (the fixed-point
//      map is a sum of a monotone function of x and a monotone function of y).  Warp w stages
//      box rows w, w+4, ...: lane l loads the 12 bytes of granule l (+4 bytes of look-ahead) of
//      each of its rows with 32-bit loads, all issued before the first is consumed, and writes
//      one 8-byte entry per source pixel x:
//          .x = (B[x], B[x+1], G[x], G[x+1])      .y = (R[x], R[x+1], 0, 0)
//      i.e. the two horizontal taps of every channel sit in adjacent bytes.  Texels outside
//      the image are staged as the border value (0, or the clamped edge texel), so sampling
//      needs no border logic.
//   3. a pixel is then two conflict-free 8-byte shared loads (top and bottom row) and six
//      IDP.2A dot products against the packed 16-bit weight pairs (w00|w10<<16, w01|w11<<16);
//      four lanes' pixels are packed into 12 bytes with one shuffle + PRMT per lane and
//      collected in a shared tile that is written out with 8-byte stores.
// A tile whose bounding box does not fit (large rotations or scales) takes the direct
// global-memory path, decided per CTA.
constexpr int WT_W = 120, WT_H = 16, WT_THREADS = 128, WT_WARPS = WT_THREADS / 32;
constexpr int WT_SRC_ENTRIES = 2560;                     // staged source capacity (20 KB): 32 granules x 20 rows
constexpr int WT_OUT_ROW_WORDS = WT_W * 3 / 4;           // 90 words: packed BGR bytes of one output row
constexpr int WT_OUT_WORDS = WT_OUT_ROW_WORDS * WT_H;    // 5.6 KB
constexpr int WT_SMEM_BYTES = WT_SRC_ENTRIES * 8 + WT_OUT_WORDS * 4;
constexpr int WT_ROWS_PER_WARP = 5;                      // staged rows per warp kept in flight (covers 20 rows)

template <int BORDER>
__device__ __forceinline__ uint32_t bgr_texel_word(const uint8_t* __restrict__ src, int64_t stride, int w, int h, int x, int y)
{
    if (BORDER == VS_BORDER_REPEAT_EDGE) {
        x = vs_clampi(x, 0, w - 1); y = vs_clampi(y, 0, h - 1);
    } else if (x < 0 || x >= w || y < 0 || y >= h) {
        return 0u;
    }
    const uint8_t* p = src + (size_t)y * stride + 3 * x;
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}

// staged entry of pixel x from the BGRX words of x and x+1
__device__ __forceinline__ uint2 wt_entry(uint32_t p, uint32_t pn)
{
    return make_uint2(__byte_perm(p, pn, 0x5140), __byte_perm(p, pn, 0x7362));
}

// One output pixel from the staged entries of (sx,sy) and (sx,sy+1).
// (sum w p + 16384) >> 15 with w = 32 wx wy  ==  (sum wx wy p + 512) >> 10
__device__ __forceinline__ uint32_t cv_blend(uint2 top, uint2 bot, int fx, int fy)
{
    const uint32_t hpair = (uint32_t)fx * 65535u + 32u;          // (32-fx) | fx << 16
    const uint32_t wt = (uint32_t)(32 - fy) * hpair;             // w00 | w10 << 16, each <= 1024
    const uint32_t wb = (uint32_t)fy * hpair;                    // w01 | w11 << 16
    uint32_t b = __dp2a_lo(wt, top.x, 512u), g = __dp2a_hi(wt, top.x, 512u), r = __dp2a_lo(wt, top.y, 512u);
    b = __dp2a_lo(wb, bot.x, b); g = __dp2a_hi(wb, bot.x, g); r = __dp2a_lo(wb, bot.y, r);
    return (b >> 10) | ((g >> 10) << 8) | ((r >> 10) << 16);
}

// one pixel from two staged entries: the cv grid blends with IDP.2A (weights <= 1024); the fine grid's weights reach
// 65536, so it unpacks the entry (this kernel is the fallback path: borders, small or unaligned images)
template <int MODE>
__device__ __forceinline__ uint32_t wt_blend(uint2 top, uint2 bot, int sfx, int sfy)
{
    if (MODE == VS_WARP_CV_EXACT_BILINEAR) return cv_blend(top, bot, (sfx >> 5) & 31, (sfy >> 5) & 31);
    // entry = (B, B', G, G'), (R, R')
    const uint32_t t00 = (top.x & 0xffu) | ((top.x >> 8) & 0xff00u) | ((top.y & 0xffu) << 16);
    const uint32_t t10 = ((top.x >> 8) & 0xffu) | ((top.x >> 16) & 0xff00u) | ((top.y & 0xff00u) << 8);
    const uint32_t t01 = (bot.x & 0xffu) | ((bot.x >> 8) & 0xff00u) | ((bot.y & 0xffu) << 16);
    const uint32_t t11 = ((bot.x >> 8) & 0xffu) | ((bot.x >> 16) & 0xff00u) | ((bot.y & 0xff00u) << 8);
    return warp_blend_taps<MODE>(t00, t10, t01, t11, sfx, sfy);
}

template <int MODE, int BORDER>
__global__ void __launch_bounds__(WT_THREADS)
k_bgr_warp_cv_tiled(const uint8_t* __restrict__ src_base, int64_t src_stride, int64_t src_bs, int w, int h,
                    const int32_t* __restrict__ slots, const VsWarpCoef* __restrict__ coefs,
                    uint8_t* __restrict__ dst_base, int64_t dst_stride, int64_t dst_bs, int dw, int dh,
                    int dst_x0, int dst_y0, int src_al4, int dst_al8)
{
    constexpr int P = WarpGrid<MODE>::P;
    constexpr double SCALE = (double)(1 << P);
    constexpr int ROUND = 1 << (P - WarpGrid<MODE>::W - 1);
    extern __shared__ __align__(16) uint32_t wt_smem[];
    // staged entries as two planes (SX: .x words, SY: .y words): a lane's four entries are then 16
    // contiguous bytes per plane, so the staging stores of a warp are conflict-free
    uint32_t* const SX = wt_smem;
    uint32_t* const SY = wt_smem + WT_SRC_ENTRIES;
    uint32_t* const O = wt_smem + WT_SRC_ENTRIES * 2;
    __shared__ int2 sXY0[WT_H];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * WT_W, oy0 = blockIdx.y * WT_H;
    const int tw = min(WT_W, dw - ox0), th = min(WT_H, dh - oy0);
    const int slot = slots ? slots[b] : b;
    const uint8_t* src = src_base + (size_t)slot * src_bs;
    uint8_t* dst = dst_base + (size_t)b * dst_bs;
    const VsWarpCoef cf = coefs[b];

    // column terms of this thread, row terms of the tile (cv::warpAffine's adelta/bdelta, X0/Y0)
    const int xcol = ox0 + min(tid, tw - 1) + dst_x0;
    const int adelta = __double2int_rn(cf.i00 * (double)xcol * SCALE);
    const int bdelta = __double2int_rn(cf.i10 * (double)xcol * SCALE);
    if (tid < WT_H) {
        const int y = oy0 + min(tid, th - 1) + dst_y0;
        sXY0[tid] = make_int2(__double2int_rn((cf.i01 * (double)y + cf.i02) * SCALE) + ROUND,
                              __double2int_rn((cf.i11 * (double)y + cf.i12) * SCALE) + ROUND);
    }
    __syncthreads();

    // source bounding box from the four corners (each term is monotone in its variable): the
    // column terms of the first and last column are those of threads 0 and tw-1
    const int xr = ox0 + tw - 1 + dst_x0;
    const int aL = __shfl_sync(0xffffffffu, adelta, 0), bL = __shfl_sync(0xffffffffu, bdelta, 0);
    const int aL0 = warp == 0 ? aL : __double2int_rn(cf.i00 * (double)(ox0 + dst_x0) * SCALE);
    const int bL0 = warp == 0 ? bL : __double2int_rn(cf.i10 * (double)(ox0 + dst_x0) * SCALE);
    const int aR = __double2int_rn(cf.i00 * (double)xr * SCALE), bR = __double2int_rn(cf.i10 * (double)xr * SCALE);
    const int2 xyT = sXY0[0], xyB = sXY0[th - 1];
    const int sxmin = (min(xyT.x, xyB.x) + min(aL0, aR)) >> P, sxmax = (max(xyT.x, xyB.x) + max(aL0, aR)) >> P;
    const int symin = (min(xyT.y, xyB.y) + min(bL0, bR)) >> P, symax = (max(xyT.y, xyB.y) + max(bL0, bR)) >> P;
    const int bx0 = (sxmin >> 2) * 4;                            // 4-pixel (12-byte, 3-word) staging granules
    const int ngran = ((sxmax - bx0) >> 2) + 1;                  // entries bx0 .. sxmax (entry x also carries x+1)
    const int pitch = ngran * 4;                                 // entries per staged row
    const int by0 = symin, nrows = symax + 1 - symin + 1;
    const bool staged = pitch * nrows <= WT_SRC_ENTRIES && nrows <= 4096;

    if (staged) {
        // rows of the box inside the image: by0 + r in [0, h)  <=>  r - rlo < rspan (unsigned)
        const int rlo = -by0;
        const unsigned rspan = (unsigned)h;
        for (int q = lane; q < ngran; q += 32) {
            const int x = bx0 + 4 * q;
            const bool xfast = src_al4 && x >= 0 && x + 5 < w;   // the 4th word ends inside the row
            for (int r0 = warp; r0 < nrows; r0 += WT_WARPS * WT_ROWS_PER_WARP) {
                uint32_t wq[WT_ROWS_PER_WARP][4];
                // one pointer and one shared-memory offset, advanced by 4 rows per step
                const uint8_t* g = src + (ptrdiff_t)(by0 + r0) * src_stride + 3 * x;
                const ptrdiff_t gstep = (ptrdiff_t)WT_WARPS * src_stride;
#pragma unroll
                for (int k = 0; k < WT_ROWS_PER_WARP; k++) {
                    const int r = r0 + k * WT_WARPS;
                    if (xfast && r < nrows && (unsigned)(r - rlo) < rspan) {
                        const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
                        wq[k][0] = __ldg(gw); wq[k][1] = __ldg(gw + 1); wq[k][2] = __ldg(gw + 2); wq[k][3] = __ldg(gw + 3);
                    }
                    g += gstep;
                }
                int off = r0 * pitch + 4 * q;
#pragma unroll
                for (int k = 0; k < WT_ROWS_PER_WARP; k++) {
                    const int r = r0 + k * WT_WARPS;
                    if (r >= nrows) break;
                    uint4 ex, ey;
                    if (xfast && (unsigned)(r - rlo) < rspan) {
                        const uint32_t w0 = wq[k][0], w1 = wq[k][1], w2 = wq[k][2], w3 = wq[k][3];
                        // stream bytes of pixel j start at 3j; entry j = (b[3j], b[3j+3], b[3j+1], b[3j+4]), (b[3j+2], b[3j+5]);
                        // the .y words keep two don't-care upper bytes: IDP.2A.LO only reads the lower two
                        ex = make_uint4(__byte_perm(w0, w1, 0x4130), __byte_perm(w0, w1, 0x7463),
                                        __byte_perm(w1, w2, 0x6352), __byte_perm(w2, w3, 0x5241));
                        ey = make_uint4(__byte_perm(w0, w1, 0x0052), __byte_perm(w1, w2, 0x0041),
                                        __byte_perm(w2, w2, 0x0030), __byte_perm(w2, w3, 0x0063));
                    } else {
                        // border or unaligned source: texel by texel (clamped / zero-filled)
                        const int y = by0 + r;
                        const uint32_t p0 = bgr_texel_word<BORDER>(src, src_stride, w, h, x, y);
                        const uint32_t p1 = bgr_texel_word<BORDER>(src, src_stride, w, h, x + 1, y);
                        const uint32_t p2 = bgr_texel_word<BORDER>(src, src_stride, w, h, x + 2, y);
                        const uint32_t p3 = bgr_texel_word<BORDER>(src, src_stride, w, h, x + 3, y);
                        const uint32_t p4 = bgr_texel_word<BORDER>(src, src_stride, w, h, x + 4, y);
                        const uint2 e0 = wt_entry(p0, p1), e1 = wt_entry(p1, p2), e2 = wt_entry(p2, p3), e3 = wt_entry(p3, p4);
                        ex = make_uint4(e0.x, e1.x, e2.x, e3.x);
                        ey = make_uint4(e0.y, e1.y, e2.y, e3.y);
                    }
                    *reinterpret_cast<uint4*>(SX + off) = ex;
                    *reinterpret_cast<uint4*>(SY + off) = ey;
                    off += WT_WARPS * pitch;
                }
            }
        }
        __syncthreads();
    }

    // every thread runs the loops below (columns beyond the tile edge repeat the last column) so the
    // warp shuffles are always executed by full warps
    const int k4 = tid & 3;
    // lane 4j+k (k<3) assembles packed word k of the 12 bytes of pixels 4j..4j+3 from its own
    // BGRX word and its right neighbour's: B0G0R0B1 | G1R1B2G2 | R2B3G3R3
    const uint32_t sel = k4 == 0 ? 0x4210u : (k4 == 1 ? 0x5421u : 0x6542u);
    uint32_t* const Orow = O + 3 * (tid >> 2) + k4;
    const bool keep = k4 < 3 && tid < WT_W;
    if (staged) {
        const int sorg = -by0 * pitch - bx0;
#pragma unroll 4
        for (int r = 0; r < WT_H; r++) {
            if (r >= th) break;
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + adelta, sfy = xy0.y + bdelta;
            const int e = sorg + (sfy >> P) * pitch + (sfx >> P);
            const uint32_t px = wt_blend<MODE>(make_uint2(SX[e], SY[e]), make_uint2(SX[e + pitch], SY[e + pitch]), sfx, sfy);
            const uint32_t nx = __shfl_down_sync(0xffffffffu, px, 1);
            if (keep) Orow[r * WT_OUT_ROW_WORDS] = __byte_perm(px, nx, sel);
        }
    } else {
        for (int r = 0; r < th; r++) {
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + adelta, sfy = xy0.y + bdelta;
            const int sx = sfx >> P, sy = sfy >> P;
            const uint32_t t00 = bgr_texel_word<BORDER>(src, src_stride, w, h, sx, sy);
            const uint32_t t10 = bgr_texel_word<BORDER>(src, src_stride, w, h, sx + 1, sy);
            const uint32_t t01 = bgr_texel_word<BORDER>(src, src_stride, w, h, sx, sy + 1);
            const uint32_t t11 = bgr_texel_word<BORDER>(src, src_stride, w, h, sx + 1, sy + 1);
            const uint32_t px = warp_blend_taps<MODE>(t00, t10, t01, t11, sfx, sfy);
            const uint32_t nx = __shfl_down_sync(0xffffffffu, px, 1);
            if (keep) Orow[r * WT_OUT_ROW_WORDS] = __byte_perm(px, nx, sel);
        }
    }
    __syncthreads();

    // write the tile: 8-byte vectors when the destination rows allow it.  A full row is 45 of them:
    // warp w writes rows w, w+4, ... with lanes 0..31 and then lanes 0..12
    const int row_bytes = tw * 3;
    uint8_t* const drow0 = dst + (size_t)oy0 * dst_stride + (size_t)ox0 * 3;
    if (dst_al8 && tw == WT_W) {
        constexpr int VPR = WT_W * 3 / 8;
        uint2* d = reinterpret_cast<uint2*>(drow0 + (size_t)warp * dst_stride) + lane;
        const uint2* o = reinterpret_cast<const uint2*>(O + warp * WT_OUT_ROW_WORDS) + lane;
        for (int r = warp; r < th; r += WT_WARPS) {
            d[0] = o[0];
            if (lane < VPR - 32) d[32] = o[32];
            d = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(d) + (size_t)WT_WARPS * dst_stride);
            o += WT_WARPS * WT_OUT_ROW_WORDS / 2;
        }
    } else {
        const uint8_t* Ob = reinterpret_cast<const uint8_t*>(O);
        for (int i = tid; i < th * row_bytes; i += WT_THREADS) {
            const int r = i / row_bytes, c = i - r * row_bytes;
            drow0[(size_t)r * dst_stride + c] = Ob[r * (WT_OUT_ROW_WORDS * 4) + c];
        }
    }
}

// ------------------------------------------------------------------ BGR warp, cv-exact, row groups on the raw box
// Same bits as the tiled kernel above, about half its instructions and a third of its shared-memory
// traffic: no expanded copy of the source is built.  A thread owns four consecutive output pixels
// (12 output bytes = three whole words) of one row.  For the near-identity similarity transforms of a
// stabiliser those four pixels almost always read four consecutive source pixels of ONE source row pair
// ("regular group"): then the 15 source bytes of a row are five aligned shared-memory words of the
// TMA-fetched raw box, brought to byte alignment by four funnel shifts, and the paired taps
// (B,B',G,G') / (R,R') of all four pixels are eight PRMTs with fixed selectors.  A warp walks down six
// consecutive rows, and the bottom taps of one row are normally the top taps of the next: they stay in
// registers.  Weights are scaled by 64 so that every blended channel lands in byte 2 of its accumulator
// ((64 S + 32768) >> 16 == (S + 512) >> 10) and the 12 output bytes are assembled by nine PRMTs, stored as
// three words (lane stride 3 words: conflict-free); the finished 128 x 24 tile leaves through one TMA
// store.  Groups whose pixels do not share one source row or consecutive source columns, groups with a
// pixel at fx = fy = 0 (scaled top-left weight 65536: not a 16-bit value) and groups cut by the right image
// edge are pushed to a list and redone pixel by pixel from the same raw box after the main loop — same
// arithmetic, nothing approximated.  The fixed-point column and row terms (cv::warpAffine's adelta/bdelta,
// X0/Y0: the only f64 work) and each tile's source box come from a small table kernel that runs once per
// launch instead of once per tile.
constexpr int WG_W = 128, WG_H = 24, WG_THREADS = 128, WG_WARPS = 4, WG_ROWS_PER_WARP = WG_H / WG_WARPS;
constexpr int WG_BOX_WORDS = VS_WARP_ROWS_BOX_WORDS, WG_BOX_ROWS = VS_WARP_ROWS_BOX_ROWS;   // 160 pixels x 28 rows
constexpr int WG_BOX_PIXELS = WG_BOX_WORDS * 4 / 3;
constexpr int WG_RAW_PITCH = WG_BOX_WORDS * 4;                       // bytes per box row
constexpr int WG_RAW_BYTES = WG_RAW_PITCH * WG_BOX_ROWS;             // 13440: the TMA transaction size
constexpr int WG_OUT_OFF = (WG_RAW_BYTES + 64 + 127) / 128 * 128;    // raw box + over-read pad of the last lane's window
constexpr int WG_OUT_ROW_WORDS = WG_W * 3 / 4;                       // 96 words = 384 bytes per output row
constexpr int WG_LIST_OFF = WG_OUT_OFF + WG_OUT_ROW_WORDS * 4 * WG_H;
constexpr int WG_SMEM_BYTES = WG_LIST_OFF + WG_H * 32 * 2;           // one list entry per group at most
static_assert(WG_BOX_PIXELS == 160 && WG_H % WG_WARPS == 0, "box / tile geometry");
static_assert(WG_W == VS_WARP_ROWS_TILE_W && WG_H == VS_WARP_ROWS_TILE_H, "vs_warp_rows_tab_ints sizes the tables by tile");

// Per launch and image b (vs_warp_rows_tab_ints int32 each): AD[dwp], BD[dwp] (column terms), XY0[dhp] (row terms,
// rounding offset included), TILE[tiles_y][tiles_x] = {first TMA word, first source row, first source pixel, box fits}.
template <int MODE>
__global__ void __launch_bounds__(256)
k_warp_tables(const VsWarpCoef* __restrict__ coefs, int dw, int dh, int dwp, int dhp, int per, int dst_x0, int dst_y0,
              int32_t* __restrict__ tab)
{
    constexpr int P = WarpGrid<MODE>::P, W = WarpGrid<MODE>::W;
    constexpr double SCALE = (double)(1 << P);
    constexpr int ROUND = 1 << (P - W - 1);
    // taps reach LO pixels before and HI pixels after the integer position (2 x 2 bilinear, 4 x 4 Lanczos-2)
    constexpr int LO = MODE == VS_WARP_LANCZOS2 ? 1 : 0, HI = MODE == VS_WARP_LANCZOS2 ? 2 : 1;
    constexpr int BOX_ROWS = MODE == VS_WARP_LANCZOS2 ? VS_WARP_LZ_BOX_ROWS : WG_BOX_ROWS;
    const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    const VsWarpCoef cf = coefs[b];
    int32_t* const t = tab + (size_t)b * per;
    auto colx = [&](int x) { return __double2int_rn(cf.i00 * (double)(x + dst_x0) * SCALE); };
    auto coly = [&](int x) { return __double2int_rn(cf.i10 * (double)(x + dst_x0) * SCALE); };
    auto rowx = [&](int y) { return __double2int_rn((cf.i01 * (double)(y + dst_y0) + cf.i02) * SCALE) + ROUND; };
    auto rowy = [&](int y) { return __double2int_rn((cf.i11 * (double)(y + dst_y0) + cf.i12) * SCALE) + ROUND; };
    const int tiles_x = dwp / WG_W, tiles_y = dhp / WG_H;
    if (i < dwp) {
        t[i] = colx(i);
        t[dwp + i] = coly(i);
    } else if (i < dwp + dhp) {
        reinterpret_cast<int2*>(t + 2 * dwp)[i - dwp] = make_int2(rowx(i - dwp), rowy(i - dwp));
    } else if (i < dwp + dhp + tiles_x * tiles_y) {
        // source bounding box of the tile from its four corners (column and row terms are monotone)
        const int q = i - dwp - dhp, ty = q / tiles_x, tx = q - ty * tiles_x;
        const int xl = tx * WG_W, xr = min(xl + WG_W, dw) - 1, yt = ty * WG_H, yb = min(yt + WG_H, dh) - 1;
        const int aL = colx(xl), aR = colx(xr), bL = coly(xl), bR = coly(xr);
        const int XT = rowx(yt), XB = rowx(yb), YT = rowy(yt), YB = rowy(yb);
        const int sxmin = (min(XT, XB) + min(aL, aR)) >> P, sxmax = (max(XT, XB) + max(aL, aR)) >> P;
        const int symin = (min(YT, YB) + min(bL, bR)) >> P, symax = (max(YT, YB) + max(bL, bR)) >> P;
        const int bx0 = ((sxmin - LO) >> 4) * 16;                 // a TMA box starts 16-byte aligned: 16 pixels = 48 bytes
        const int by0 = symin - LO;
        // pixels bx0 .. sxmax + HI and rows by0 .. symax + HI must be inside the box
        const bool fits = sxmax + HI - bx0 < WG_BOX_PIXELS && symax + HI - by0 < BOX_ROWS &&
                          sxmin > -(1 << 14) && sxmax < (1 << 14) && symin > -(1 << 14) && symax < (1 << 14);
        // bit 1: fx == fy == 0 at all four corners of the tile (then, for a near-identity transform, almost everywhere in it)
        const int fr = ((XT + aL) | (XT + aR) | (XB + aL) | (XB + aR) | (YT + bL) | (YT + bR) | (YB + bL) | (YB + bR)) & (((1 << W) - 1) << (P - W));
        reinterpret_cast<int4*>(t + 2 * dwp + 2 * dhp)[q] = make_int4((bx0 >> 4) * 12, by0, bx0, (fits ? 1 : 0) | (fr == 0 ? 2 : 0));
    }
}

// packed 16-bit weight pairs of one pixel, summing to 65536: wt = w00 | w10 << 16, wb = w01 | w11 << 16 — the cv grid's
// (32 - fx | fx)(32 - fy | fy) scaled by 64, the fine grid's (256 - fx | fx)(256 - fy | fy).  w00 = 65536 when
// fx == fy == 0: then wt == 0x10000 exactly, and the caller takes the pixel-by-pixel path (on the cv grid bit 16 is never
// set otherwise, the upper half being a multiple of 64; on the fine grid it can be, so the test is an equality there).
// Written so that most of the work is multiply-adds: the shift / logic pipe is the busy one in this kernel.
template <int MODE>
__device__ __forceinline__ void wg_weights(int sfx, int sfy, uint32_t& wt, uint32_t& wb)
{
    if (MODE == VS_WARP_CV_EXACT_BILINEAR) {
        const uint32_t gx = (uint32_t)sfx & 0x3e0u;                        // 32 fx
        const uint32_t hp2 = gx * 131070u + 2048u;                         // 64 (32 - fx) | 64 fx << 16
        const uint32_t hp64 = gx * (131070u * 32u) + 65536u;               // 32 hp2 as a multiply-add of its own: wt = 32 hp2 - fy hp2
        uint32_t t, fy;                                                    // then is one subtraction, not (32 - fy) on the logic pipe + a multiply
        asm("shl.b32 %0, %1, 22;" : "=r"(t) : "r"(sfy));                  // two shifts, not shift + mask: the left one can be a multiply
        asm("shr.u32 %0, %1, 27;" : "=r"(fy) : "r"(t));
        wb = fy * hp2;
        wt = hp64 - wb;
    } else {
        // 8-bit fractions: (256 - fx | fx << 16) times (256 - fy | fy); the products are 16-bit values except
        // (256 - 0)(256 - 0) = 65536, which makes wt == 0x10000 exactly (and only then)
        uint32_t t, fx, fy;
        asm("shl.b32 %0, %1, 16;" : "=r"(t) : "r"(sfx));
        asm("shr.u32 %0, %1, 24;" : "=r"(fx) : "r"(t));
        asm("shl.b32 %0, %1, 16;" : "=r"(t) : "r"(sfy));
        asm("shr.u32 %0, %1, 24;" : "=r"(fy) : "r"(t));
        const uint32_t hp = fx * 65535u + 256u;                            // 256 - fx | fx << 16
        const uint32_t hp256 = fx * (65535u * 256u) + 65536u;              // 256 hp
        wb = fy * hp;
        wt = hp256 - wb;
    }
}

__device__ __forceinline__ void wg_blend(uint32_t wt, uint32_t wb, uint32_t tx, uint32_t ty, uint32_t bx, uint32_t by,
                                         uint32_t& b, uint32_t& g, uint32_t& r)
{
    b = __dp2a_lo(wt, tx, 32768u); g = __dp2a_hi(wt, tx, 32768u); r = __dp2a_lo(wt, ty, 32768u);
    b = __dp2a_lo(wb, bx, b); g = __dp2a_hi(wb, bx, g); r = __dp2a_lo(wb, by, r);
}

// byte 2 of four accumulators -> one output word
__device__ __forceinline__ uint32_t wg_pack(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
}

template <int MODE>
__global__ void __launch_bounds__(WG_THREADS, 8)
k_bgr_warp_cv_rows(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map,
                   const uint8_t* __restrict__ src_base, int64_t src_stride, int64_t src_bs, int w, int h,
                   const int32_t* __restrict__ slots, const int32_t* __restrict__ tab, int dwp, int dhp, int per,
                   uint8_t* __restrict__ dst_base, int64_t dst_stride, int64_t dst_bs, int dw, int dh, int dst_tma, int dst_al16)
{
    constexpr int P = WarpGrid<MODE>::P;
    extern __shared__ __align__(128) uint32_t wg_smem[];
    uint32_t* const RAW = wg_smem;                                 // [WG_BOX_ROWS][WG_BOX_WORDS]
    uint32_t* const O = wg_smem + WG_OUT_OFF / 4;                  // [WG_H][WG_OUT_ROW_WORDS]
    uint16_t* const LIST = reinterpret_cast<uint16_t*>(wg_smem + WG_LIST_OFF / 4);
    __shared__ int2 sXY0[WG_H];                                    // row terms relative to the box origin
    __shared__ int sCount;
    __shared__ __align__(8) unsigned long long tma_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * WG_W, oy0 = blockIdx.y * WG_H;
    const int tw = min(WG_W, dw - ox0), th = min(WG_H, dh - oy0);
    const int32_t* const T = tab + (size_t)b * per;
    const int32_t* const AD = T + ox0;
    const int32_t* const BD = AD + dwp;
    const int2* const XY = reinterpret_cast<const int2*>(T + 2 * dwp) + oy0;
    const int4 tile = __ldg(reinterpret_cast<const int4*>(T + 2 * dwp + 2 * dhp) + blockIdx.y * gridDim.x + blockIdx.x);
    const bool staged = (tile.w & 1) != 0;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
    if (tid == 0) {
        sCount = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (staged) {
            const int slot = slots ? __ldg(slots + b) : b;
            const uint32_t dstsm = (uint32_t)__cvta_generic_to_shared(RAW);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WG_RAW_BYTES) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(dstsm), "l"(reinterpret_cast<uint64_t>(&src_map)), "r"(tile.x), "r"(tile.y), "r"(slot), "r"(bar)
                : "memory");
        }
    }
    const int orgx = staged ? tile.z << P : 0, orgy = staged ? tile.y << P : 0;
    if (tid < WG_H) {
        const int2 xy = __ldg(XY + min(tid, th - 1));
        sXY0[tid] = make_int2(xy.x - orgx, xy.y - orgy);
    }
    const int4 ad4 = __ldg(reinterpret_cast<const int4*>(AD) + lane);
    const int4 bd4 = __ldg(reinterpret_cast<const int4*>(BD) + lane);
    __syncthreads();

    uint8_t* const Ob = reinterpret_cast<uint8_t*>(O);
    if (staged) {
        // column terms with the pixel's offset inside the group taken out: a regular group has one integer part
        const int a0 = ad4.x, a1 = ad4.y - (1 << P), a2 = ad4.z - (2 << P), a3 = ad4.w - (3 << P);
        const bool whole = 4 * lane + 3 < tw;
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar) : "memory");
            if (!done && ++spins > (1u << 24)) __trap();   // a lost transaction must fail loudly, not hang the GPU
        }
        auto rows = [&](auto fix_tag) {
            constexpr bool FIX = decltype(fix_tag)::value;
            // A warp walks down six consecutive output rows.  From one row to the next a regular group normally moves
            // down by exactly one source row at the same source column: its top taps are then the previous row's bottom
            // taps, already aligned and paired in registers.  The test is warp-uniform (every regular lane must agree).
            uint32_t pbyte = 0xffffffffu;                                 // raw-box byte offset of the paired bottom taps held in p*
            uint32_t px0 = 0, px1 = 0, px2 = 0, px3 = 0, py0 = 0, py1 = 0, py2 = 0, py3 = 0;
    #pragma unroll
            for (int k = 0; k < WG_ROWS_PER_WARP; k++) {
                const int r = warp * WG_ROWS_PER_WARP + k;
                if (r >= th) break;
                const int2 xy0 = sXY0[r];
                const int x0 = xy0.x + a0, x1 = xy0.x + a1, x2 = xy0.x + a2, x3 = xy0.x + a3;
                const int y0 = xy0.y + bd4.x, y1 = xy0.y + bd4.y, y2 = xy0.y + bd4.z, y3 = xy0.y + bd4.w;
                uint32_t wt0, wb0, wt1, wb1, wt2, wb2, wt3, wb3;
                wg_weights<MODE>(x0, y0, wt0, wb0); wg_weights<MODE>(x1, y1, wt1, wb1);
                wg_weights<MODE>(x2, y2, wt2, wb2); wg_weights<MODE>(x3, y3, wt3, wb3);
                // one source row pair (the row term is monotone in x: the ends decide), consecutive source columns,
                // no pixel at fx == fy == 0
                const uint32_t spread = (uint32_t)((x1 ^ x0) | (x2 ^ x0) | (x3 ^ x0) | (y3 ^ y0));
                // FIX (tiles whose corners all have fx == fy == 0: whole-pixel shifts, a camera standing still): 65535 in place
                // of the unrepresentable 65536 gives the same byte 2 (65535 p + 32768 = 65536 p + (32768 - p)), two more
                // instructions per pixel; elsewhere such a pixel is one in 1024 and its group goes to the list
                if (FIX) {
                    if (wt0 == 0x10000u) wt0 = 0xffffu;
                    if (wt1 == 0x10000u) wt1 = 0xffffu;
                    if (wt2 == 0x10000u) wt2 = 0xffffu;
                    if (wt3 == 0x10000u) wt3 = 0xffffu;
                }
                uint32_t ovf = 0u;
                if (!FIX) {
                    if (MODE == VS_WARP_CV_EXACT_BILINEAR) ovf = (wt0 | wt1 | wt2 | wt3) & 0x10000u;
                    else ovf = (wt0 == 0x10000u) | (wt1 == 0x10000u) | (wt2 == 0x10000u) | (wt3 == 0x10000u);
                }
                const bool regular = whole && spread < (1u << P) && ovf == 0u;
                const uint32_t byte = (uint32_t)(y0 >> P) * (uint32_t)WG_RAW_PITCH + (uint32_t)(x0 >> P) * 3u;
                const bool reuse = __all_sync(0xffffffffu, !regular || byte == pbyte);
                if (regular) {
                    const uint32_t* const p = RAW + (byte >> 2);
                    const uint32_t sh = byte << 3;                        // funnel shifts use the low five bits: 8 (byte & 3)
                    uint32_t tx0, tx1, tx2, tx3, ty0, ty1, ty2, ty3;
                    if (reuse) {
                        tx0 = px0; tx1 = px1; tx2 = px2; tx3 = px3; ty0 = py0; ty1 = py1; ty2 = py2; ty3 = py3;
                    } else {
                        const uint32_t t0 = p[0], t1 = p[1], t2 = p[2], t3 = p[3], t4 = p[4];
                        const uint32_t u0 = __funnelshift_r(t0, t1, sh), u1 = __funnelshift_r(t1, t2, sh),
                                       u2 = __funnelshift_r(t2, t3, sh), u3 = __funnelshift_r(t3, t4, sh);
                        // stream bytes of pixel j start at 3j: taps (s[3j], s[3j+3], s[3j+1], s[3j+4]) and (s[3j+2], s[3j+5])
                        tx0 = __byte_perm(u0, u1, 0x4130); tx1 = __byte_perm(u0, u1, 0x7463);
                        tx2 = __byte_perm(u1, u2, 0x6352); tx3 = __byte_perm(u2, u3, 0x5241);
                        ty0 = __byte_perm(u0, u1, 0x0052); ty1 = __byte_perm(u1, u2, 0x0041);
                        ty2 = __byte_perm(u2, u2, 0x0030); ty3 = __byte_perm(u2, u3, 0x0063);
                    }
                    const uint32_t c0 = p[WG_BOX_WORDS], c1 = p[WG_BOX_WORDS + 1], c2 = p[WG_BOX_WORDS + 2],
                                   c3 = p[WG_BOX_WORDS + 3], c4 = p[WG_BOX_WORDS + 4];
                    const uint32_t v0 = __funnelshift_r(c0, c1, sh), v1 = __funnelshift_r(c1, c2, sh),
                                   v2 = __funnelshift_r(c2, c3, sh), v3 = __funnelshift_r(c3, c4, sh);
                    px0 = __byte_perm(v0, v1, 0x4130); px1 = __byte_perm(v0, v1, 0x7463);
                    px2 = __byte_perm(v1, v2, 0x6352); px3 = __byte_perm(v2, v3, 0x5241);
                    py0 = __byte_perm(v0, v1, 0x0052); py1 = __byte_perm(v1, v2, 0x0041);
                    py2 = __byte_perm(v2, v2, 0x0030); py3 = __byte_perm(v2, v3, 0x0063);
                    pbyte = byte + WG_RAW_PITCH;
                    uint32_t b0, g0, r0, b1, g1, r1, b2, g2, r2, b3, g3, r3;
                    wg_blend(wt0, wb0, tx0, ty0, px0, py0, b0, g0, r0);
                    wg_blend(wt1, wb1, tx1, ty1, px1, py1, b1, g1, r1);
                    wg_blend(wt2, wb2, tx2, ty2, px2, py2, b2, g2, r2);
                    wg_blend(wt3, wb3, tx3, ty3, px3, py3, b3, g3, r3);
                    uint32_t* const o = O + r * WG_OUT_ROW_WORDS + 3 * lane;
                    o[0] = wg_pack(b0, g0, r0, b1);
                    o[1] = wg_pack(g1, r1, b2, g2);
                    o[2] = wg_pack(r2, b3, g3, r3);
                } else {
                    // (the held taps are dead: defining them here lets the register moves of the merge land on this rare path)
                    pbyte = 0xffffffffu;
                    px0 = px1 = px2 = px3 = py0 = py1 = py2 = py3 = 0u;
                    if (4 * lane < tw) LIST[atomicAdd(&sCount, 1)] = (uint16_t)(r * 32 + lane);
                }
            }
        };
        if (tile.w & 2) rows(std::true_type{}); else rows(std::false_type{});
        __syncthreads();
        // irregular groups, pixel by pixel from the raw box (every tap of the tile is inside it)
        const int nfix = sCount * 4;
        const uint8_t* const Rb = reinterpret_cast<const uint8_t*>(RAW);
        for (int i = tid; i < nfix; i += WG_THREADS) {
            const int e = LIST[i >> 2], r = e >> 5, x = 4 * (e & 31) + (i & 3);
            if (x >= tw) continue;
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + __ldg(AD + x), sfy = xy0.y + __ldg(BD + x);
            const uint8_t* const q = Rb + (sfy >> P) * WG_RAW_PITCH + (sfx >> P) * 3;
            const uint32_t t00 = q[0] | (q[1] << 8) | (q[2] << 16), t10 = q[3] | (q[4] << 8) | (q[5] << 16);
            const uint32_t t01 = q[WG_RAW_PITCH] | (q[WG_RAW_PITCH + 1] << 8) | (q[WG_RAW_PITCH + 2] << 16),
                           t11 = q[WG_RAW_PITCH + 3] | (q[WG_RAW_PITCH + 4] << 8) | (q[WG_RAW_PITCH + 5] << 16);
            const uint32_t px = warp_blend_taps<MODE>(t00, t10, t01, t11, sfx, sfy);
            uint8_t* const o = Ob + r * (WG_OUT_ROW_WORDS * 4) + 3 * x;
            o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
        }
    } else {
        // the box does not fit (large rotation or scale): every pixel straight from global memory
        const uint8_t* const src = src_base + (size_t)(slots ? __ldg(slots + b) : b) * src_bs;
        for (int i = tid; i < tw * th; i += WG_THREADS) {
            const int r = i / tw, x = i - r * tw;
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + __ldg(AD + x), sfy = xy0.y + __ldg(BD + x);
            const int sx = sfx >> P, sy = sfy >> P;
            const uint32_t t00 = bgr_texel_word<VS_BORDER_CONSTANT0>(src, src_stride, w, h, sx, sy);
            const uint32_t t10 = bgr_texel_word<VS_BORDER_CONSTANT0>(src, src_stride, w, h, sx + 1, sy);
            const uint32_t t01 = bgr_texel_word<VS_BORDER_CONSTANT0>(src, src_stride, w, h, sx, sy + 1);
            const uint32_t t11 = bgr_texel_word<VS_BORDER_CONSTANT0>(src, src_stride, w, h, sx + 1, sy + 1);
            const uint32_t px = warp_blend_taps<MODE>(t00, t10, t01, t11, sfx, sfy);
            uint8_t* const o = Ob + r * (WG_OUT_ROW_WORDS * 4) + 3 * x;
            o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
        }
    }

    if (dst_tma) {
        // the tile leaves through one TMA store (rows / words beyond the image are clipped by the tensor map)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t srcsm = (uint32_t)__cvta_generic_to_shared(O);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(reinterpret_cast<uint64_t>(&dst_map)), "r"(blockIdx.x * WG_OUT_ROW_WORDS), "r"(oy0), "r"(b), "r"(srcsm)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    // destinations a tensor map cannot describe: 16-byte vectors when possible (24 per full row), else bytes
    uint8_t* const drow0 = dst_base + (size_t)b * dst_bs + (size_t)oy0 * dst_stride + (size_t)ox0 * 3;
    if (dst_al16 && tw == WG_W) {
        constexpr int VPR = WG_W * 3 / 16;
        for (int i = tid; i < th * VPR; i += WG_THREADS) {
            const int r = i / VPR, c = i - r * VPR;
            reinterpret_cast<uint4*>(drow0 + (size_t)r * dst_stride)[c] = reinterpret_cast<const uint4*>(O + r * WG_OUT_ROW_WORDS)[c];
        }
    } else {
        const int row_bytes = tw * 3;
        for (int i = tid; i < th * row_bytes; i += WG_THREADS) {
            const int r = i / row_bytes, c = i - r * row_bytes;
            drow0[(size_t)r * dst_stride + c] = Ob[r * (WG_OUT_ROW_WORDS * 4) + c];
        }
    }
}

// ------------------------------------------------------------------ BGR warp, Lanczos-2, row groups on the raw box
// 4 x 4 Lanczos-2 on the 16.16 grid, exact integer arithmetic (the mode has no counterpart upstream; the oracle defines it
// the same way): weights of the 64 fractions q/64 in Q11 from a table; per pixel and channel the vertical sums of the four
// source columns (Q11), then their horizontal sum (Q22: at most 1.4e9 in magnitude, no intermediate shift), rounded,
// clamped.  Same skeleton as k_bgr_warp_cv_rows: 128 x 24 output tile, raw source box by one TMA load (160 pixels x 32
// rows), a thread owns four consecutive pixels of a row, a warp walks down six rows, the tile leaves by one TMA store.
// A row of the window is 7 aligned shared-memory words + 6 funnel shifts; two rows are interleaved by PRMT into byte
// pairs and a vertical sum is two IDP.2A (signed 16-bit weight pairs x unsigned bytes).  The four window rows live in a
// ring of two interleaved pairs: from one output row to the next a regular group normally moves down by exactly one
// source row at the same column, so only the new bottom row is loaded and merged (by PRMT) into the pair that held the
// old top row, and the weight pairs are taken from the table rotated by the ring's phase (warp-uniform).  A "regular" group — consecutive source columns, one source
// row quadruple — whose pixels also share one vertical fraction shares its vertical sums: the 7 source pixels x 3
// channels it touches are 21 sums for 4 output pixels instead of 48.  Everything else (groups that are not regular,
// tiles whose box does not fit) goes pixel by pixel through lz_pixel.
constexpr int LZ_BOX_ROWS = VS_WARP_LZ_BOX_ROWS;
constexpr int LZ_RAW_BYTES = WG_RAW_PITCH * LZ_BOX_ROWS;             // 15360: the TMA transaction size
constexpr int LZ_OUT_OFF = (LZ_RAW_BYTES + 64 + 127) / 128 * 128;
constexpr int LZ_LIST_OFF = LZ_OUT_OFF + WG_OUT_ROW_WORDS * 4 * WG_H;
constexpr int LZ_TAB_OFF = LZ_LIST_OFF + WG_H * 32 * 2;
constexpr int LZ_SMEM_BYTES = LZ_TAB_OFF + 64 * 16 + 4 * 64 * 8;

struct LzTables {
    int32_t wx[64][4];       // Q11 weights of taps -1, 0, 1, 2 for fraction q / 64
};

__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// shared-memory loads by 32-bit shared address (the weight tables: keeps the lookups off generic addressing)
__device__ __forceinline__ uint2 lds_u2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// one pixel, any position, through a byte fetcher tap(x, y, c): returns B | G << 8 | R << 16
template <typename TAP>
__device__ __forceinline__ uint32_t lz_pixel(const int32_t (*wxt)[4], int sfx, int sfy, TAP tap)
{
    const int ix = sfx >> 16, iy = sfy >> 16;
    const int32_t* wx = wxt[(sfx >> 10) & 63];
    const int32_t* wy = wxt[(sfy >> 10) & 63];
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int hsum = 1 << 21;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int v = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) v += wy[r] * (int)tap(ix - 1 + t, iy - 1 + r, c);
            hsum += wx[t] * v;
        }
        out |= (uint32_t)__vimin_s32_relu(hsum >> 22, 255) << (8 * c);
    }
    return out;
}

__global__ void __launch_bounds__(WG_THREADS, 6)
k_bgr_warp_lz_rows(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map,
                   const __grid_constant__ LzTables tables,
                   const uint8_t* __restrict__ src_base, int64_t src_stride, int64_t src_bs, int w, int h,
                   const int32_t* __restrict__ slots, const int32_t* __restrict__ tab, int dwp, int dhp, int per,
                   uint8_t* __restrict__ dst_base, int64_t dst_stride, int64_t dst_bs, int dw, int dh, int dst_tma, int dst_al16,
                   int border_flags)
{
    constexpr int P = 16;
    extern __shared__ __align__(128) uint32_t lz_smem[];
    uint32_t* const RAW = lz_smem;                                 // [LZ_BOX_ROWS][WG_BOX_WORDS]
    uint32_t* const O = lz_smem + LZ_OUT_OFF / 4;                  // [WG_H][WG_OUT_ROW_WORDS]
    uint16_t* const LIST = reinterpret_cast<uint16_t*>(lz_smem + LZ_LIST_OFF / 4);
    int32_t (*const WX)[4] = reinterpret_cast<int32_t (*)[4]>(lz_smem + LZ_TAB_OFF / 4);
    // vertical weight pairs for the four phases of the row ring: slot s holds window row (s - phase) & 3
    uint2* const WY = reinterpret_cast<uint2*>(lz_smem + LZ_TAB_OFF / 4 + 64 * 4);      // [4][64]
    __shared__ int2 sXY0[WG_H];                                    // row terms relative to the box origin
    __shared__ int sCount;
    __shared__ __align__(8) unsigned long long tma_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * WG_W, oy0 = blockIdx.y * WG_H;
    const int tw = min(WG_W, dw - ox0), th = min(WG_H, dh - oy0);
    const int32_t* const T = tab + (size_t)b * per;
    const int32_t* const AD = T + ox0;
    const int32_t* const BD = AD + dwp;
    const int2* const XY = reinterpret_cast<const int2*>(T + 2 * dwp) + oy0;
    const int4 tile = __ldg(reinterpret_cast<const int4*>(T + 2 * dwp + 2 * dhp) + blockIdx.y * gridDim.x + blockIdx.x);
    const bool staged = (tile.w & 1) != 0 && border_flags == VS_BORDER_CONSTANT0;   // bit 1 of border_flags: no source map
    const int border = border_flags & 1;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
    if (tid == 0) {
        sCount = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (staged) {
            const int slot = slots ? __ldg(slots + b) : b;
            const uint32_t dstsm = (uint32_t)__cvta_generic_to_shared(RAW);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)LZ_RAW_BYTES) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(dstsm), "l"(reinterpret_cast<uint64_t>(&src_map)), "r"(tile.x), "r"(tile.y), "r"(slot), "r"(bar)
                : "memory");
        }
    }
    if (tid < 64) {
        // one 16-byte constant load per fraction: the horizontal weights as they are, the vertical ones as 16-bit pairs in
        // the four rotations of the row ring
        const int4 wq = *reinterpret_cast<const int4*>(tables.wx[tid]);
        *reinterpret_cast<int4*>(WX[tid]) = wq;
        const uint32_t w0 = (uint32_t)wq.x & 0xffffu, w1 = (uint32_t)wq.y & 0xffffu, w2 = (uint32_t)wq.z & 0xffffu, w3 = (uint32_t)wq.w & 0xffffu;
        WY[0 * 64 + tid] = make_uint2(w0 | w1 << 16, w2 | w3 << 16);      // slots hold rows 0 1 2 3
        WY[1 * 64 + tid] = make_uint2(w3 | w0 << 16, w1 | w2 << 16);      //                 3 0 1 2
        WY[2 * 64 + tid] = make_uint2(w2 | w3 << 16, w0 | w1 << 16);      //                 2 3 0 1
        WY[3 * 64 + tid] = make_uint2(w1 | w2 << 16, w3 | w0 << 16);      //                 1 2 3 0
    }
    const uint32_t wx_s = (uint32_t)__cvta_generic_to_shared(WX), wy_s = (uint32_t)__cvta_generic_to_shared(WY);
    const int orgx = staged ? tile.z << P : 0, orgy = staged ? tile.y << P : 0;
    if (tid < WG_H) {
        const int2 xy = __ldg(XY + min(tid, th - 1));
        sXY0[tid] = make_int2(xy.x - orgx, xy.y - orgy);
    }
    const int4 ad4 = __ldg(reinterpret_cast<const int4*>(AD) + lane);
    const int4 bd4 = __ldg(reinterpret_cast<const int4*>(BD) + lane);
    __syncthreads();

    uint8_t* const Ob = reinterpret_cast<uint8_t*>(O);
    if (staged) {
        const int a0 = ad4.x, a1 = ad4.y - (1 << P), a2 = ad4.z - (2 << P), a3 = ad4.w - (3 << P);
        const bool whole = 4 * lane + 3 < tw;
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar) : "memory");
            if (!done && ++spins > (1u << 24)) __trap();   // a lost transaction must fail loudly, not hang the GPU
        }
        // the ring: slots 0, 1 interleaved in a01 / b01 and slots 2, 3 in a23 / b23 (21 stream bytes of a row = 6 words)
        uint32_t a01[6], b01[6], a23[6], b23[6];
        uint32_t prev_byte = 0xffffffffu;         // window offset of this thread's previous row when that row was regular
#pragma unroll 1
        for (int k = 0; k < WG_ROWS_PER_WARP; k++) {
            const int phase = k & 3;              // slot of the window's top row (the same for the whole warp)
            const int r = warp * WG_ROWS_PER_WARP + k;
            if (r >= th) break;
            const int2 xy0 = sXY0[r];
            const int x0 = xy0.x + a0, x1 = xy0.x + a1, x2 = xy0.x + a2, x3 = xy0.x + a3;
            const int y0 = xy0.y + bd4.x, y1 = xy0.y + bd4.y, y2 = xy0.y + bd4.z, y3 = xy0.y + bd4.w;
            // consecutive source columns and one source row quadruple for the four pixels; when they also share one vertical
            // fraction (pure translations, and wherever the rotation moves y by less than 1/64 px over the group) the
            // vertical sums are shared by the group, else every pixel forms its own from the same interleaved rows.  Which
            // of the two forms runs is decided per warp, so no lane waits for code it does not need.
            const uint32_t spread = (uint32_t)((x1 ^ x0) | (x2 ^ x0) | (x3 ^ x0) | (y1 ^ y0) | (y2 ^ y0) | (y3 ^ y0)) >> P;
            const uint32_t yfrac = (uint32_t)((y1 ^ y0) | (y2 ^ y0) | (y3 ^ y0)) >> 10;
            const bool regular = whole && spread == 0u;
            const bool shared_wy = __all_sync(0xffffffffu, !regular || yfrac == 0u);
            if (regular) {
                const uint32_t byte = (uint32_t)((y0 >> P) - 1) * (uint32_t)WG_RAW_PITCH + (uint32_t)((x0 >> P) - 1) * 3u;
                const uint32_t* const p = RAW + (byte >> 2);
                const uint32_t sh = byte << 3;                        // funnel shifts use the low five bits: 8 (byte & 3)
                // window row rr (0 = top) as six byte-aligned words
                auto load_row = [&](int rr, uint32_t* u) {
#pragma unroll
                    for (int q = 0; q < 6; q++)
                        u[q] = __funnelshift_r(p[rr * WG_BOX_WORDS + q], p[rr * WG_BOX_WORDS + q + 1], sh);
                };
                if (k > 0 && byte == prev_byte + (uint32_t)WG_RAW_PITCH) {
                    // one source row down at the same column: the new bottom row replaces the old top row in its pair
                    uint32_t u[6];
                    load_row(3, u);
                    const int slot = (phase + 3) & 3;
                    const uint32_t selA = (slot & 1) ? 0x5240u : 0x3514u, selB = (slot & 1) ? 0x7260u : 0x3716u;   // odd / even bytes of the pair
                    if (slot < 2) {
#pragma unroll
                        for (int q = 0; q < 6; q++) { a01[q] = __byte_perm(a01[q], u[q], selA); b01[q] = __byte_perm(b01[q], u[q], selB); }
                    } else {
#pragma unroll
                        for (int q = 0; q < 6; q++) { a23[q] = __byte_perm(a23[q], u[q], selA); b23[q] = __byte_perm(b23[q], u[q], selB); }
                    }
                } else {
#pragma unroll
                    for (int pr = 0; pr < 2; pr++) {                  // the rows of slots 2 pr and 2 pr + 1
                        uint32_t ue[6], uo[6];
                        load_row((2 * pr - phase) & 3, ue);
                        load_row((2 * pr + 1 - phase) & 3, uo);
#pragma unroll
                        for (int q = 0; q < 6; q++) {
                            (pr ? a23 : a01)[q] = __byte_perm(ue[q], uo[q], 0x5140);
                            (pr ? b23 : b01)[q] = __byte_perm(ue[q], uo[q], 0x7362);
                        }
                    }
                }
                prev_byte = byte;
                // vertical sum of stream byte i (Q11) under the weight pairs wy of this phase
                auto vsum = [&](int i, const uint2 wy) -> int {
                    const int q = i >> 2;
                    switch (i & 3) {
                    case 0: return dp2a_lo_su(wy.y, a23[q], dp2a_lo_su(wy.x, a01[q], 0));
                    case 1: return dp2a_hi_su(wy.y, a23[q], dp2a_hi_su(wy.x, a01[q], 0));
                    case 2: return dp2a_lo_su(wy.y, b23[q], dp2a_lo_su(wy.x, b01[q], 0));
                    default: return dp2a_hi_su(wy.y, b23[q], dp2a_hi_su(wy.x, b01[q], 0));
                    }
                };
                const uint32_t wyp = wy_s + (uint32_t)phase * 512u;
                const int xs[4] = {x0, x1, x2, x3};
                const int ys[4] = {y0, y1, y2, y3};
                uint32_t px[4];
                if (shared_wy) {
                    const uint2 wy = lds_u2(wyp + ((((uint32_t)y0 >> 10) & 63u) << 3));
                    int vs[21];
#pragma unroll
                    for (int i = 0; i < 21; i++) vs[i] = vsum(i, wy);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int4 wx = lds_i4(wx_s + ((((uint32_t)xs[j] >> 10) & 63u) << 4));
                        uint32_t o = 0;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const int hsum = wx.x * vs[3 * j + c] + wx.y * vs[3 * j + 3 + c] + wx.z * vs[3 * j + 6 + c] +
                                             wx.w * vs[3 * j + 9 + c] + (1 << 21);
                            o |= (uint32_t)__vimin_s32_relu(hsum >> 22, 255) << (8 * c);
                        }
                        px[j] = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint2 wy = lds_u2(wyp + ((((uint32_t)ys[j] >> 10) & 63u) << 3));
                        const int4 wx = lds_i4(wx_s + ((((uint32_t)xs[j] >> 10) & 63u) << 4));
                        uint32_t o = 0;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const int hsum = wx.x * vsum(3 * j + c, wy) + wx.y * vsum(3 * j + 3 + c, wy) + wx.z * vsum(3 * j + 6 + c, wy) +
                                             wx.w * vsum(3 * j + 9 + c, wy) + (1 << 21);
                            o |= (uint32_t)__vimin_s32_relu(hsum >> 22, 255) << (8 * c);
                        }
                        px[j] = o;
                    }
                }
                uint32_t* const o = O + r * WG_OUT_ROW_WORDS + 3 * lane;
                o[0] = __byte_perm(px[0], px[1], 0x4210);
                o[1] = __byte_perm(px[1], px[2], 0x5421);
                o[2] = __byte_perm(px[2], px[3], 0x6542);
            } else {
                prev_byte = 0xffffffffu;
                if (4 * lane < tw) LIST[atomicAdd(&sCount, 1)] = (uint16_t)(r * 32 + lane);
            }
        }
        __syncthreads();
        // irregular groups, pixel by pixel from the raw box (every tap of the tile is inside it)
        const int nfix = sCount * 4;
        const uint8_t* const Rb = reinterpret_cast<const uint8_t*>(RAW);
        for (int i = tid; i < nfix; i += WG_THREADS) {
            const int e = LIST[i >> 2], r = e >> 5, x = 4 * (e & 31) + (i & 3);
            if (x >= tw) continue;
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + __ldg(AD + x), sfy = xy0.y + __ldg(BD + x);
            const uint32_t px = lz_pixel(WX, sfx, sfy, [&](int xx, int yy, int c) { return Rb[yy * WG_RAW_PITCH + xx * 3 + c]; });
            uint8_t* const o = Ob + r * (WG_OUT_ROW_WORDS * 4) + 3 * x;
            o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
        }
    } else {
        // the box does not fit (large rotation or scale) or the border repeats the edge: every pixel from global memory
        const uint8_t* const src = src_base + (size_t)(slots ? __ldg(slots + b) : b) * src_bs;
        for (int i = tid; i < tw * th; i += WG_THREADS) {
            const int r = i / tw, x = i - r * tw;
            const int2 xy0 = sXY0[r];
            const int sfx = xy0.x + __ldg(AD + x), sfy = xy0.y + __ldg(BD + x);
            const uint32_t px = lz_pixel(WX, sfx, sfy, [&](int xx, int yy, int c) -> uint32_t {
                if (border == VS_BORDER_REPEAT_EDGE) { xx = vs_clampi(xx, 0, w - 1); yy = vs_clampi(yy, 0, h - 1); }
                else if (xx < 0 || xx >= w || yy < 0 || yy >= h) return 0u;
                return __ldg(src + (size_t)yy * src_stride + 3 * xx + c);
            });
            uint8_t* const o = Ob + r * (WG_OUT_ROW_WORDS * 4) + 3 * x;
            o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
        }
    }

    if (dst_tma) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t srcsm = (uint32_t)__cvta_generic_to_shared(O);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(reinterpret_cast<uint64_t>(&dst_map)), "r"(blockIdx.x * WG_OUT_ROW_WORDS), "r"(oy0), "r"(b), "r"(srcsm)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    uint8_t* const drow0 = dst_base + (size_t)b * dst_bs + (size_t)oy0 * dst_stride + (size_t)ox0 * 3;
    const int row_bytes = tw * 3;
    for (int i = tid; i < th * row_bytes; i += WG_THREADS) {
        const int r = i / row_bytes, c = i - r * row_bytes;
        drow0[(size_t)r * dst_stride + c] = Ob[r * (WG_OUT_ROW_WORDS * 4) + c];
    }
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace

// ================================================================== launchers

int vsk_bgr2gray(vs_ctx* ctx, const VsDevImg& bgr, const VsDevImg& gray)
{
    VS_REQUIRE(ctx, bgr.w == gray.w && bgr.h == gray.h && bgr.batch == gray.batch, "bgr2gray: shape mismatch");
    if (bgr.w <= 0 || bgr.h <= 0) return VS_OK;
    VS_REQUIRE(ctx, bgr.h <= 65535 && bgr.batch <= 65535, "bgr2gray: image too tall / batch too large");
    int vec_ok = aligned_to(bgr.data, 16) && bgr.stride % 16 == 0 && bgr.batch_stride % 16 == 0 &&
                 aligned_to(gray.data, 16) && gray.stride % 16 == 0 && gray.batch_stride % 16 == 0;
    dim3 block(128), grid(vs_cdiv(vs_cdiv(bgr.w, 16), 128), bgr.h, bgr.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_BGR2GRAY);
    k_bgr2gray<<<grid, block, 0, ctx->stream>>>((const uint8_t*)bgr.data, bgr.stride, bgr.batch_stride,
                                                (uint8_t*)gray.data, gray.stride, gray.batch_stride,
                                                bgr.w, bgr.h, vec_ok);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

// BGR -> gray level 0 + level 1 in one pass when the geometry allows the fused kernel; returns VS_OK and sets *fused
int vsk_ingest_bgr_gray_l1(vs_ctx* ctx, const VsDevImg& bgr, const VsDevImg& g0, const VsDevImg& g1, bool* fused)
{
    *fused = false;
    VS_REQUIRE(ctx, bgr.w == g0.w && bgr.h == g0.h && bgr.batch == g0.batch && bgr.batch == g1.batch, "ingest: shape mismatch");
    if (bgr.w <= 0 || bgr.h <= 0) return VS_OK;
    const bool ok = bgr.w % 16 == 0 && bgr.h >= 2 && g1.w == bgr.w / 2 && g1.h == bgr.h / 2 &&
                    aligned_to(bgr.data, 16) && bgr.stride % 16 == 0 && bgr.batch_stride % 16 == 0 &&
                    aligned_to(g0.data, 16) && g0.stride % 16 == 0 && g0.batch_stride % 16 == 0 &&
                    aligned_to(g1.data, 8) && g1.stride % 8 == 0 && g1.batch_stride % 8 == 0 &&
                    bgr.batch <= 65535 && vs_cdiv(g1.h, IG_L1_ROWS * 4) <= 65535;
    if (!ok) return VS_OK;
    dim3 block(32, 4), grid(vs_cdiv(bgr.w / 16, 32), vs_cdiv(vs_cdiv(g1.h, IG_L1_ROWS), 4), bgr.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_INGEST);
    k_ingest_bgr_gray_l1<<<grid, block, 0, ctx->stream>>>((const uint8_t*)bgr.data, bgr.stride, bgr.batch_stride,
                                                          (uint8_t*)g0.data, g0.stride, g0.batch_stride,
                                                          (uint8_t*)g1.data, g1.stride, g1.batch_stride, bgr.w, bgr.h, g1.w, g1.h);
    VS_LAUNCH_CHECK(ctx);
    *fused = true;
    return VS_OK;
}

int vsk_pyr_down(vs_ctx* ctx, const VsDevImg& in, const VsDevImg& out)
{
    VS_REQUIRE(ctx, in.batch == out.batch, "pyr_down: batch mismatch");
    VS_REQUIRE(ctx, in.w > 0 && in.h > 0, "pyr_down: empty input");
    if (out.w <= 0 || out.h <= 0) return VS_OK;
    const int in_al8 = aligned_to(in.data, 8) && in.stride % 8 == 0 && in.batch_stride % 8 == 0;
    const int out_al4 = aligned_to(out.data, 4) && out.stride % 4 == 0 && out.batch_stride % 4 == 0;
    VS_REQUIRE(ctx, out.batch <= 65535 && vs_cdiv(out.h, PD_ROWS_DEFAULT) <= 65535, "pyr_down: grid too large");
    // the wide kernel (8 outputs per thread, 4 rows) whenever alignment and sizes allow it, else the 4-column kernel that
    // takes any alignment and size (the other shapes that were measured are in profiles/r01_v5_kernel_bench.jsonl)
    const bool ok16 = aligned_to(in.data, 16) && in.stride % 16 == 0 && in.batch_stride % 16 == 0 &&
                      aligned_to(out.data, 8) && out.stride % 8 == 0 && out.batch_stride % 8 == 0 &&
                      out.w % 8 == 0 && 2 * out.w <= in.w;
    VS_LAUNCH_BEGIN(ctx, VSK_PYR_DOWN);
    if (ok16) {
        dim3 wblock(32, 4), wgrid(vs_cdiv(out.w / 8, 32), vs_cdiv(out.h, 4 * 4), out.batch);
        k_pyr_down_wide<4><<<wgrid, wblock, 0, ctx->stream>>>((const uint8_t*)in.data, in.stride, in.batch_stride, in.w, in.h,
                                                              (uint8_t*)out.data, out.stride, out.batch_stride, out.w, out.h);
    } else {
        dim3 block(32, 4), grid(vs_cdiv(vs_cdiv(out.w, 4), 32), vs_cdiv(out.h, 4 * PD_ROWS_DEFAULT), out.batch);
        k_pyr_down<PD_ROWS_DEFAULT, 7><<<grid, block, 0, ctx->stream>>>((const uint8_t*)in.data, in.stride, in.batch_stride, in.w, in.h,
                                                                        (uint8_t*)out.data, out.stride, out.batch_stride, out.w, out.h,
                                                                        in_al8, out_al4);
    }
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_grad_xy(vs_ctx* ctx, const VsDevImg& in, const VsDevImg& gx, const VsDevImg& gy)
{
    VS_REQUIRE(ctx, gx.w == gy.w && gx.h == gy.h && in.batch == gx.batch && in.batch == gy.batch, "grad_xy: shape mismatch");
    VS_REQUIRE(ctx, in.w > 0 && in.h > 0, "grad_xy: empty input");
    if (gx.w <= 0 || gx.h <= 0) return VS_OK;
    VS_REQUIRE(ctx, gx.h <= 65535 && gx.batch <= 65535, "grad_xy: grid too large");
    int vec_ok = aligned_to(gx.data, 16) && gx.stride % 4 == 0 && gx.batch_stride % 4 == 0 &&
                 aligned_to(gy.data, 16) && gy.stride % 4 == 0 && gy.batch_stride % 4 == 0;
    dim3 block(128), grid(vs_cdiv(vs_cdiv(gx.w, 4), 128), gx.h, gx.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_GRAD_XY);
    k_grad_xy<<<grid, block, 0, ctx->stream>>>((const uint8_t*)in.data, in.stride, in.batch_stride, in.w, in.h,
                                               (float*)gx.data, gx.stride, gx.batch_stride,
                                               (float*)gy.data, gy.stride, gy.batch_stride, gx.w, gx.h, vec_ok);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_image_warp(vs_ctx* ctx, const VsDevImg& in, const float* d_params4, const VsDevImg& out)
{
    VS_REQUIRE(ctx, in.batch == out.batch, "image_warp: batch mismatch");
    VS_REQUIRE(ctx, in.w > 0 && in.h > 0, "image_warp: empty input");
    if (out.w <= 0 || out.h <= 0) return VS_OK;
    VS_REQUIRE(ctx, out.h <= 65535 && out.batch <= 65535, "image_warp: grid too large");
    int vec_ok = aligned_to(out.data, 16) && out.stride % 4 == 0 && out.batch_stride % 4 == 0;
    dim3 block(128), grid(vs_cdiv(vs_cdiv(out.w, 4), 128), out.h, out.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_IMAGE_WARP);
    k_image_warp<<<grid, block, 0, ctx->stream>>>((const uint8_t*)in.data, in.stride, in.batch_stride, in.w, in.h,
                                                  d_params4, (float*)out.data, out.stride, out.batch_stride,
                                                  out.w, out.h, vec_ok);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

// cv::warpAffine inverts the forward matrix in f64 (no WARP_INVERSE_MAP at imgproc.cpp:472)
void vs_warp_coef_from_forward(const double* M, VsWarpCoef* o)
{
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1.0 / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    o->i00 = A11; o->i01 = M[1] * (-D); o->i10 = M[3] * (-D); o->i11 = A22;
    o->i02 = -o->i00 * M[2] - o->i01 * M[5];
    o->i12 = -o->i10 * M[2] - o->i11 * M[5];
}

void vs_forward_matrix_from_transform(const double* T, int cols, int rows, double* M)
{
    double cx = (cols - 1) * 0.5, cy = (rows - 1) * 0.5;
    double tx_ul = T[2] - T[0] * cx + T[1] * cy;
    double ty_ul = T[3] - T[1] * cx - T[0] * cy;
    M[0] = 1.0 + T[0]; M[1] = -T[1]; M[2] = tx_ul;
    M[3] = T[1]; M[4] = 1.0 + T[0]; M[5] = ty_ul;
}

template <int MODE, int BORDER>
static int launch_bgr_warp_tiled(vs_ctx* ctx, const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef,
                                 const VsDevImg& dst, int dst_x0, int dst_y0)
{
    // per device, so set on every launch (multi-GPU processes drive several devices)
    VS_CUDA(ctx, cudaFuncSetAttribute(k_bgr_warp_cv_tiled<MODE, BORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES));
    const int src_al4 = aligned_to(src.data, 4) && src.stride % 4 == 0 && src.batch_stride % 4 == 0;
    const int dst_al8 = aligned_to(dst.data, 8) && dst.stride % 8 == 0 && dst.batch_stride % 8 == 0;
    dim3 tgrid(vs_cdiv(dst.w, WT_W), vs_cdiv(dst.h, WT_H), dst.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_BGR_WARP);
    k_bgr_warp_cv_tiled<MODE, BORDER><<<tgrid, WT_THREADS, WT_SMEM_BYTES, ctx->stream>>>(
        (const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h, d_slots, d_coef,
        (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_x0, dst_y0, src_al4, dst_al8);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

// the fine grid keeps 16 fractional bits in an int32: positions must stay below 2^14 pixels
static bool warp_fits_fine_grid(const VsDevImg& src, const VsDevImg& dst, int dst_x0, int dst_y0)
{
    return src.w <= 16384 && src.h <= 16384 && dst.w + dst_x0 <= 16384 && dst.h + dst_y0 <= 16384 && dst_x0 >= 0 && dst_y0 >= 0;
}

int vsk_bgr_warp_slots(vs_ctx* ctx, const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef,
                       const VsDevImg& dst, int dst_x0, int dst_y0, int mode, int border)
{
    VS_REQUIRE(ctx, src.w > 0 && src.h > 0, "bgr_warp: empty source");
    VS_REQUIRE(ctx, mode >= 0 && mode <= 2 && (border == 0 || border == 1), "bgr_warp: bad mode/border");
    if (dst.w <= 0 || dst.h <= 0 || dst.batch <= 0) return VS_OK;
    VS_REQUIRE(ctx, dst.h <= 65535 && dst.batch <= 65535, "bgr_warp: grid too large");
    if (mode == VS_WARP_CV_EXACT_BILINEAR || mode == VS_WARP_FLOAT_BILINEAR) {
        VS_REQUIRE(ctx, vs_cdiv(dst.h, WT_H) <= 65535, "bgr_warp: grid too large");
        if (mode == VS_WARP_FLOAT_BILINEAR)
            VS_REQUIRE(ctx, warp_fits_fine_grid(src, dst, dst_x0, dst_y0), "bgr_warp: the 16.16 grid holds frames up to 16384 pixels");
        if (mode == VS_WARP_CV_EXACT_BILINEAR)
            return border == VS_BORDER_REPEAT_EDGE
                       ? launch_bgr_warp_tiled<VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_REPEAT_EDGE>(ctx, src, d_slots, d_coef, dst, dst_x0, dst_y0)
                       : launch_bgr_warp_tiled<VS_WARP_CV_EXACT_BILINEAR, VS_BORDER_CONSTANT0>(ctx, src, d_slots, d_coef, dst, dst_x0, dst_y0);
        return border == VS_BORDER_REPEAT_EDGE
                   ? launch_bgr_warp_tiled<VS_WARP_FLOAT_BILINEAR, VS_BORDER_REPEAT_EDGE>(ctx, src, d_slots, d_coef, dst, dst_x0, dst_y0)
                   : launch_bgr_warp_tiled<VS_WARP_FLOAT_BILINEAR, VS_BORDER_CONSTANT0>(ctx, src, d_slots, d_coef, dst, dst_x0, dst_y0);
    }
    return vs_set_error(ctx, VS_ERR_INVALID, "bgr_warp: the Lanczos-2 mode runs through vsk_bgr_warp_lz");
}

// ------------------------------------------------------------------ planar warp (NV12 frames), cv-exact
// cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of a 1-channel (the Y plane) or 2-channel (the interleaved UV plane of
// an NV12 frame) u8 image: the same fixed-point grid as the BGR warp (AB_BITS 10, INTER_BITS 5, 15-bit weights).  A lane
// owns four output pixels of a row (one 4- or 8-byte store), a CTA a 128 x 32 tile: the column terms are computed once
// per lane, the row terms once per CTA.  Regular groups read two aligned word windows through L1, the others byte taps.
constexpr int PW_THREADS = 256, PW_W = 128, PW_H = 32;
static_assert(PW_THREADS == 256 && PW_W == 128 && PW_H == 32, "k_plane_warp_cv maps 8 warps onto 4 x 2 blocks of 32 x 16 pixels");

template <int CH>
__global__ void __launch_bounds__(PW_THREADS)
k_plane_warp_cv(const uint8_t* __restrict__ src_base, int64_t src_stride, int64_t src_bs, int w, int h,
                const int32_t* __restrict__ slots, const VsWarpCoef* __restrict__ coefs,
                uint8_t* __restrict__ dst_base, int64_t dst_stride, int64_t dst_bs, int dw, int dh,
                int dst_x0, int dst_y0, int src_al, int src_al4, int dst_al)
{
    constexpr int P = 10;
    constexpr double SCALE = 1024.0;
    constexpr int ROUND = 16;
    __shared__ int2 sXY0[PW_H];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * PW_W, oy0 = blockIdx.y * PW_H;
    const int slot = slots ? slots[b] : b;
    const uint8_t* src = src_base + (size_t)slot * src_bs;
    uint8_t* dst = dst_base + (size_t)b * dst_bs;
    const VsWarpCoef cf = coefs[b];
    if (tid < PW_H) {
        const double y = (double)(oy0 + tid + dst_y0);
        sXY0[tid] = make_int2(__double2int_rn((cf.i01 * y + cf.i02) * SCALE) + ROUND,
                              __double2int_rn((cf.i11 * y + cf.i12) * SCALE) + ROUND);
    }
    // a warp covers 32 pixels x 4 rows per step (eight lanes of four pixels per row), a quarter of the tile's width over 16
    // of its rows: a similarity's irregular groups (where the source column or row index steps by 0 or 2) lie on two nearly
    // straight lines about 1 / |A| and 1 / |B| pixels apart, and a compact footprint meets them four times less often
    // than a 128-pixel row segment would
    const int xo = ox0 + 32 * (warp & 3) + 4 * (lane & 7);
    int ad[4], bd[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double x = (double)(xo + j + dst_x0);
        ad[j] = __double2int_rn(cf.i00 * x * SCALE);
        bd[j] = __double2int_rn(cf.i10 * x * SCALE);
    }
    __syncthreads();
    if (xo >= dw) return;
    const int npx = min(4, dw - xo);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = 16 * (warp >> 2) + 4 * k + (lane >> 3);
        const int yo = oy0 + r;
        if (yo >= dh) continue;
        const int2 xy0 = sXY0[r];
        uint32_t px[4][CH];
        int sfx[4], sfy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { sfx[j] = xy0.x + ad[j]; sfy[j] = xy0.y + bd[j]; }
        const int sx0 = sfx[0] >> P, sy0 = sfy[0] >> P;
        // regular group: the four pixels read source columns sx0 .. sx0 + 4 of rows sy0, sy0 + 1, all inside the image (any
        // similarity close to the identity, away from the border): two aligned word windows instead of 16 byte loads, the
        // vertical blend of a column pair as one packed multiply-add, the horizontal one as IDP.2A; fractions stay per pixel
        // (the aligned words that hold the window must end inside the row: the last row of a tightly packed image has
        // nothing behind it)
        bool regular = src_al4 && sx0 >= 0 && sx0 + 4 < w && sy0 >= 0 && sy0 + 1 < h &&
                       (((sx0 * CH) >> 2) + CH + 1) * 4 <= w * CH;
#pragma unroll
        for (int j = 1; j < 4; j++) regular = regular && (sfx[j] >> P) == sx0 + j && (sfy[j] >> P) == sy0;
        if (regular) {
            const uint8_t* row = src + (size_t)sy0 * src_stride;
            const int byte0 = sx0 * CH;
            const uint32_t* wt_ = reinterpret_cast<const uint32_t*>(row) + (byte0 >> 2);
            const uint32_t* wb_ = reinterpret_cast<const uint32_t*>(row + src_stride) + (byte0 >> 2);
            const uint32_t sh = 8u * (uint32_t)(byte0 & 3);
            uint32_t qt[4], qb[4];     // CH 1: [p_j, 0, p_j+1, 0] per pixel; CH 2: [U_j V_j U_j+1 V_j+1]
            if (CH == 1) {
                const uint32_t t0 = __ldg(wt_), t1 = __ldg(wt_ + 1), b0 = __ldg(wb_), b1 = __ldg(wb_ + 1);
                const uint32_t ta = __funnelshift_r(t0, t1, sh), te = (t1 >> sh) & 0xffu;
                const uint32_t ba = __funnelshift_r(b0, b1, sh), be = (b1 >> sh) & 0xffu;
                qt[0] = __byte_perm(ta, te, 0x5150); qt[1] = __byte_perm(ta, te, 0x5251);
                qt[2] = __byte_perm(ta, te, 0x5352); qt[3] = __byte_perm(ta, te, 0x5453);
                qb[0] = __byte_perm(ba, be, 0x5150); qb[1] = __byte_perm(ba, be, 0x5251);
                qb[2] = __byte_perm(ba, be, 0x5352); qb[3] = __byte_perm(ba, be, 0x5453);
            } else {
                const uint32_t t0 = __ldg(wt_), t1 = __ldg(wt_ + 1), t2 = __ldg(wt_ + 2);
                const uint32_t b0 = __ldg(wb_), b1 = __ldg(wb_ + 1), b2 = __ldg(wb_ + 2);
                qt[0] = __funnelshift_r(t0, t1, sh); qt[2] = __funnelshift_r(t1, t2, sh);
                qb[0] = __funnelshift_r(b0, b1, sh); qb[2] = __funnelshift_r(b1, b2, sh);
                const uint32_t te = t2 >> sh, be = b2 >> sh;
                qt[1] = __byte_perm(qt[0], qt[2], 0x5432); qt[3] = __byte_perm(qt[2], te, 0x5432);
                qb[1] = __byte_perm(qb[0], qb[2], 0x5432); qb[3] = __byte_perm(qb[2], be, 0x5432);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t fx = ((uint32_t)sfx[j] >> 5) & 31u, fy = ((uint32_t)sfy[j] >> 5) & 31u;
                const uint32_t wx = fx * 255u + 32u;                     // (32 - fx) | fx << 8
                // (sum w p + 16384) >> 15 with w = 32 wx wy  ==  ((32 - fx) v_j + fx v_j+1 + 512) >> 10, v = (32 - fy) top + fy bottom
                if (CH == 1) {
                    const uint32_t v = (32u - fy) * qt[j] + fy * qb[j];
                    px[j][0] = __dp2a_lo(v, wx, 512u) >> 10;
                } else {
                    const uint32_t vu = (32u - fy) * (qt[j] & 0x00ff00ffu) + fy * (qb[j] & 0x00ff00ffu);
                    const uint32_t vv = (32u - fy) * ((qt[j] >> 8) & 0x00ff00ffu) + fy * ((qb[j] >> 8) & 0x00ff00ffu);
                    px[j][0] = __dp2a_lo(vu, wx, 512u) >> 10;
                    px[j][CH - 1] = __dp2a_lo(vv, wx, 512u) >> 10;
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int sx = sfx[j] >> P, sy = sfy[j] >> P;
            const uint32_t fx = ((uint32_t)sfx[j] >> 5) & 31u, fy = ((uint32_t)sfy[j] >> 5) & 31u;
            const bool x0 = (unsigned)sx < (unsigned)w, x1 = (unsigned)(sx + 1) < (unsigned)w;
            const bool y0 = (unsigned)sy < (unsigned)h, y1 = (unsigned)(sy + 1) < (unsigned)h;
            const uint8_t* p = src + (ptrdiff_t)sy * src_stride + (ptrdiff_t)sx * CH;
            uint32_t t[4] = {0u, 0u, 0u, 0u};        // taps (x, y), (x+1, y), (x, y+1), (x+1, y+1), channels in bytes
            if (CH == 1) {
                if (x0 && y0) t[0] = __ldg(p);
                if (x1 && y0) t[1] = __ldg(p + 1);
                if (x0 && y1) t[2] = __ldg(p + src_stride);
                if (x1 && y1) t[3] = __ldg(p + src_stride + 1);
            } else if (src_al) {
                if (x0 && y0) t[0] = __ldg(reinterpret_cast<const uint16_t*>(p));
                if (x1 && y0) t[1] = __ldg(reinterpret_cast<const uint16_t*>(p + 2));
                if (x0 && y1) t[2] = __ldg(reinterpret_cast<const uint16_t*>(p + src_stride));
                if (x1 && y1) t[3] = __ldg(reinterpret_cast<const uint16_t*>(p + src_stride + 2));
            } else {
                if (x0 && y0) t[0] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8);
                if (x1 && y0) t[1] = (uint32_t)__ldg(p + 2) | ((uint32_t)__ldg(p + 3) << 8);
                if (x0 && y1) t[2] = (uint32_t)__ldg(p + src_stride) | ((uint32_t)__ldg(p + src_stride + 1) << 8);
                if (x1 && y1) t[3] = (uint32_t)__ldg(p + src_stride + 2) | ((uint32_t)__ldg(p + src_stride + 3) << 8);
            }
            // (sum w p + 16384) >> 15 with w = 32 wx wy  ==  (sum wx wy p + 512) >> 10
            const uint32_t w00 = (32u - fx) * (32u - fy), w10 = fx * (32u - fy), w01 = (32u - fx) * fy, w11 = fx * fy;
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const uint32_t v = w00 * ((t[0] >> (8 * c)) & 0xffu) + w10 * ((t[1] >> (8 * c)) & 0xffu) +
                                   w01 * ((t[2] >> (8 * c)) & 0xffu) + w11 * ((t[3] >> (8 * c)) & 0xffu);
                px[j][c] = (v + 512u) >> 10;
            }
        }
        }
        uint8_t* d = dst + (size_t)yo * dst_stride + (size_t)xo * CH;
        if (dst_al && npx == 4) {
            if (CH == 1) {
                *reinterpret_cast<uint32_t*>(d) = px[0][0] | (px[1][0] << 8) | (px[2][0] << 16) | (px[3][0] << 24);
            } else {
                *reinterpret_cast<uint2*>(d) = make_uint2(px[0][0] | (px[0][CH - 1] << 8) | (px[1][0] << 16) | (px[1][CH - 1] << 24),
                                                          px[2][0] | (px[2][CH - 1] << 8) | (px[3][0] << 16) | (px[3][CH - 1] << 24));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j < npx) {
#pragma unroll
                    for (int c = 0; c < CH; c++) d[j * CH + c] = (uint8_t)px[j][c];
                }
        }
    }
}

// One pixel of a one-channel plane from its four byte taps (zero outside the image): the general path.
__device__ __forceinline__ uint32_t plane_pixel_taps(const uint8_t* __restrict__ src, int64_t src_stride, int w, int h, int sfx, int sfy)
{
    const int sx = sfx >> 10, sy = sfy >> 10;
    const uint32_t fx = ((uint32_t)sfx >> 5) & 31u, fy = ((uint32_t)sfy >> 5) & 31u;
    const bool x0 = (unsigned)sx < (unsigned)w, x1 = (unsigned)(sx + 1) < (unsigned)w;
    const bool y0 = (unsigned)sy < (unsigned)h, y1 = (unsigned)(sy + 1) < (unsigned)h;
    const uint8_t* p = src + (ptrdiff_t)sy * src_stride + sx;
    uint32_t t00 = 0u, t10 = 0u, t01 = 0u, t11 = 0u;
    if (x0 && y0) t00 = __ldg(p);
    if (x1 && y0) t10 = __ldg(p + 1);
    if (x0 && y1) t01 = __ldg(p + src_stride);
    if (x1 && y1) t11 = __ldg(p + src_stride + 1);
    const uint32_t v = (32u - fx) * (32u - fy) * t00 + fx * (32u - fy) * t10 + (32u - fx) * fy * t01 + fx * fy * t11;
    return (v + 512u) >> 10;
}

// The Y plane, eight pixels per lane (one 8-byte store).  A warp covers 32 pixels x 8 rows per step (compact footprint, as
// above), a CTA a 128 x 64 tile.  The column terms of a lane are kept relative to its first pixel: with
// r = (X0 + adelta[0]) & 1023 the pixel j of a regular group (source column sx0 + j) has r + off_j in [0, 1024), where
// off_j = adelta[j] - adelta[0] - 1024 j is a lane constant of a few units; the eight values travel as four packed pairs of
// 16-bit halves biased by 1024 (one add per pair, one LOP3 per pair for the test, the horizontal fractions by one shift
// and mask per pair, the IDP.2A weight bytes of two pixels by one multiply-add).  The row terms stay per pixel (their
// column part is monotone, so the eight rows coincide when the first and the last do).  A regular group reads three
// aligned words per source row; pixel j blends its column pair [p_j, p_j+1] vertically as one packed multiply-add
// (32 top + fy (bottom - top) on the packed word: the halves cannot interfere because the true result has none) and
// horizontally by IDP.2A.  Anything else (borders, a skipped source column or row, tile tails) goes pixel by pixel.
constexpr int PW8_H = 64;      // rows of the Y kernel's tile: the column terms of a lane serve four steps of eight rows

template <int DUMMY>
__global__ void __launch_bounds__(PW_THREADS)
k_plane_warp_y8(const uint8_t* __restrict__ src_base, int64_t src_stride, int64_t src_bs, int w, int h,
                const int32_t* __restrict__ slots, const VsWarpCoef* __restrict__ coefs,
                uint8_t* __restrict__ dst_base, int64_t dst_stride, int64_t dst_bs, int dw, int dh,
                int dst_x0, int dst_y0)
{
    constexpr double SCALE = 1024.0;
    constexpr int ROUND = 16;
    __shared__ int2 sXY0[PW8_H];
    __shared__ __align__(16) int2 sCol[PW_W];     // (adelta, bdelta) of the tile's columns: two f64 products per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * PW_W, oy0 = blockIdx.y * PW8_H;
    const int slot = slots ? slots[b] : b;
    const uint8_t* src = src_base + (size_t)slot * src_bs;
    uint8_t* dst = dst_base + (size_t)b * dst_bs;
    const VsWarpCoef cf = coefs[b];
    if (tid < PW8_H) {
        const double y = (double)(oy0 + tid + dst_y0);
        sXY0[tid] = make_int2(__double2int_rn((cf.i01 * y + cf.i02) * SCALE) + ROUND,
                              __double2int_rn((cf.i11 * y + cf.i12) * SCALE) + ROUND);
    }
    if (tid >= PW_THREADS - PW_W) {
        const int c = tid - (PW_THREADS - PW_W);
        const double x = (double)(ox0 + c + dst_x0);
        sCol[c] = make_int2(__double2int_rn(cf.i00 * x * SCALE), __double2int_rn(cf.i10 * x * SCALE));
    }
    __syncthreads();
    const int xo = ox0 + 32 * (warp & 3) + 8 * (lane & 3);
    int ad0, bd[8];
    uint32_t cax[4];            // (off_2k + 1024) | (off_2k+1 + 1024) << 16
    bool lane_ok = true;        // offsets small enough for the packed form
    {
        int ad[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const int4 c2 = *reinterpret_cast<const int4*>(&sCol[xo - ox0 + j]);
            ad[j] = c2.x; bd[j] = c2.y; ad[j + 1] = c2.z; bd[j + 1] = c2.w;
        }
        ad0 = ad[0];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o0 = ad[2 * k] - ad0 - 1024 * (2 * k), o1 = ad[2 * k + 1] - ad0 - 1024 * (2 * k + 1);
            lane_ok = lane_ok && o0 >= -1024 && o0 <= 1024 && o1 >= -1024 && o1 <= 1024;
            cax[k] = (uint32_t)(o0 + 1024) | ((uint32_t)(o1 + 1024) << 16);
        }
    }
    if (xo >= dw) return;
    const bool full = xo + 8 <= dw;
    uint8_t* const dcol = dst + xo;
#pragma unroll
    for (int k = 0; k < PW8_H / 16; k++) {
        const int r = (PW8_H / 2) * (warp >> 2) + 8 * k + (lane >> 2);
        const int yo = oy0 + r;
        if (yo >= dh) continue;
        const int2 xy0 = sXY0[r];
        uint8_t* const d = dcol + (size_t)yo * dst_stride;
        const int s0x = xy0.x + ad0, sx0 = s0x >> 10;
        const int sfy0 = xy0.y + bd[0], sfy7 = xy0.y + bd[7], sy0 = sfy0 >> 10;
        const uint32_t rx2 = (uint32_t)(s0x & 1023) * 0x10001u;
        uint32_t rel[4];
        uint32_t bad = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            rel[q] = rx2 + cax[q];
            bad |= rel[q] ^ 0x04000400u;
        }
        // the three aligned words that hold source columns sx0 .. sx0 + 8 must end inside the row
        const bool regular = full && lane_ok && (bad & 0xfc00fc00u) == 0u && (sfy7 >> 10) == sy0 && sx0 >= 0 && sy0 >= 0 &&
                             sy0 + 1 < h && ((sx0 >> 2) + 3) * 4 <= w;
        if (regular) {
            const uint32_t* wt_ = reinterpret_cast<const uint32_t*>(src + (size_t)sy0 * src_stride) + (sx0 >> 2);
            const uint32_t* wb_ = reinterpret_cast<const uint32_t*>(src + (size_t)(sy0 + 1) * src_stride) + (sx0 >> 2);
            const uint32_t t0 = __ldg(wt_), t1 = __ldg(wt_ + 1), t2 = __ldg(wt_ + 2);
            const uint32_t b0 = __ldg(wb_), b1 = __ldg(wb_ + 1), b2 = __ldg(wb_ + 2);
            const uint32_t sh = 8u * (uint32_t)(sx0 & 3);
            const uint32_t ta = __funnelshift_r(t0, t1, sh), tb = __funnelshift_r(t1, t2, sh), tc = (t2 >> sh) & 0xffu;
            const uint32_t ba = __funnelshift_r(b0, b1, sh), bb = __funnelshift_r(b1, b2, sh), bc = (b2 >> sh) & 0xffu;
            uint32_t out[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // [p_j, 0, p_j+1, 0] of both rows (the second operand supplies the zero bytes, and p_8)
                uint32_t top, bot;
                if (j < 3)       { top = __byte_perm(ta, 0u, 0x4140 + 0x0101 * j); bot = __byte_perm(ba, 0u, 0x4140 + 0x0101 * j); }
                else if (j == 3) { top = __byte_perm(ta, tb & 0xffu, 0x5453); bot = __byte_perm(ba, bb & 0xffu, 0x5453); }
                else if (j < 7)  { top = __byte_perm(tb, 0u, 0x4140 + 0x0101 * (j - 4)); bot = __byte_perm(bb, 0u, 0x4140 + 0x0101 * (j - 4)); }
                else             { top = __byte_perm(tb, tc, 0x5453); bot = __byte_perm(bb, bc, 0x5453); }
                const uint32_t fy = ((uint32_t)(xy0.y + bd[j]) >> 5) & 31u;
                const uint32_t v = (top << 5) + fy * (bot - top);                  // halves: 32 top + fy (bottom - top) <= 8160
                const uint32_t fxp = (rel[j >> 1] >> 5) & 0x001f001fu;             // fx of pixels j & ~1 and j | 1
                const uint32_t wxp = fxp * 255u + 0x00200020u;                     // (32 - fx) | fx << 8 in each half
                // (sum w p + 16384) >> 15 with w = 32 wx wy  ==  ((32 - fx) v_j + fx v_j+1 + 512) >> 10
                out[j] = ((j & 1) ? __dp2a_hi(v, wxp, 512u) : __dp2a_lo(v, wxp, 512u)) >> 10;
            }
            *reinterpret_cast<uint2*>(d) = make_uint2(out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24),
                                                      out[4] | (out[5] << 8) | (out[6] << 16) | (out[7] << 24));
        } else {
            const int npx = min(8, dw - xo);
#pragma unroll 1
            for (int j = 0; j < npx; j++) {
                // the column terms again (a dynamic index into the register arrays would spill them; this path is rare)
                const double x = (double)(xo + j + dst_x0);
                const int a = __double2int_rn(cf.i00 * x * SCALE), bq = __double2int_rn(cf.i10 * x * SCALE);
                d[j] = (uint8_t)plane_pixel_taps(src, src_stride, w, h, xy0.x + a, xy0.y + bq);
            }
        }
    }
}

int vsk_plane_warp_slots(vs_ctx* ctx, const VsDevImg& src, int channels, const int32_t* d_slots, const VsWarpCoef* d_coef,
                         const VsDevImg& dst, int dst_x0, int dst_y0)
{
    VS_REQUIRE(ctx, channels == 1 || channels == 2, "plane_warp: 1 or 2 channels");
    VS_REQUIRE(ctx, src.w > 0 && src.h > 0, "plane_warp: empty source");
    if (dst.w <= 0 || dst.h <= 0 || dst.batch <= 0) return VS_OK;
    VS_REQUIRE(ctx, vs_cdiv(dst.h, PW_H) <= 65535 && dst.batch <= 65535, "plane_warp: grid too large");
    const int src_al = aligned_to(src.data, 2) && src.stride % 2 == 0 && src.batch_stride % 2 == 0;
    const int src_al4 = aligned_to(src.data, 4) && src.stride % 4 == 0 && src.batch_stride % 4 == 0;
    const int va = 4 * channels;
    const int dst_al = aligned_to(dst.data, va) && dst.stride % va == 0 && dst.batch_stride % va == 0;
    dim3 grid(vs_cdiv(dst.w, PW_W), vs_cdiv(dst.h, PW_H), dst.batch);
    VS_LAUNCH_BEGIN(ctx, VSK_BGR_WARP);
    const bool y8 = channels == 1 && src_al4 && aligned_to(dst.data, 8) && dst.stride % 8 == 0 && dst.batch_stride % 8 == 0;
    if (y8)
        k_plane_warp_y8<0><<<dim3(vs_cdiv(dst.w, PW_W), vs_cdiv(dst.h, PW8_H), dst.batch), PW_THREADS, 0, ctx->stream>>>((const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h,
            d_slots, d_coef, (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_x0, dst_y0);
    else if (channels == 1)
        k_plane_warp_cv<1><<<grid, PW_THREADS, 0, ctx->stream>>>((const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h,
            d_slots, d_coef, (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_x0, dst_y0, src_al, src_al4, dst_al);
    else
        k_plane_warp_cv<2><<<grid, PW_THREADS, 0, ctx->stream>>>((const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h,
            d_slots, d_coef, (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_x0, dst_y0, src_al, src_al4, dst_al);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

// Q11 Lanczos-2 weights of the 64 fractions q / 64 (taps at -1, 0, 1, 2): the reference's lanczos2 polynomial
// (generators.cpp:31-47) in double, normalised; the rounding residue goes to the largest weight (rows sum to 2048)
static const LzTables& lz_tables()
{
    static const LzTables t = [] {
        LzTables r;
        for (int q = 0; q < 64; q++) {
            const double f = q / 64.0;
            double wgt[4], sum = 0.0;
            for (int k = 0; k < 4; k++) {
                const double x = (double)(k - 1) - f, x2 = x * x;
                double v = 0.000858519;
                v = -0.0158853 + v * x2;
                v = 0.128693 + v * x2;
                v = -0.583468 + v * x2;
                v = 1.52229 + v * x2;
                v = -2.05238 + v * x2;
                v = 0.999861 + v * x2;
                wgt[k] = fabs(x) >= 2.0 ? 0.0 : v;
                sum += wgt[k];
            }
            int iw[4], isum = 0, big = 0;
            for (int k = 0; k < 4; k++) {
                iw[k] = (int)lrint(wgt[k] / sum * 2048.0);
                isum += iw[k];
                if (iw[k] > iw[big]) big = k;
            }
            iw[big] += 2048 - isum;
            for (int k = 0; k < 4; k++) r.wx[q][k] = iw[k];
        }
        return r;
    }();
    return t;
}

// Lanczos-2 BGR warp.  tensor_map: a CUtensorMap over the source viewed as u32 [image][row][word] with box
// {120, VS_WARP_LZ_BOX_ROWS, 1}, or null (then every tile reads global memory directly: small or unaligned sources).
// d_tab: vs_warp_rows_tab_ints(dst.w, dst.h) int32 per image.
int vsk_bgr_warp_lz(vs_ctx* ctx, const void* tensor_map, const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef,
                    const VsDevImg& dst, int dst_x0, int dst_y0, int border, int32_t* d_tab)
{
    VS_REQUIRE(ctx, d_tab && src.w > 0 && src.h > 0, "bgr_warp_lz: bad source / scratch");
    VS_REQUIRE(ctx, border == VS_BORDER_CONSTANT0 || border == VS_BORDER_REPEAT_EDGE, "bgr_warp_lz: bad border");
    if (dst.w <= 0 || dst.h <= 0 || dst.batch <= 0) return VS_OK;
    VS_REQUIRE(ctx, vs_cdiv(dst.h, WG_H) <= 65535 && dst.batch <= 65535, "bgr_warp_lz: grid too large");
    VS_REQUIRE(ctx, warp_fits_fine_grid(src, dst, dst_x0, dst_y0), "bgr_warp: the 16.16 grid holds frames up to 16384 pixels");
    const int dwp = vs_cdiv(dst.w, WG_W) * WG_W, dhp = vs_cdiv(dst.h, WG_H) * WG_H;
    const int ntiles = (dwp / WG_W) * (dhp / WG_H);
    const int per = (int)vs_warp_rows_tab_ints(dst.w, dst.h);
    const int dst_al16 = aligned_to(dst.data, 16) && dst.stride % 16 == 0 && dst.batch_stride % 16 == 0;
    CUtensorMap dst_map, src_map;
    memset(&dst_map, 0, sizeof(dst_map));
    memset(&src_map, 0, sizeof(src_map));
    int dst_tma = 0;
    if (dst_al16 && dst.w % 4 == 0 && vs_tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t)(dst.w * 3 / 4), (cuuint64_t)dst.h, (cuuint64_t)dst.batch};
        const cuuint64_t strides[2] = {(cuuint64_t)dst.stride, (cuuint64_t)(dst.batch > 1 ? dst.batch_stride : dst.stride * dst.h)};
        const cuuint32_t box[3] = {(cuuint32_t)WG_OUT_ROW_WORDS, (cuuint32_t)WG_H, 1}, estr[3] = {1, 1, 1};
        CUresult r = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(vs_tensor_map_encoder())(
            &dst_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dst.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        dst_tma = (r == CUDA_SUCCESS);
    }
    // without a source map every tile takes the direct path: the kernel is told through the border argument's high bit
    const bool have_map = tensor_map != nullptr;
    if (have_map) src_map = *reinterpret_cast<const CUtensorMap*>(tensor_map);
    VS_CUDA(ctx, cudaFuncSetAttribute(k_bgr_warp_lz_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ_SMEM_BYTES));
    VS_LAUNCH_BEGIN(ctx, VSK_BGR_WARP);
    k_warp_tables<VS_WARP_LANCZOS2><<<dim3(vs_cdiv(dwp + dhp + ntiles, 256), dst.batch), 256, 0, ctx->stream>>>(
        d_coef, dst.w, dst.h, dwp, dhp, per, dst_x0, dst_y0, d_tab);
    ctx->launches++;
    dim3 tgrid(dwp / WG_W, dhp / WG_H, dst.batch);
    k_bgr_warp_lz_rows<<<tgrid, WG_THREADS, LZ_SMEM_BYTES, ctx->stream>>>(
        src_map, dst_map, lz_tables(), (const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h, d_slots, d_tab, dwp, dhp,
        per, (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_tma, dst_al16, have_map ? border : (border | 2));
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

// cuTensorMapEncodeTiled lives in the driver library; resolve it through the runtime so that
// libvstab.so keeps linking against cudart only
void* vs_tensor_map_encoder()
{
    // a function-local static is initialised once even when several host threads (one VideoStabilizer each) arrive together
    static void* const fn = []() -> void* {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return p;
        cudaGetLastError();
        return nullptr;
    }();
    return fn;
}

// clip-resident sources, row-group kernel: table kernel + warp kernel on the same stream
template <int MODE>
static void launch_warp_rows(cudaStream_t s, dim3 tab_grid, dim3 tgrid, const CUtensorMap& src_map, const CUtensorMap& dst_map,
                             const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef, const VsDevImg& dst,
                             int dst_x0, int dst_y0, int32_t* d_tab, int dwp, int dhp, int per, int dst_tma, int dst_al16)
{
    cudaFuncSetAttribute(k_bgr_warp_cv_rows<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    k_warp_tables<MODE><<<tab_grid, 256, 0, s>>>(d_coef, dst.w, dst.h, dwp, dhp, per, dst_x0, dst_y0, d_tab);
    k_bgr_warp_cv_rows<MODE><<<tgrid, WG_THREADS, WG_SMEM_BYTES, s>>>(
        src_map, dst_map, (const uint8_t*)src.data, src.stride, src.batch_stride, src.w, src.h, d_slots, d_tab, dwp, dhp, per,
        (uint8_t*)dst.data, dst.stride, dst.batch_stride, dst.w, dst.h, dst_tma, dst_al16);
}

int vsk_bgr_warp_slots_rows(vs_ctx* ctx, const void* tensor_map, const VsDevImg& src, const int32_t* d_slots,
                            const VsWarpCoef* d_coef, const VsDevImg& dst, int dst_x0, int dst_y0, int32_t* d_tab, int mode)
{
    VS_REQUIRE(ctx, mode == VS_WARP_CV_EXACT_BILINEAR || mode == VS_WARP_FLOAT_BILINEAR, "bgr_warp_rows: bilinear modes only");
    if (mode == VS_WARP_FLOAT_BILINEAR)
        VS_REQUIRE(ctx, warp_fits_fine_grid(src, dst, dst_x0, dst_y0), "bgr_warp: the 16.16 grid holds frames up to 16384 pixels");
    VS_REQUIRE(ctx, tensor_map && d_tab && src.w > 0 && src.h > 0, "bgr_warp_rows: bad source");
    if (dst.w <= 0 || dst.h <= 0 || dst.batch <= 0) return VS_OK;
    VS_REQUIRE(ctx, vs_cdiv(dst.h, WG_H) <= 65535 && dst.batch <= 65535, "bgr_warp_rows: grid too large");
    const int dwp = vs_cdiv(dst.w, WG_W) * WG_W, dhp = vs_cdiv(dst.h, WG_H) * WG_H;
    const int ntiles = (dwp / WG_W) * (dhp / WG_H);
    const int per = (int)vs_warp_rows_tab_ints(dst.w, dst.h);
    const int dst_al16 = aligned_to(dst.data, 16) && dst.stride % 16 == 0 && dst.batch_stride % 16 == 0;
    // the destination as a u32 [image][row][word] tensor for the TMA store: needs whole words per row and 16-byte strides
    CUtensorMap dst_map;
    memset(&dst_map, 0, sizeof(dst_map));
    int dst_tma = 0;
    if (dst_al16 && dst.w % 4 == 0 && vs_tensor_map_encoder()) {
        const cuuint64_t dims[3] = {(cuuint64_t)(dst.w * 3 / 4), (cuuint64_t)dst.h, (cuuint64_t)dst.batch};
        const cuuint64_t strides[2] = {(cuuint64_t)dst.stride, (cuuint64_t)(dst.batch > 1 ? dst.batch_stride : dst.stride * dst.h)};
        const cuuint32_t box[3] = {(cuuint32_t)WG_OUT_ROW_WORDS, (cuuint32_t)WG_H, 1}, estr[3] = {1, 1, 1};
        CUresult r = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(vs_tensor_map_encoder())(
            &dst_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dst.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        dst_tma = (r == CUDA_SUCCESS);
    }
    VS_LAUNCH_BEGIN(ctx, VSK_BGR_WARP);
    const dim3 tab_grid(vs_cdiv(dwp + dhp + ntiles, 256), dst.batch), tgrid(dwp / WG_W, dhp / WG_H, dst.batch);
    const CUtensorMap& smap = *reinterpret_cast<const CUtensorMap*>(tensor_map);
    if (mode == VS_WARP_CV_EXACT_BILINEAR)
        launch_warp_rows<VS_WARP_CV_EXACT_BILINEAR>(ctx->stream, tab_grid, tgrid, smap, dst_map, src, d_slots, d_coef, dst, dst_x0,
                                                    dst_y0, d_tab, dwp, dhp, per, dst_tma, dst_al16);
    else
        launch_warp_rows<VS_WARP_FLOAT_BILINEAR>(ctx->stream, tab_grid, tgrid, smap, dst_map, src, d_slots, d_coef, dst, dst_x0,
                                                 dst_y0, d_tab, dwp, dhp, per, dst_tma, dst_al16);
    ctx->launches++;            // the table kernel
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

bool vs_warp_rows_usable(const VsDevImg& src, int mode, int border)
{
    if (mode == VS_WARP_LANCZOS2)
        return border == VS_BORDER_CONSTANT0 && vs_tensor_map_encoder() != nullptr && aligned_to(src.data, 16) &&
               src.stride % 16 == 0 && (src.batch <= 1 || src.batch_stride % 16 == 0) && src.w % 4 == 0 &&
               src.w * 3 / 4 >= VS_WARP_ROWS_BOX_WORDS && src.h >= VS_WARP_LZ_BOX_ROWS;
    // whole words per row (elements past 3w/4 words are out of bounds = zero-filled = BORDER_CONSTANT(0)), a box that fits
    return (mode == VS_WARP_CV_EXACT_BILINEAR || mode == VS_WARP_FLOAT_BILINEAR) && border == VS_BORDER_CONSTANT0 &&
           vs_tensor_map_encoder() != nullptr &&
           aligned_to(src.data, 16) && src.stride % 16 == 0 && (src.batch <= 1 || src.batch_stride % 16 == 0) &&
           src.w % 4 == 0 && src.w * 3 / 4 >= VS_WARP_ROWS_BOX_WORDS && src.h >= VS_WARP_ROWS_BOX_ROWS;
}

int vsk_bgr_warp(vs_ctx* ctx, const VsDevImg& src, const VsWarpCoef* d_coef, const VsDevImg& dst,
                 int dst_x0, int dst_y0, int mode, int border, int32_t* d_tab)
{
    VS_REQUIRE(ctx, src.batch == dst.batch, "bgr_warp: batch mismatch");
    if (mode == VS_WARP_LANCZOS2) {
        VS_REQUIRE(ctx, d_tab, "bgr_warp: the Lanczos-2 mode needs the table scratch");
        if (vs_warp_rows_usable(src, mode, border)) {
            CUtensorMap map;
            const cuuint64_t dims[3] = {(cuuint64_t)(src.w * 3 / 4), (cuuint64_t)src.h, (cuuint64_t)src.batch};
            const cuuint64_t strides[2] = {(cuuint64_t)src.stride, (cuuint64_t)(src.batch > 1 ? src.batch_stride : src.stride * src.h)};
            const cuuint32_t box[3] = {(cuuint32_t)VS_WARP_ROWS_BOX_WORDS, (cuuint32_t)VS_WARP_LZ_BOX_ROWS, 1}, estr[3] = {1, 1, 1};
            CUresult r = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(vs_tensor_map_encoder())(
                &map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) return vsk_bgr_warp_lz(ctx, &map, src, nullptr, d_coef, dst, dst_x0, dst_y0, border, d_tab);
        }
        return vsk_bgr_warp_lz(ctx, nullptr, src, nullptr, d_coef, dst, dst_x0, dst_y0, border, d_tab);
    }
    if (d_tab && vs_warp_rows_usable(src, mode, border)) {
        // any device image as a u32 [image][row][word] tensor
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)(src.w * 3 / 4), (cuuint64_t)src.h, (cuuint64_t)src.batch};
        const cuuint64_t strides[2] = {(cuuint64_t)src.stride, (cuuint64_t)(src.batch > 1 ? src.batch_stride : src.stride * src.h)};
        const cuuint32_t box[3] = {(cuuint32_t)VS_WARP_ROWS_BOX_WORDS, (cuuint32_t)VS_WARP_ROWS_BOX_ROWS, 1}, estr[3] = {1, 1, 1};
        CUresult r = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(vs_tensor_map_encoder())(
            &map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            return vsk_bgr_warp_slots_rows(ctx, &map, src, nullptr, d_coef, dst, dst_x0, dst_y0, d_tab, mode);
    }
    return vsk_bgr_warp_slots(ctx, src, nullptr, d_coef, dst, dst_x0, dst_y0, mode, border);
}
