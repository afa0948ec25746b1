// vs_kernels_sparse.cu — keypoint kernels: per-tile gradient argmax, Jacobians, Lanczos-2
// warp-diff, and the per-pair inverse-compositional Gauss-Newton solver that runs the whole
// coarse-to-fine loop of VideoAligner::AlignNextFrame (alignment.cpp:390-693) on the device,
// one CTA per frame pair, many pairs per launch.
#include "vs_internal.h"
#include "vs_introselect.cuh"
#include "vs_linalg4.cuh"

namespace {

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// ------------------------------------------------------------- grad_argmax (f32 planes)
// generators.cpp:260-294.  One warp per tile.  Key = (bits(|g|)+1) << 32 | (N*N-1 - scan
// index): the maximum key is the first maximum in (r.y outer, r.x inner) order; a NaN never
// wins (as with Halide's strict '>' against the running best).
__global__ void __launch_bounds__(256)
k_grad_argmax(const float* __restrict__ gx, int64_t gxs, const float* __restrict__ gy, int64_t gys,
              int tile, int tw, int th, uint16_t* __restrict__ lmx, uint16_t* __restrict__ lmy)
{
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= tw * th) return;
    const int tx = wid % tw, ty = wid / tw;
    const int n = tile * tile;
    unsigned long long bx = 0, by = 0;
    for (int p = lane; p < n; p += 32) {
        int ry = p / tile, rx = p - ry * tile;
        float a = fabsf(__ldg(gx + (size_t)(ty * tile + ry) * gxs + tx * tile + rx));
        float b = fabsf(__ldg(gy + (size_t)(ty * tile + ry) * gys + tx * tile + rx));
        unsigned long long inv = (unsigned long long)(n - 1 - p);
        unsigned long long ka = isnan(a) ? 0ull : (((unsigned long long)__float_as_uint(a) + 1ull) << 32) | inv;
        unsigned long long kb = isnan(b) ? 0ull : (((unsigned long long)__float_as_uint(b) + 1ull) << 32) | inv;
        bx = ka > bx ? ka : bx;
        by = kb > by ? kb : by;
    }
    bx = warp_max_u64(bx);
    by = warp_max_u64(by);
    if (lane == 0) {
        int px = (bx >> 32) == 0 ? 0 : n - 1 - (int)(bx & 0xffffffffu);
        int py = (by >> 32) == 0 ? 0 : n - 1 - (int)(by & 0xffffffffu);
        size_t plane = (size_t)tw * th, t = (size_t)ty * tw + tx;
        lmx[t] = (uint16_t)(px % tile + tx * tile);
        lmx[plane + t] = (uint16_t)(px / tile + ty * tile);
        lmy[t] = (uint16_t)(py % tile + tx * tile);
        lmy[plane + t] = (uint16_t)(py / tile + ty * tile);
    }
}

// ------------------------------------------------------------- sparse_jac (f32 planes)
// generators.cpp:332-386
__device__ __forceinline__ float4 jac_x(float g, int ix, int iy, int w, int h)
{
    float cx = __fmul_rn((float)w, 0.5f), cy = __fmul_rn((float)h, 0.5f);
    float scale = __fdiv_rn(1.f, (float)w);
    float u = __fsub_rn((float)ix, cx), v = __fsub_rn((float)iy, cy);
    float g2 = __fmul_rn(2.f, g);
    return make_float4(__fmul_rn(__fmul_rn(g2, u), scale), __fmul_rn(__fmul_rn(g2, -v), scale), g2, 0.f);
}
__device__ __forceinline__ float4 jac_y(float g, int ix, int iy, int w, int h)
{
    float cx = __fmul_rn((float)w, 0.5f), cy = __fmul_rn((float)h, 0.5f);
    float scale = __fdiv_rn(1.f, (float)w);
    float u = __fsub_rn((float)ix, cx), v = __fsub_rn((float)iy, cy);
    float g2 = __fmul_rn(2.f, g);
    return make_float4(__fmul_rn(__fmul_rn(g2, v), scale), __fmul_rn(__fmul_rn(g2, u), scale), 0.f, g2);
}

__global__ void __launch_bounds__(128)
k_sparse_jac(const float* __restrict__ gx, int64_t gxs, const float* __restrict__ gy, int64_t gys,
             int w, int h, const uint16_t* __restrict__ lmx, const uint16_t* __restrict__ lmy,
             int tw, int th, float* __restrict__ jx, float* __restrict__ jy)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = tw * th;
    if (t >= plane) return;
    int ix0 = min((int)lmx[t], w - 1), iy0 = min((int)lmx[plane + t], h - 1);
    int ix1 = min((int)lmy[t], w - 1), iy1 = min((int)lmy[plane + t], h - 1);
    float4 a = jac_x(__ldg(gx + (size_t)iy0 * gxs + ix0), ix0, iy0, w, h);
    float4 b = jac_y(__ldg(gy + (size_t)iy1 * gys + ix1), ix1, iy1, w, h);
    jx[t] = a.x; jx[plane + t] = a.y; jx[2 * plane + t] = a.z; jx[3 * plane + t] = a.w;
    jy[t] = b.x; jy[plane + t] = b.y; jy[2 * plane + t] = b.z; jy[3 * plane + t] = b.w;
}

// ------------------------------------------------------------- keyframe features (fused)
// ComputeKeyFrame (alignment.cpp:237-276) without materialising the f32 gradient planes:
// gray level -> per-tile argmax of |gx| and |gy| -> keypoint + Jacobian.  One warp per tile,
// all levels and all requested slots in one launch.  2|g| = |I(+1) - I(-1)| is an integer in
// [0,255], so key = 2|g| << 16 | (N*N-1 - scan index) reproduces Halide's first-maximum rule.
__global__ void __launch_bounds__(256)
k_keyframe_features_generic(VsClipGeom g, const uint8_t* __restrict__ pyr, const int32_t* __restrict__ slots,
                            uint32_t* __restrict__ kp, float4* __restrict__ jac)
{
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= g.total_tiles) return;
    const int slot = slots[blockIdx.y];
    int lvl = 0;
    while (lvl + 1 < g.levels && wid >= (int)g.lv[lvl + 1].tile_off) lvl++;
    const VsLevel L = g.lv[lvl];
    const int t = wid - (int)L.tile_off;
    const int tx = t % L.tw, ty = t / L.tw;
    const uint8_t* img = pyr + (size_t)slot * g.pyr_slot_bytes + L.img_off;
    const int N = L.tile, n = N * N;
    uint32_t bx = 0, by = 0;
    for (int p = lane; p < n; p += 32) {
        int ry = p / N, rx = p - ry * N;
        int x = tx * N + rx, y = ty * N + ry;
        const uint8_t* row = img + (size_t)y * L.pitch;
        int xp = __ldg(row + min(x + 1, L.w - 1)), xm = __ldg(row + max(x - 1, 0));
        int yp = __ldg(img + (size_t)min(y + 1, L.h - 1) * L.pitch + x);
        int ym = __ldg(img + (size_t)max(y - 1, 0) * L.pitch + x);
        uint32_t inv = (uint32_t)(n - 1 - p);
        bx = max(bx, ((uint32_t)abs(xp - xm) << 16) | inv);
        by = max(by, ((uint32_t)abs(yp - ym) << 16) | inv);
    }
    bx = __reduce_max_sync(0xffffffffu, bx);
    by = __reduce_max_sync(0xffffffffu, by);
    if (lane < 2) {
        const int axis = lane;
        int p = n - 1 - (int)((axis == 0 ? bx : by) & 0xffffu);
        int x = tx * N + p % N, y = ty * N + p / N;
        float gval;
        float4 J;
        if (axis == 0) {
            const uint8_t* row = img + (size_t)y * L.pitch;
            gval = __fmul_rn(0.5f, __fsub_rn((float)__ldg(row + min(x + 1, L.w - 1)), (float)__ldg(row + max(x - 1, 0))));
            J = jac_x(gval, x, y, L.w, L.h);
        } else {
            gval = __fmul_rn(0.5f, __fsub_rn((float)__ldg(img + (size_t)min(y + 1, L.h - 1) * L.pitch + x),
                                             (float)__ldg(img + (size_t)max(y - 1, 0) * L.pitch + x)));
            J = jac_y(gval, x, y, L.w, L.h);
        }
        size_t o = ((size_t)slot * 2 + axis) * g.total_tiles + wid;
        kp[o] = ((uint32_t)y << 16) | (uint32_t)x;
        jac[o] = J;
    }
}

// ------------------------------------------------------------- keyframe features (banded)
// Same result as k_keyframe_features_generic, organised for bandwidth: one CTA owns one tile
// row (a band of N image rows, full width) of one level of one slot.  A thread streams down
// the band on a 16-pixel column group: one 16-byte load per row (+ two halo bytes), a
// three-row register window, |I(x+1)-I(x-1)| and |I(y+1)-I(y-1)| four pixels at a time with
// funnel shifts and VABSDIFF4, and a running per-column maximum of (2|g| << 8 | 255 - ry),
// one PRMT + one max per pixel and axis.  After the band the 16 column maxima are merged per
// tile in registers, pushed to shared-memory accumulators with atomicMax on the full key
// (2|g| << 16 | N*N-1 - scan index: the maximum is Halide's first maximum in (ry, rx) order),
// and one thread per tile emits keypoint + Jacobian.
constexpr int KF_THREADS = 128;
constexpr int KF_MAX_TW = 512;     // tiles per band the shared accumulators hold

__device__ __forceinline__ uint32_t kf_word(const uint4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// replicate pixel (nvalid-1) into bytes nvalid..15 (repeat-edge for the last column group of a row)
__device__ __forceinline__ void kf_clamp_tail(uint4& v, int nvalid)
{
    if (nvalid >= 16) return;
    uint32_t wv[4] = {v.x, v.y, v.z, v.w};
    const int last = nvalid - 1;
    const uint32_t edge = (wv[last >> 2] >> (8 * (last & 3))) & 0xffu;
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (i >= nvalid) wv[i >> 2] = (wv[i >> 2] & ~(0xffu << (8 * (i & 3)))) | (edge << (8 * (i & 3)));
    v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
}

__device__ __forceinline__ uint4 kf_load_row(const uint8_t* __restrict__ img, int pitch, int w, int h, int y, int x0)
{
    uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)vs_clampi(y, 0, h - 1) * pitch + x0));
    kf_clamp_tail(v, w - x0);
    return v;
}

// bands one CTA handles on a level: a band of the narrow levels keeps only one or two warps busy (16 pixels per
// thread), and a CTA lasts as long as the band is tall whatever its width, so the narrow levels put several bands
// side by side in one CTA (a team of 1, 2 or 4 warps per band).  Shared by the kernel and its launcher.
__host__ __device__ inline int kf_bands_per_cta(const VsLevel& L)
{
    const int groups = (L.tw * L.tile + 15) / 16;
    int b = groups <= 32 ? 4 : (groups <= 64 ? 2 : 1);
    while (b > 1 && L.tw * b > KF_MAX_TW) b >>= 1;
    return b;
}

__global__ void __launch_bounds__(KF_THREADS, 8)
k_keyframe_features(VsClipGeom g, const uint8_t* __restrict__ pyr, const int32_t* __restrict__ slots,
                    uint32_t* __restrict__ kp, float4* __restrict__ jac)
{
    __shared__ uint32_t acc_all[2 * KF_MAX_TW];
    // which (level, band group) is this CTA?
    int lvl = 0, grp = blockIdx.x;
    int bpc = kf_bands_per_cta(g.lv[0]);
    while (grp >= (g.lv[lvl].th + bpc - 1) / bpc) {
        grp -= (g.lv[lvl].th + bpc - 1) / bpc;
        lvl++;
        bpc = kf_bands_per_cta(g.lv[lvl]);
    }
    const VsLevel L = g.lv[lvl];
    const int team_size = KF_THREADS / bpc, team = threadIdx.x / team_size, ttid = threadIdx.x - team * team_size;
    const int band = grp * bpc + team;
    const bool live = band < L.th;
    // accumulators of this team: [axis][KF_MAX_TW / bpc]
    uint32_t* const acc0 = acc_all + team * (2 * KF_MAX_TW / bpc);
    uint32_t* const acc1 = acc0 + KF_MAX_TW / bpc;
    const int slot = slots[blockIdx.y];
    const uint8_t* img = pyr + (size_t)slot * g.pyr_slot_bytes + L.img_off;
    const int N = L.tile, NN1 = N * N - 1;
    const int y0 = band * N;
    const int wtiles = live ? L.tw * N : 0;           // columns that belong to a tile
    for (int i = ttid; i < L.tw; i += team_size) { acc0[i] = 0u; acc1[i] = 0u; }
    __syncthreads();

    for (int x0 = ttid * 16; x0 < wtiles; x0 += team_size * 16) {
        uint32_t mx[8], my[8];          // running column maxima, two 16-bit keys per register (VIMNMX.U16x2)
#pragma unroll
        for (int i = 0; i < 8; i++) { mx[i] = 0u; my[i] = 0u; }
        uint4 prev = kf_load_row(img, L.pitch, L.w, L.h, y0 - 1, x0);
        uint4 cur = kf_load_row(img, L.pitch, L.w, L.h, y0, x0);
#pragma unroll 2
        for (int ry = 0; ry < N; ry++) {
            const int y = y0 + ry;
            const uint4 next = kf_load_row(img, L.pitch, L.w, L.h, y + 1, x0);
            const uint8_t* row = img + (size_t)y * L.pitch;
            const uint32_t lb = x0 > 0 ? (uint32_t)__ldg(row + x0 - 1) : (cur.x & 0xffu);
            const uint32_t rb = x0 + 16 < L.w ? (uint32_t)__ldg(row + x0 + 16) : (cur.w >> 24);
            const uint32_t crow = 255u - (uint32_t)ry;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t wc = kf_word(cur, k);
                const uint32_t wl = k == 0 ? (lb << 24) : kf_word(cur, k - 1);
                const uint32_t wr = k == 3 ? rb : kf_word(cur, k + 1);
                const uint32_t left = __funnelshift_r(wl, wc, 24);     // pixels x-1 .. x+2
                const uint32_t right = __funnelshift_r(wc, wr, 8);     // pixels x+1 .. x+4
                const uint32_t gx4 = __vabsdiffu4(right, left);
                const uint32_t gy4 = __vabsdiffu4(kf_word(next, k), kf_word(prev, k));
                // key' = 2|g| << 8 | 255 - ry, 16 bits per pixel: bytes (crow, g_a, crow, g_b) hold two pixels
                mx[2 * k + 0] = __vmaxu2(mx[2 * k + 0], __byte_perm(gx4, crow, 0x1404));
                mx[2 * k + 1] = __vmaxu2(mx[2 * k + 1], __byte_perm(gx4, crow, 0x3424));
                my[2 * k + 0] = __vmaxu2(my[2 * k + 0], __byte_perm(gy4, crow, 0x1404));
                my[2 * k + 1] = __vmaxu2(my[2 * k + 1], __byte_perm(gy4, crow, 0x3424));
            }
            prev = cur; cur = next;
        }
        // merge the 16 columns per tile, then one shared atomic per (tile, axis)
        int tile = x0 / N;
        int rx = x0 - tile * N;
        uint32_t bx = 0u, by = 0u;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (x0 + i < wtiles) {
                const uint32_t mxi = (mx[i >> 1] >> (16 * (i & 1))) & 0xffffu, myi = (my[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                const uint32_t kx = ((mxi >> 8) << 16) | (uint32_t)(NN1 - (int)(255u - (mxi & 0xffu)) * N - rx);
                const uint32_t ky = ((myi >> 8) << 16) | (uint32_t)(NN1 - (int)(255u - (myi & 0xffu)) * N - rx);
                bx = max(bx, kx); by = max(by, ky);
            }
            rx++;
            if (rx == N || i == 15) {
                if (bx | by) { atomicMax(&acc0[tile], bx); atomicMax(&acc1[tile], by); }
                bx = by = 0u; rx = 0; tile++;
            }
        }
    }
    __syncthreads();

    for (int i = ttid; i < (live ? 2 * L.tw : 0); i += team_size) {
        const int axis = i >= L.tw, tx = i - axis * L.tw;
        const int p = NN1 - (int)((axis ? acc1 : acc0)[tx] & 0xffffu);
        const int x = tx * N + p % N, y = y0 + p / N;
        float4 J;
        if (axis == 0) {
            const uint8_t* row = img + (size_t)y * L.pitch;
            const float gval = __fmul_rn(0.5f, __fsub_rn((float)__ldg(row + min(x + 1, L.w - 1)), (float)__ldg(row + max(x - 1, 0))));
            J = jac_x(gval, x, y, L.w, L.h);
        } else {
            const float gval = __fmul_rn(0.5f, __fsub_rn((float)__ldg(img + (size_t)min(y + 1, L.h - 1) * L.pitch + x),
                                                         (float)__ldg(img + (size_t)max(y - 1, 0) * L.pitch + x)));
            J = jac_y(gval, x, y, L.w, L.h);
        }
        const size_t o = ((size_t)slot * 2 + axis) * g.total_tiles + L.tile_off + (size_t)band * L.tw + tx;
        kp[o] = ((uint32_t)y << 16) | (uint32_t)x;
        jac[o] = J;
    }
}

// ------------------------------------------------------------- sparse_warpdiff (standalone)
// generators.cpp:646-700
__global__ void __launch_bounds__(128)
k_sparse_warpdiff(const uint8_t* __restrict__ tmpl, int64_t ts, const uint8_t* __restrict__ key, int64_t ks,
                  int w, int h, const uint16_t* __restrict__ lm, int plane,
                  float A, float B, float TX, float TY, uint16_t* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plane) return;
    int px = min((int)lm[t], w - 1), py = min((int)lm[plane + t], h - 1);
    float s = vs_lanczos_sample(key, w, h, (int)ks, (float)px, (float)py, A, B, TX, TY);
    float d = fabsf(__fsub_rn(s, (float)__ldg(tmpl + (size_t)py * ts + px)));
    d = fmaxf(fminf(d, 65535.0f), 0.0f);
    out[t] = (uint16_t)d;
}

// ------------------------------------------------------------- sparse_ica (standalone)
// generators.cpp:429-596.  Single CTA; f32 J*r products accumulated in f64.
__global__ void __launch_bounds__(256)
k_sparse_ica(const uint8_t* __restrict__ tmpl, int64_t ts, const uint8_t* __restrict__ key, int64_t ks,
             int w, int h, const uint16_t* __restrict__ selx, int kx, const uint16_t* __restrict__ sely, int ky,
             const float* __restrict__ jx, const float* __restrict__ jy,
             float A, float B, float TX, float TY, double* __restrict__ out)
{
    __shared__ double red[8][4];
    double acc[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < kx + ky; i += blockDim.x) {
        const bool isx = i < kx;
        const int j = isx ? i : i - kx, k = isx ? kx : ky;
        const uint16_t* sel = isx ? selx : sely;
        const float* jac = isx ? jx : jy;
        int px = sel[j], py = sel[k + j];
        float warped = vs_lanczos_sample(key, w, h, (int)ks, (float)px, (float)py, A, B, TX, TY);
        int tx = min(px, w - 1), ty = min(py, h - 1);
        float residual = __fsub_rn((float)__ldg(tmpl + (size_t)ty * ts + tx), warped);
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] += (double)__fmul_rn(jac[(size_t)c * k + j], residual);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double v = vs_warp_reduce_sum(acc[c]);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) s += red[wv][threadIdx.x];
        out[threadIdx.x] = s * 0.5;
    }
}

// ------------------------------------------------------------- per-pair solver
// CTA size is a template parameter; three CTAs per SM so that the 299 pairs of a 300-frame clip run as one wave
constexpr int SOLVE_MAX_WARPS = 32;
constexpr int SOLVE_WARPS = SOLVE_MAX_WARPS;   // array extents only
enum { FLAG_CONTINUE = 0, FLAG_CONVERGED = 1, FLAG_FAIL = 2 };

// (sized by the CTA's own warp count: every KB of shared memory is a KB of L1 taken from the gathers)
template <int NWARPS>
struct SolveShared {
    double T[4];
    double Hinv[16];
    double red[NWARPS][12];           // Hessian partial sums; its first 16 doubles then carry H to the lanes of warp 0
    double red4[NWARPS][4];           // Gauss-Newton partial sums (a buffer of their own: thread 0 may still be busy after the Hessian reduce)
    double c0[4][2], c1[4][2];
    int flag;
    int status;
};

// ---- block-parallel exact replay of std::nth_element (libstdc++ introselect) ------------
// The serial algorithm (vs_introselect.cuh) is: loop { median-of-3 pivot to `first`;
// unguarded Hoare partition of [first+1,last); keep the side that holds nth } until the range
// is <= 3 long, then insertion sort.  Only the Hoare partition is long, and it is
// order-independent enough to run in parallel while producing the identical permutation:
//   the left pointer stops at the successive positions a_0 < a_1 < ... whose value is >= pivot,
//   the right pointer at the successive positions b_0 > b_1 > ... whose value is <= pivot
//   (both in the array as it was when the partition started), and swap k exchanges a_k and b_k
//   as long as a_k < b_k.  With m such swaps the function returns min(a_m, b_{m-1}): the first
//   position past a_{m-1} that now holds a value >= pivot.
// So a round is: every thread counts the candidates of its slice, a block scan turns counts
// into ranks, candidates are scattered to posL[rank] / posR[rank], m is a block-wide count of
// a_k < b_k, and the m swaps touch disjoint positions.  Median-of-3, the bookkeeping and the
// tail (ranges <= SEL_SERIAL long, or an exhausted depth limit -> heap select) run on one thread
// per axis with the serial code, so every comparison-dependent choice is libstdc++'s own.
// Both keypoint axes are processed in the same rounds.
constexpr int SEL_SERIAL = 32;

struct SelAxis {
    int first, last, depth, done;
    uint32_t pivot;
    int nL, nR;
    int cutL, cutR;     // posL[m] and posR[m-1] of the current round (see the swap phase)
};

template <int NWARPS>
struct SelShared {
    SelAxis ax[2];
    uint32_t warp_tot[2][NWARPS];
    int warp_cnt[2][NWARPS];
};

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// keys[a]: n packed keys of axis a; pos[a]: 2*n u16 (posL then posR).  nth < n.
template <int SOLVE_THREADS>
__device__ void block_nth_element2(uint32_t* const keys0, uint32_t* const keys1, uint16_t* const pos0, uint16_t* const pos1,
                                   const int n, const int nth, SelShared<SOLVE_THREADS / 32>& ss, long long* rounds = nullptr,
                                   long long* serial_cycles = nullptr)
{
    constexpr int NWARPS = SOLVE_THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n == 0 || nth == n) return;
    if (tid < 2) {
        ss.ax[tid].first = 0; ss.ax[tid].last = n; ss.ax[tid].depth = vs_sel::lg(n) * 2; ss.ax[tid].done = 0;
    }
    __syncthreads();
    while (true) {
        // ---- per-axis serial step: finish, or pick the pivot of the next round
        if (tid == 0 || tid == 32) {
            const int a = tid >> 5;
            SelAxis& st = ss.ax[a];
            uint32_t* v = a ? keys1 : keys0;
            const long long t0 = serial_cycles ? clock64() : 0;
            if (!st.done) {
                if (st.last - st.first <= SEL_SERIAL || st.depth == 0) {
                    vs_sel::introselect_from(v, st.first, nth, st.last, st.depth);
                    st.done = 1;
                } else {
                    --st.depth;
                    st.cutL = st.cutR = 0x7fffffff;
                    const int mid = st.first + (st.last - st.first) / 2;
                    vs_sel::move_median_to_first(v, st.first, st.first + 1, mid, st.last - 1);
                    st.pivot = v[st.first];
                }
            }
            if (serial_cycles) *serial_cycles += clock64() - t0;
        }
        __syncthreads();
        const bool act0 = !ss.ax[0].done, act1 = !ss.ax[1].done;
        if (!act0 && !act1) break;
        if (rounds) ++*rounds;

        // ---- count the candidates of this thread's slice, both axes
        int c0[2], c1[2];
        uint32_t cnt[2] = {0u, 0u};          // nL | nR << 16
#pragma unroll
        for (int a = 0; a < 2; a++) {
            c0[a] = c1[a] = 0;
            if (!(a ? act1 : act0)) continue;
            const SelAxis& st = ss.ax[a];
            const uint32_t* v = a ? keys1 : keys0;
            const int f0 = st.first + 1, len = st.last - f0;
            const int chunk = (len + SOLVE_THREADS - 1) / SOLVE_THREADS;
            c0[a] = min(f0 + tid * chunk, st.last);
            c1[a] = min(c0[a] + chunk, st.last);
            const uint32_t pv = st.pivot;
            uint32_t c = 0;
            for (int i = c0[a]; i < c1[a]; i++) {
                const uint32_t e = v[i];
                c += (vs_sel::less(e, pv) ? 0u : 1u) + (vs_sel::less(pv, e) ? 0u : 0x10000u);
            }
            cnt[a] = c;
        }
        uint32_t incl[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
            incl[a] = warp_incl_scan_u32(cnt[a]);
            if (lane == 31) ss.warp_tot[a][warp] = incl[a];
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 2; a++) {
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NWARPS; w++) {
                const uint32_t t = ss.warp_tot[a][w];
                if (w < warp) before += t;
                total += t;
            }
            incl[a] += before;
            if (!(a ? act1 : act0)) continue;
            const uint32_t* v = a ? keys1 : keys0;
            uint16_t* posL = a ? pos1 : pos0;
            uint16_t* posR = posL + n;
            const uint32_t pv = ss.ax[a].pivot;
            const int nL = (int)(total & 0xffffu), nR = (int)(total >> 16);
            int offL = (int)((incl[a] - cnt[a]) & 0xffffu);            // candidates in lower slices
            int offR = nR - (int)(incl[a] >> 16);                      // candidates in higher slices
            for (int i = c0[a]; i < c1[a]; i++)
                if (!vs_sel::less(v[i], pv)) posL[offL++] = (uint16_t)i;
            for (int i = c1[a] - 1; i >= c0[a]; i--)
                if (!vs_sel::less(pv, v[i])) posR[offR++] = (uint16_t)i;
            if (tid == 0) { ss.ax[a].nL = nL; ss.ax[a].nR = nR; }
        }
        __syncthreads();

        // ---- the swaps.  a_k increases and b_k decreases with k, so "a_k < b_k" holds exactly for k < m: every thread
        //      decides its own swaps without knowing m, m is the block-wide count, and the two positions the bookkeeping
        //      needs (posL[m], posR[m-1]) are the smallest a_k not swapped and the smallest b_k swapped.
#pragma unroll
        for (int a = 0; a < 2; a++) {
            int c = 0;
            if (a ? act1 : act0) {
                uint32_t* v = a ? keys1 : keys0;
                const uint16_t* posL = a ? pos1 : pos0;
                const uint16_t* posR = posL + n;
                const int nL = ss.ax[a].nL, lim = min(nL, ss.ax[a].nR);
                int minL = 0x7fffffff, minR = 0x7fffffff;
                for (int k = tid; k <= lim; k += SOLVE_THREADS) {
                    if (k == lim) { if (k < nL) minL = min(minL, (int)posL[k]); break; }
                    const int i = posL[k], j = posR[k];
                    if (i < j) {
                        const uint32_t t = v[i]; v[i] = v[j]; v[j] = t;
                        c++;
                        minR = min(minR, j);
                    } else {
                        minL = min(minL, i);
                    }
                }
                minL = __reduce_min_sync(0xffffffffu, minL);
                minR = __reduce_min_sync(0xffffffffu, minR);
                if (lane == 0) {
                    if (minL != 0x7fffffff) atomicMin(&ss.ax[a].cutL, minL);
                    if (minR != 0x7fffffff) atomicMin(&ss.ax[a].cutR, minR);
                }
            }
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) ss.warp_cnt[a][warp] = c;
        }
        __syncthreads();
        // ---- the bookkeeping of __introselect
        if (tid == 0 || tid == 32) {
            const int a = tid >> 5;
            SelAxis& st = ss.ax[a];
            if (!st.done) {
                int m = 0;
#pragma unroll
                for (int w = 0; w < NWARPS; w++) m += ss.warp_cnt[a][w];
                int cut = st.last;
                if (m < st.nL) cut = st.cutL;                       // posL[m]
                if (m > 0) cut = min(cut, st.cutR);                  // posR[m - 1]
                if (cut <= nth) st.first = cut; else st.last = cut;
            }
        }
        // the next iteration's serial step runs on the same threads (0 and 32): no barrier needed here
    }
}

// ---- the same replay without candidate lists, for CTAs that have an SM to themselves ------------
// (512 / 1024 threads: 4K clips, up to 148 pairs per launch, the per-frame API.)  The range is cut into chunks
// of 32 consecutive keys; a warp ballot gives every chunk one bit mask of ">= pivot" and one of "<= pivot" keys,
// a warp scan of the chunks' population counts gives their ranks, and the candidate of rank k from the left
// finds its partner, rank k from the right, by a binary search over the chunk ranks and a bit-select in that
// chunk's mask.  a_k - b_k grows with k, so every candidate decides its own swap and a warp stops at its first
// chunk holding a candidate that does not swap.  The state is 16 bytes per 32 keys, in shared memory at every
// level (no global scratch, no L2 round trips), and the two keypoint axes run on their own half of the warps
// with their own named barrier: an axis never waits for the rounds of the other.  Measured against the list
// form above (mean cycles per pair): 4K 1.92 M -> 1.78 M, 720p 0.68 M -> 0.63 M, one 1080p pair 0.75 M -> 0.70 M.
// With three 256-thread CTAs per SM (the 299 pairs of a 1080p clip) the extra 5 KB per CTA push the
// shared-memory carve-out from 132 to 164 KB, and the L1 lost to it costs more than the selection gains
// (0.774 -> 0.803 ms): that configuration keeps the lists.
__host__ __device__ inline int sel_chunks(int max_tiles) { return (max_tiles + 31) / 32 + 1; }

// position of the set bit of rank j (from bit 0) in a mask that has more than j bits set
__device__ __forceinline__ int sel_nth_bit(uint32_t mask, int j)
{
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const int c = __popc(mask & (((1u << w) - 1u) << pos));
        if (j >= c) { j -= c; pos += w; }
    }
    return pos;
}

template <int GROUP_THREADS>
__device__ __forceinline__ void sel_group_bar(int axis)
{
    asm volatile("bar.sync %0, %1;" ::"r"(axis + 1), "n"(GROUP_THREADS) : "memory");
}

// keys[a]: n packed keys of axis a; selbuf: [2][4][nc] words (masks and ranks of the chunks).  nth < n.
template <int SOLVE_THREADS>
__device__ void block_nth_element2_masks(uint32_t* const keys0, uint32_t* const keys1, uint32_t* const selbuf, const int nc,
                                         const int n, const int nth, SelAxis* const ax, long long* rounds = nullptr,
                                         long long* serial_cycles = nullptr)
{
    constexpr int GT = SOLVE_THREADS / 2;      // threads of an axis group
    constexpr int G = GT / 32;                 // its warps
    if (n == 0 || nth == n) return;
    const int axis = threadIdx.x / GT, gtid = threadIdx.x % GT, gw = gtid >> 5, lane = gtid & 31;
    uint32_t* const v = axis ? keys1 : keys0;
    SelAxis& st = ax[axis];
    uint32_t* const maskL = selbuf + (size_t)axis * 4 * nc;
    uint32_t* const maskR = maskL + nc;
    uint32_t* const preL = maskR + nc;         // candidates of lower chunks, ">= pivot"
    uint32_t* const preR = preL + nc;          // candidates of lower chunks, "<= pivot"
    const uint32_t lt = (1u << lane) - 1u;
    if (gtid == 0) { st.first = 0; st.last = n; st.depth = vs_sel::lg(n) * 2; st.done = 0; }
    while (true) {
        // ---- serial step: finish, or pick the pivot of the next round (the thread that did the bookkeeping)
        if (gtid == 0) {
            const long long t0 = serial_cycles && axis == 0 ? clock64() : 0;
            if (st.last - st.first <= SEL_SERIAL || st.depth == 0) {
                vs_sel::introselect_from(v, st.first, nth, st.last, st.depth);
                st.done = 1;
            } else {
                --st.depth;
                st.nL = 0;                              // swaps of the round
                st.cutL = st.cutR = 0x7fffffff;
                const int mid = st.first + (st.last - st.first) / 2;
                vs_sel::move_median_to_first(v, st.first, st.first + 1, mid, st.last - 1);
                st.pivot = v[st.first];
            }
            if (serial_cycles && axis == 0) *serial_cycles += clock64() - t0;
        }
        sel_group_bar<GT>(axis);
        if (st.done) break;
        if (rounds && axis == 0) ++*rounds;
        const int f0 = st.first + 1, last = st.last;
        const uint32_t pv = vs_sel::key_value(st.pivot);
        const int nch = (last - f0 + 31) >> 5;

        // ---- masks of the chunks
        for (int c = gw; c < nch; c += G) {
            const int i = f0 + 32 * c + lane;
            const bool valid = i < last;
            const uint32_t e = valid ? vs_sel::key_value(v[i]) : 0u;
            const uint32_t mL = __ballot_sync(0xffffffffu, valid && e >= pv);
            const uint32_t mR = __ballot_sync(0xffffffffu, valid && e <= pv);
            if (lane == 0) { maskL[c] = mL; maskR[c] = mR; }
        }
        sel_group_bar<GT>(axis);

        // ---- ranks of the chunks (every warp computes and writes the same values: no barrier)
        uint32_t nR = 0;
        {
            uint32_t nL = 0;
            for (int base = 0; base < nch; base += 32) {
                const int c = base + lane;
                const uint32_t cl = c < nch ? __popc(maskL[c]) : 0u, cr = c < nch ? __popc(maskR[c]) : 0u;
                uint32_t il = cl, ir = cr;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t tl = __shfl_up_sync(0xffffffffu, il, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
                    if (lane >= o) { il += tl; ir += tr; }
                }
                if (c < nch) { preL[c] = nL + il - cl; preR[c] = nR + ir - cr; }
                nL += __shfl_sync(0xffffffffu, il, 31);
                nR += __shfl_sync(0xffffffffu, ir, 31);
            }
        }
        __syncwarp();

        // ---- the swaps
        const int steps = 32 - __clz(nch - 1);      // 2^steps >= nch (nch >= 1)
        int cnt = 0, minL = 0x7fffffff, minR = 0x7fffffff;
        for (int c = gw; c < nch; c += G) {
            const uint32_t mL = maskL[c];
            if (mL == 0u) continue;
            const bool cand = (mL >> lane) & 1u;
            const uint32_t k = preL[c] + __popc(mL & lt);
            const int a = f0 + 32 * c + lane;
            bool swapped = false;
            if (cand && k < nR) {
                const uint32_t r = nR - 1u - k;          // the partner's rank from the left
                int lo = 0, hi = nch - 1;
                for (int s = 0; s < steps; s++) {        // largest chunk whose rank is <= r
                    const int mid = (lo + hi + 1) >> 1;
                    if (preR[mid] <= r) lo = mid; else hi = mid - 1;
                }
                const int b = f0 + 32 * lo + sel_nth_bit(maskR[lo], (int)(r - preR[lo]));
                if (a < b) {
                    const uint32_t t = v[a]; v[a] = v[b]; v[b] = t;
                    swapped = true;
                    cnt++;
                    minR = min(minR, b);
                }
            }
            if (cand && !swapped) minL = min(minL, a);
            if (__any_sync(0xffffffffu, cand && !swapped)) break;   // no later candidate swaps either
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        minL = __reduce_min_sync(0xffffffffu, minL);
        minR = __reduce_min_sync(0xffffffffu, minR);
        if (lane == 0) {
            if (cnt) atomicAdd(&st.nL, cnt);
            if (minL != 0x7fffffff) atomicMin(&st.cutL, minL);
            if (minR != 0x7fffffff) atomicMin(&st.cutR, minR);
        }
        sel_group_bar<GT>(axis);
        // ---- the bookkeeping of __introselect
        if (gtid == 0) {
            int cut = st.last;
            if (st.cutL != 0x7fffffff) cut = st.cutL;             // a_m
            if (st.nL > 0) cut = min(cut, st.cutR);               // b_{m-1}
            if (cut <= nth) st.first = cut; else st.last = cut;
        }
    }
}

template <int N, int NWARPS, int STRIDE>
__device__ __forceinline__ void block_reduce(double* v, double (*red)[STRIDE], double* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; i++) {
        double r = vs_warp_reduce_sum(v[i]);
        if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = 0;
            for (int w = 0; w < NWARPS; w++) s += red[w][i];
            total[i] = s;
        }
    }
}

__device__ __forceinline__ double dist2d(const double* a, const double* b)
{
    double dx = a[0] - b[0], dy = a[1] - b[1];
    return sqrt(dx * dx + dy * dy);
}

template <int SOLVE_THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(SOLVE_THREADS, MIN_CTAS)
k_solve_pairs(VsClipGeom g, VsSolveArgs a)
{
    constexpr int NWARPS = SOLVE_THREADS / 32;
    extern __shared__ uint32_t dyn_keys[];   // keys[2][max_tiles]: abs_delta << KEY_SHIFT | tile, then (optionally) pos[2][2*max_tiles] u16
    __shared__ SolveShared<SOLVE_THREADS / 32> sh;
    __shared__ SelShared<SOLVE_THREADS / 32> sel;
    const int tid = threadIdx.x;
    const int pair = blockIdx.x;          // the job: index of scratch and outputs
    if (pair >= a.n_pairs) return;
    // a parameter sweep runs every pair once per parameter set in the same launch
    const int src_pair = a.sweep ? pair % a.sweep_pairs : pair;
    double p_threshold = a.threshold, p_max_displacement = a.max_displacement;
    float p_fraction = a.fraction;
    int p_max_iters = a.max_iters;
    bool p_seed = a.init_T != nullptr;
    if (a.sweep) {
        const VsSweepSet ss = a.sweep[pair / a.sweep_pairs];
        p_threshold = ss.threshold; p_max_displacement = ss.max_displacement; p_fraction = ss.fraction;
        p_max_iters = ss.max_iters; p_seed = p_seed && ss.use_seed != 0;
    }
    const vs_pair pr = a.pairs[src_pair];
    const uint8_t* tpyr = a.pyr + (size_t)pr.template_slot * g.pyr_slot_bytes;
    const uint8_t* kpyr = a.pyr + (size_t)pr.keyframe_slot * g.pyr_slot_bytes;
    const size_t feat = (size_t)pr.keyframe_slot * 2 * g.total_tiles;
    // levels too large for shared memory (8K: 82 944 tiles, 663 KB of keys) keep the keys of the pair in global memory
    // (L2-resident); only the chunk masks of the selection stay on chip
    uint32_t* const keys0 = a.key_scratch ? a.key_scratch + (size_t)pair * 2 * g.max_tiles : dyn_keys;
    uint32_t* const keys1 = keys0 + g.max_tiles;
    // the parallel selection's state: chunk masks and ranks in shared memory when the CTA has its SM to itself, else
    // candidate position lists in shared memory when they fit, else in a global scratch slice
    constexpr bool SEL_MASKS = MIN_CTAS == 1;
    const int sel_nc = sel_chunks(g.max_tiles);
    uint32_t* const selbuf = a.key_scratch ? dyn_keys : dyn_keys + 2 * g.max_tiles;
    uint16_t* const pos0 = a.pos_scratch ? a.pos_scratch + (size_t)pair * 4 * g.max_tiles
                                         : reinterpret_cast<uint16_t*>(dyn_keys + 2 * g.max_tiles);
    uint16_t* const pos1 = pos0 + 2 * g.max_tiles;
    // signed residual template - keyframe(W(p)) of every tile from the warp-diff pass: the first Gauss-Newton iteration of a
    // level evaluates exactly these samples again (same transform), so it reads them back instead of the images
    float* const res = a.res_scratch ? a.res_scratch + (size_t)pair * 2 * g.max_tiles : nullptr;
    // patch cache: the 4x4 keyframe window and the template byte of every tile as the warp-diff pass gathered them.  A later
    // Gauss-Newton iteration whose sample still has the same integer origin (the transform moves by a fraction of a pixel
    // within a level) reads 16 contiguous bytes instead of four sectors of the image; the window's bytes depend on the
    // origin alone, so both routes give the same bits.
    uint4* const patch = a.patch_scratch ? a.patch_scratch + (size_t)pair * 2 * g.max_tiles : nullptr;
    uint8_t* const tbs = a.tb_scratch ? a.tb_scratch + (size_t)pair * 2 * g.max_tiles : nullptr;

    if (tid == 0) {
        sh.T[0] = sh.T[1] = sh.T[2] = sh.T[3] = 0.0;
        if (p_seed) { sh.T[2] = a.init_T[(size_t)src_pair * 2]; sh.T[3] = a.init_T[(size_t)src_pair * 2 + 1]; }   // alignment.cpp:379-387
        sh.status = 1;
        if (a.out_iters) for (int l = 0; l < g.levels; l++) a.out_iters[(size_t)pair * g.levels + l] = 0;
    }

    // debug taps: cycles per phase {warpdiff, select, hessian sums, svd beside the first iteration, gn reduce+update,
    // (rounds of the selection), gn gathers of the later iterations}, summed over levels and per level
    long long clk[7] = {0, 0, 0, 0, 0, 0, 0};
    long long t_prev = a.dbg_clock ? clock64() : 0;
    int lvl = g.levels - 1;
#define VS_CLK(slot) do { if (a.dbg_clock && tid == 0) { long long _t = clock64(); clk[slot] += _t - t_prev;                     \
                          a.dbg_clock[(size_t)pair * VS_CLK_STRIDE + 8 + lvl * 8 + (slot)] += _t - t_prev; t_prev = _t; } } while (0)
    for (; lvl >= 0; lvl--) {
        const VsLevel L = g.lv[lvl];
        const uint8_t* timg = tpyr + L.img_off;
        const uint8_t* kimg = kpyr + L.img_off;
        const uint32_t* const kpl0 = a.kp + feat + L.tile_off;
        const uint32_t* const kpl1 = kpl0 + g.total_tiles;
        const float4* const jcl0 = a.jac + feat + L.tile_off;
        const float4* const jcl1 = jcl0 + g.total_tiles;
        const int nt = L.ntiles;
        const int k = vs_sel::selected_count(nt, p_fraction);
        __syncthreads();   // sh.T of the previous level (or the initial identity) is visible

        // ---- SparseWarpDiff for both keypoint sets with the incoming transform (alignment.cpp:409-431)
        float P[4];
        vs_ul_params_half(sh.T, L.w, L.h, P);
        // the keypoint words go to the key arrays first (coalesced, all in flight together): a sample is then one
        // dependent round trip to L2 (key word in shared memory -> taps) instead of two (keypoint -> taps)
        for (int t = tid; t < nt; t += SOLVE_THREADS) {
            keys0[t] = __ldg(kpl0 + t);
            keys1[t] = __ldg(kpl1 + t);
        }
        __syncthreads();
        // keys[t] is read before the same thread overwrites it with the result
        for (int i = tid; i < 2 * nt; i += SOLVE_THREADS) {
            const int axis = i >= nt, t = i - axis * nt;
            const uint32_t kv = (axis ? keys1 : keys0)[t];
            const int px = min((int)(kv & 0xffffu), L.w - 1), py = min((int)(kv >> 16), L.h - 1);
            VsLzTaps taps;
            vs_lz_fetch(kimg, L.w, L.h, L.pitch, (float)px, (float)py, P[0], P[1], P[2], P[3], taps);
            const uint32_t tb = __ldg(timg + (size_t)py * L.pitch + px);
            if (patch) {
                patch[axis * g.max_tiles + t] = make_uint4(taps.row[0], taps.row[1], taps.row[2], taps.row[3]);
                tbs[axis * g.max_tiles + t] = (uint8_t)tb;
            }
            const float s = vs_lz_eval(taps);
            if (res) res[axis * g.max_tiles + t] = __fsub_rn((float)tb, s);
            float d = fabsf(__fsub_rn(s, (float)tb));
            d = fmaxf(fminf(d, 65535.0f), 0.0f);
            const uint32_t u = (uint32_t)d;
            (axis ? keys1 : keys0)[t] = vs_sel::make_key(u, (uint32_t)t);
            if (a.dbg_warpdiff)
                a.dbg_warpdiff[((size_t)pair * 2 + axis) * g.total_tiles + L.tile_off + t] = (uint16_t)u;
        }
        __syncthreads();

        VS_CLK(0);
        // ---- keep the k smallest: exact replay of std::nth_element (alignment.cpp:460-486)
        // candidate lists: the unused tail of the key arrays when this level's lists fit there (every level but the
        // largest of a clip: no shared memory is added, and the lists of those levels stay out of L2)
        const bool tail_lists = a.pos_scratch != nullptr && 2 * nt <= g.max_tiles;
        const long long rounds0 = clk[5];
        long long serial_cyc = 0;
        if (SEL_MASKS)
            block_nth_element2_masks<SOLVE_THREADS>(keys0, keys1, selbuf, sel_nc, nt, k, sel.ax, a.dbg_clock ? &clk[5] : nullptr,
                                                    a.dbg_clock ? &serial_cyc : nullptr);
        else
            block_nth_element2<SOLVE_THREADS>(keys0, keys1, tail_lists ? reinterpret_cast<uint16_t*>(keys0 + nt) : pos0,
                                              tail_lists ? reinterpret_cast<uint16_t*>(keys1 + nt) : pos1, nt, k, sel,
                                              a.dbg_clock ? &clk[5] : nullptr, a.dbg_clock ? &serial_cyc : nullptr);
        __syncthreads();
        VS_CLK(1);
        if (a.dbg_clock && tid == 0) {
            a.dbg_clock[(size_t)pair * VS_CLK_STRIDE + 8 + lvl * 8 + 5] = clk[5] - rounds0;
            a.dbg_clock[(size_t)pair * VS_CLK_STRIDE + 8 + lvl * 8 + 7] = serial_cyc;   // thread 0's share of the serial steps
        }

        if (a.dbg_order) {
            for (int i = tid; i < 2 * k; i += SOLVE_THREADS) {
                const int axis = i >= k, j = i - axis * k;
                a.dbg_order[((size_t)pair * 2 + axis) * g.total_tiles + L.tile_off + j] = vs_sel::key_tile((axis ? keys1 : keys0)[j]);
            }
            if (tid < 2) a.dbg_count[((size_t)pair * 2 + tid) * g.levels + lvl] = k;
        }

        // ---- H = sum j j^T in f64 (alignment.cpp:278-332); X rows are (a,b,c,0), Y rows (a,b,0,c)
        {
            // The same pass rewrites the selected keys as tile column | tile row << 10 | x offset << 20 | y offset << 25:
            // the keypoint position of every Gauss-Newton sample then follows from shared memory alone.
            double hs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
            for (int i = tid; i < 2 * k; i += SOLVE_THREADS) {
                const int axis = i >= k, j = i - axis * k;
                uint32_t* const kj = (axis ? keys1 : keys0) + j;
                const int t = (int)vs_sel::key_tile(*kj);
                const uint32_t kv = __ldg((axis ? kpl1 : kpl0) + t);
                float4 J = __ldg((axis ? jcl1 : jcl0) + t);
                const int ty = t / L.tw, tx = t - ty * L.tw;
                *kj = (uint32_t)tx | ((uint32_t)ty << 10) | (((kv & 0xffffu) - (uint32_t)(tx * L.tile)) << 20) |
                      (((kv >> 16) - (uint32_t)(ty * L.tile)) << 25);
                double ja = J.x, jb = J.y, jc = axis == 0 ? J.z : J.w;
                hs[0] += ja * ja; hs[1] += ja * jb; hs[2] += jb * jb;
                if (axis == 0) { hs[3] += ja * jc; hs[4] += jb * jc; hs[5] += jc * jc; }
                else           { hs[6] += ja * jc; hs[7] += jb * jc; hs[8] += jc * jc; }
            }
            double tot[9];
            block_reduce<9, NWARPS, 12>(hs, sh.red, tot);
            VS_CLK(2);
            // warp 0: conditioning + inverse (f64 Jacobi SVD spread over the lanes of a quad, bit-identical to the serial
            // vs_condition_and_invert).  Meanwhile warps 1.. run the first Gauss-Newton iteration, which needs only the
            // incoming transform and the selected keypoints.
            if (tid < 32) {
                double* const H = &sh.red[0][0];              // the partial sums are dead once thread 0 holds the totals
                static_assert(NWARPS * 12 >= 16, "H aliases the Hessian partial sums");
                if (tid == 0) {
                    H[0] = tot[0]; H[1] = tot[1]; H[2] = tot[3]; H[3] = tot[6];
                    H[4] = tot[1]; H[5] = tot[2]; H[6] = tot[4]; H[7] = tot[7];
                    H[8] = tot[3]; H[9] = tot[4]; H[10] = tot[5]; H[11] = 0.0;
                    H[12] = tot[6]; H[13] = tot[7]; H[14] = 0.0; H[15] = tot[8];
                }
                __syncwarp();
                double hrow[4], hinv[4];
#pragma unroll
                for (int c = 0; c < 4; c++) hrow[c] = H[(tid & 3) * 4 + c];
                vs_condition_and_invert_quad(hrow, hinv);
                if (tid < 4) {
#pragma unroll
                    for (int c = 0; c < 4; c++) sh.Hinv[tid * 4 + c] = hinv[c];
                }
                if (tid == 0) {
                    // corner bookkeeping (alignment.cpp:585-598)
                    const double cx = L.w * 0.5, cy = L.h * 0.5;
                    const double xr = (double)((float)L.w - 1.f), yb = (double)((float)L.h - 1.f);
                    const double cr[4][2] = {{0.0, 0.0}, {xr, 0.0}, {0.0, yb}, {xr, yb}};
                    for (int c = 0; c < 4; c++) {
                        vs_tf_warp_center(sh.T, cr[c][0], cr[c][1], cx, cy, sh.c0[c]);
                        sh.c1[c][0] = sh.c0[c][0]; sh.c1[c][1] = sh.c0[c][1];
                    }
                    sh.flag = FLAG_CONTINUE;
                }
            }
        }

        // ---- inverse-compositional Gauss-Newton iterations (alignment.cpp:600-668)
        // sum over the selected keypoints of J * (template - keyframe(W(p))), threads first .. first + count - 1
        auto gather = [&](int first, int count, double* b) {
            float Pg[4];
            vs_ul_params_half(sh.T, L.w, L.h, Pg);
            for (int i = tid - first; i < 2 * k; i += count) {
                const int axis = i >= k, j = i - axis * k;
                const uint32_t key = (axis ? keys1 : keys0)[j];
                const int tx = (int)(key & 0x3ffu), ty = (int)((key >> 10) & 0x3ffu);
                const int px = tx * L.tile + (int)((key >> 20) & 31u), py = ty * L.tile + (int)((key >> 25) & 31u);
                const int t = ty * L.tw + tx;
                const float4 J = __ldg((axis ? jcl1 : jcl0) + t);
                VsLzTaps taps;
                uint32_t tb;
                if (patch) {
                    int ix, iy, ix0, iy0;
                    float rx0, ry0;
                    vs_lz_pos((float)px, (float)py, Pg[0], Pg[1], Pg[2], Pg[3], ix, iy, taps.rx, taps.ry);
                    vs_lz_pos((float)px, (float)py, P[0], P[1], P[2], P[3], ix0, iy0, rx0, ry0);
                    tb = __ldcg(tbs + axis * g.max_tiles + t);
                    if (ix == ix0 && iy == iy0) {
                        const uint4 q = __ldcg(patch + axis * g.max_tiles + t);
                        taps.row[0] = q.x; taps.row[1] = q.y; taps.row[2] = q.z; taps.row[3] = q.w;
                    } else {
                        vs_lz_load(kimg, L.w, L.h, L.pitch, ix, iy, taps.row);
                    }
                } else {
                    vs_lz_fetch(kimg, L.w, L.h, L.pitch, (float)px, (float)py, Pg[0], Pg[1], Pg[2], Pg[3], taps);
                    tb = __ldg(timg + (size_t)min(py, L.h - 1) * L.pitch + min(px, L.w - 1));
                }
                const float r = __fsub_rn((float)tb, vs_lz_eval(taps));
                b[0] += (double)__fmul_rn(J.x, r);
                b[1] += (double)__fmul_rn(J.y, r);
                if (i < k) b[2] += (double)__fmul_rn(J.z, r);
                else       b[3] += (double)__fmul_rn(J.w, r);
            }
        };
        // first iteration: residuals of the warp-diff pass (bit-identical: same samples, same transform)
        auto gather_cached = [&](int first, int count, double* b) {
#pragma unroll 2
            for (int i = tid - first; i < 2 * k; i += count) {
                const int axis = i >= k, j = i - axis * k;
                const uint32_t key = (axis ? keys1 : keys0)[j];
                const int t = (int)((key >> 10) & 0x3ffu) * L.tw + (int)(key & 0x3ffu);
                const float4 J = __ldg((axis ? jcl1 : jcl0) + t);
                const float r = __ldcg(res + axis * g.max_tiles + t);
                b[0] += (double)__fmul_rn(J.x, r);
                b[1] += (double)__fmul_rn(J.y, r);
                if (i < k) b[2] += (double)__fmul_rn(J.z, r);
                else       b[3] += (double)__fmul_rn(J.w, r);
            }
        };
        int iters = 0;
        int flag = FLAG_CONTINUE;
        for (int iter = 0; iter < p_max_iters; iter++) {
            iters++;
            double b[4] = {0, 0, 0, 0};
            if (iter > 0) gather(0, SOLVE_THREADS, b);
            else if (tid >= 32) {
                if (res) gather_cached(32, SOLVE_THREADS - 32, b);
                else gather(32, SOLVE_THREADS - 32, b);
            }
            VS_CLK(iter > 0 ? 6 : 3);
            double tot[4];
            block_reduce<4, NWARPS, 4>(b, sh.red4, tot);
            if (tid == 0) {
                double bb[4], dt[4];
                for (int c = 0; c < 4; c++) bb[c] = tot[c] * 0.5;
                for (int r = 0; r < 4; r++) {
                    double s = 0;
                    for (int c = 0; c < 4; c++) s += sh.Hinv[r * 4 + c] * bb[c];
                    dt[r] = s;
                }
                const double scale = 1.0 / L.w;
                double delta[4] = {dt[0] * scale, dt[1] * scale, dt[2], dt[3]};
                double Tn[4];
                vs_tf_compose(delta, sh.T, Tn);
                sh.T[0] = Tn[0]; sh.T[1] = Tn[1]; sh.T[2] = Tn[2]; sh.T[3] = Tn[3];
                const double cx = L.w * 0.5, cy = L.h * 0.5;
                const double xr = (double)((float)L.w - 1.f), yb = (double)((float)L.h - 1.f);
                const double cr[4][2] = {{0.0, 0.0}, {xr, 0.0}, {0.0, yb}, {xr, yb}};
                double d12 = 0.0;
                for (int c = 0; c < 4; c++) {
                    double c2[2];
                    vs_tf_warp_center(sh.T, cr[c][0], cr[c][1], cx, cy, c2);
                    d12 = fmax(d12, dist2d(c2, sh.c1[c]));
                    sh.c1[c][0] = c2[0]; sh.c1[c][1] = c2[1];
                }
                int f = FLAG_CONTINUE;
                if (d12 < p_threshold) f = FLAG_CONVERGED;
                else if (iter >= p_max_iters - 1) f = FLAG_FAIL;
                sh.flag = f;
            }
            __syncthreads();
            VS_CLK(4);
            flag = sh.flag;
            if (flag != FLAG_CONTINUE) break;
        }

        if (tid == 0) {
            if (a.out_iters) a.out_iters[(size_t)pair * g.levels + lvl] = iters;
            if (flag == FLAG_FAIL) {
                sh.status = 0;
            } else {
                double d01 = 0.0;
                for (int c = 0; c < 4; c++) d01 = fmax(d01, dist2d(sh.c0[c], sh.c1[c]));
                if (d01 > p_max_displacement) sh.status = 0;
                else if (lvl > 0) { sh.T[2] *= 2.0; sh.T[3] *= 2.0; }
            }
        }
        __syncthreads();
        if (sh.status == 0) break;
    }

    if (tid == 0) {
        double T[4] = {sh.T[0], sh.T[1], sh.T[2], sh.T[3]};
        if (sh.status == 1 && pr.invert) {
            double Ti[4];
            vs_tf_inverse(T, Ti);
            T[0] = Ti[0]; T[1] = Ti[1]; T[2] = Ti[2]; T[3] = Ti[3];
        }
        for (int c = 0; c < 4; c++) a.out_T[(size_t)pair * 4 + c] = T[c];
        a.out_status[pair] = sh.status;
        if (a.dbg_clock) for (int c = 0; c < 7; c++) a.dbg_clock[(size_t)pair * VS_CLK_STRIDE + c] = clk[c];
    }
}

// debug tap: one warp per matrix, quad form on all lanes and the serial form on lane 0
__global__ void __launch_bounds__(32)
k_debug_invert4(const double* __restrict__ H, double* __restrict__ out_quad, double* __restrict__ out_serial, double* __restrict__ out_cond)
{
    const int m = blockIdx.x, lane = threadIdx.x;
    double hrow[4], hinv[4];
#pragma unroll
    for (int c = 0; c < 4; c++) hrow[c] = H[m * 16 + (lane & 3) * 4 + c];
    const double cond = vs_condition_and_invert_quad(hrow, hinv);
    if (lane < 4) {
#pragma unroll
        for (int c = 0; c < 4; c++) out_quad[m * 16 + lane * 4 + c] = hinv[c];
    }
    if (lane == 0) {
        double Hs[16], Hi[16];
        for (int i = 0; i < 16; i++) Hs[i] = H[m * 16 + i];
        vs_condition_and_invert(Hs, Hi);
        for (int i = 0; i < 16; i++) out_serial[m * 16 + i] = Hi[i];
        out_cond[m] = cond;
    }
}

}  // namespace

// ================================================================== launchers

int vsk_debug_invert4(vs_ctx* ctx, const double* d_H, int n, double* d_quad, double* d_serial, double* d_cond)
{
    if (n <= 0) return VS_OK;
    VS_LAUNCH_BEGIN(ctx, VSK_SOLVE);
    k_debug_invert4<<<n, 32, 0, ctx->stream>>>(d_H, d_quad, d_serial, d_cond);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_grad_argmax(vs_ctx* ctx, const VsDevImg& gx, const VsDevImg& gy, int tile, uint16_t* d_lmx, uint16_t* d_lmy)
{
    VS_REQUIRE(ctx, tile >= 1 && gx.w == gy.w && gx.h == gy.h, "grad_argmax: bad arguments");
    VS_REQUIRE(ctx, gx.w <= 65535 && gx.h <= 65535, "grad_argmax: coordinates must fit in u16");
    int tw = gx.w / tile, th = gy.h / tile;
    if (tw * th <= 0) return VS_OK;
    VS_LAUNCH_BEGIN(ctx, VSK_GRAD_ARGMAX);
    k_grad_argmax<<<vs_cdiv(tw * th, 8), 256, 0, ctx->stream>>>((const float*)gx.data, gx.stride, (const float*)gy.data,
                                                                gy.stride, tile, tw, th, d_lmx, d_lmy);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_sparse_jac(vs_ctx* ctx, const VsDevImg& gx, const VsDevImg& gy, const uint16_t* d_lmx, const uint16_t* d_lmy,
                   int tw, int th, float* d_jx, float* d_jy)
{
    if (tw * th <= 0) return VS_OK;
    VS_LAUNCH_BEGIN(ctx, VSK_SPARSE_JAC);
    k_sparse_jac<<<vs_cdiv(tw * th, 128), 128, 0, ctx->stream>>>((const float*)gx.data, gx.stride, (const float*)gy.data,
                                                                 gy.stride, gx.w, gx.h, d_lmx, d_lmy, tw, th, d_jx, d_jy);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_sparse_warpdiff(vs_ctx* ctx, const VsDevImg& tmpl, const VsDevImg& key, const uint16_t* d_lm, int tw, int th,
                        float A, float B, float TX, float TY, uint16_t* d_out)
{
    VS_REQUIRE(ctx, tmpl.w == key.w && tmpl.h == key.h, "warpdiff: template/keyframe size mismatch");
    if (tw * th <= 0) return VS_OK;
    VS_LAUNCH_BEGIN(ctx, VSK_WARPDIFF);
    k_sparse_warpdiff<<<vs_cdiv(tw * th, 128), 128, 0, ctx->stream>>>((const uint8_t*)tmpl.data, tmpl.stride,
                                                                      (const uint8_t*)key.data, key.stride, key.w, key.h,
                                                                      d_lm, tw * th, A, B, TX, TY, d_out);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_sparse_ica(vs_ctx* ctx, const VsDevImg& tmpl, const VsDevImg& key, const uint16_t* d_selx, int kx,
                   const uint16_t* d_sely, int ky, const float* d_jx, const float* d_jy,
                   float A, float B, float TX, float TY, double* d_out4)
{
    VS_REQUIRE(ctx, tmpl.w == key.w && tmpl.h == key.h, "ica: template/keyframe size mismatch");
    VS_LAUNCH_BEGIN(ctx, VSK_ICA);
    k_sparse_ica<<<1, 256, 0, ctx->stream>>>((const uint8_t*)tmpl.data, tmpl.stride, (const uint8_t*)key.data, key.stride,
                                             key.w, key.h, d_selx, kx, d_sely, ky, d_jx, d_jy, A, B, TX, TY, d_out4);
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_keyframe_features(vs_ctx* ctx, const VsClipGeom& g, const uint8_t* d_pyr, const int32_t* d_slots, int n_slots,
                          uint32_t* d_kp, float4* d_jac)
{
    if (n_slots <= 0) return VS_OK;
    VS_REQUIRE(ctx, n_slots <= 65535, "keyframe: too many slots in one call");
    // banded kernel: needs 16-byte aligned rows and a band's tiles in its shared accumulators
    bool banded = (reinterpret_cast<uintptr_t>(d_pyr) % 16 == 0) && g.pyr_slot_bytes % 16 == 0;
    int bands = 0;
    for (int l = 0; l < g.levels; l++) {
        banded = banded && g.lv[l].pitch % 16 == 0 && g.lv[l].img_off % 16 == 0 && g.lv[l].tw <= KF_MAX_TW;
        const int bpc = kf_bands_per_cta(g.lv[l]);
        bands += (g.lv[l].th + bpc - 1) / bpc;
    }
    VS_LAUNCH_BEGIN(ctx, VSK_KEYFRAME);
    if (banded) {
        dim3 grid(bands, n_slots);
        k_keyframe_features<<<grid, KF_THREADS, 0, ctx->stream>>>(g, d_pyr, d_slots, d_kp, d_jac);
    } else {
        dim3 grid(vs_cdiv(g.total_tiles, 8), n_slots);
        k_keyframe_features_generic<<<grid, 256, 0, ctx->stream>>>(g, d_pyr, d_slots, d_kp, d_jac);
    }
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}

int vsk_solve_pairs(vs_ctx* ctx, const VsClipGeom& g, const VsSolveArgs& a)
{
    if (a.n_pairs <= 0) return VS_OK;
    // CTA size by how many pairs are in flight (see below)
    int threads = a.n_pairs <= ctx->sm_count ? 512 : 256;
    if (a.force_threads == 256 || a.force_threads == 512 || a.force_threads == 1024) threads = a.force_threads;
    // keys (8 B per tile) always live in shared memory.  A CTA that has its SM to itself adds the chunk masks and ranks
    // of the list-free selection (1 B per tile).  Three CTAs per SM keep candidate position lists (another 8 B per
    // tile): next to the keys when the total stays small, else in the unused tail of the key arrays and, for the
    // largest level, in the caller's global scratch.
    const size_t key_bytes = (size_t)2 * g.max_tiles * sizeof(uint32_t);
    const size_t pos_bytes = (size_t)4 * g.max_tiles * sizeof(uint16_t);
    const size_t sel_bytes = (size_t)2 * 4 * sel_chunks(g.max_tiles) * sizeof(uint32_t);
    VS_REQUIRE(ctx, (uint32_t)g.max_tiles <= vs_sel::KEY_TILE_MASK, "solve: more tiles per level than a key's tile index holds");
    // a level whose keys do not fit shared memory (8K) keeps them in the caller's global scratch; the masks of the
    // list-free selection stay on chip
    bool keys_global = key_bytes + sel_bytes > 220 * 1024;
    VS_REQUIRE(ctx, !keys_global || a.key_scratch, "solve: level too large for the shared-memory selection and no key scratch given");
    VS_REQUIRE(ctx, sel_bytes <= 200 * 1024, "solve: level too large for the on-chip selection masks");
    // 4K-class levels (keys beyond 100 KB: one CTA per SM either way): a launch of several pairs runs 1024 threads per pair
    // with the keys in global memory, which leaves the SM's 256 KB to L1 for the gathers (119 pairs of a 4K video: 1.79 ms
    // with 512 threads and the keys in shared memory, 1.12 with 1024 threads, 1.03 with the keys in global memory as well);
    // a single pair (the per-frame API) is quicker the old way (0.70 against 0.74 ms).  At 1080p the keys stay on chip
    // (three CTAs per SM: 0.79 against 0.86 ms).
    // More pairs than SMs: three 256-thread CTAs per SM with keys and candidate lists in global memory, so that every pair
    // is resident (the keys of ONE 4K pair would fill an SM's shared memory).
    if (keys_global) threads = 1024;
    else if (a.key_scratch && a.pos_scratch && key_bytes > VS_SOLVE_BIG_KEYS && a.n_pairs >= 8) {
        keys_global = true;
        threads = a.n_pairs <= ctx->sm_count ? 1024 : 256;
    }
    size_t smem = keys_global ? 0 : key_bytes;
    VsSolveArgs args = a;
    if (!keys_global) args.key_scratch = nullptr;
    // (Measured: keeping the lists of the shorter rounds in a shared-memory part next to the keys makes a round cheaper,
    // 9.5k -> 8.3k cycles, but every KB of shared memory is a KB of L1 taken from the gathers, which lose more: mean
    // 0.67 -> 0.97 ms per pair at 68 KB per CTA.  The keys stay alone in shared memory unless everything fits.)
    if (threads != 256) {
        smem += sel_bytes;
        args.pos_scratch = nullptr;
    } else if (key_bytes + pos_bytes <= 72 * 1024 || !a.pos_scratch) {
        VS_REQUIRE(ctx, key_bytes + pos_bytes <= 220 * 1024, "solve: level too large and no selection scratch given");
        smem += pos_bytes;
        args.pos_scratch = nullptr;
    }
    // Measured on B200 (1080p, 299 pairs in flight), all rejected (profiles/r01_*): CTAs of 320 / 384 threads (registers
    // capped at 64 / 56: 0.88 / 1.09 ms mean per pair against 0.85), register software pipelining of the gathers (1.00 ms),
    // L2 prefetch of the samples 1 / 2 / 4 iterations ahead (0.91 / 1.00 / 1.15 ms).  More loads in flight make it slower: the gathers
    // are bound by the rate of random 32-byte sector reads from DRAM (L2 hit rate 27 %), not by their latency.
    // CTA size by how many pairs are in flight: every pair must be resident at once (a second wave would double the
    // time: the kernel lasts as long as its slowest pair), and within that the largest CTA wins because the gathers
    // of a pair are independent.  256 x 3 per SM (444 pairs), 512 x 1 (148 pairs, registers uncapped), 1024 x 1 for
    // the big shared-memory footprints of 4K where only one CTA fits anyway.
    // shared-memory carve-out: just enough for the CTAs that must be resident (the rest of the 256 KB is L1 for the gathers;
    // B200, 1080p, 299 pairs: 0.857 ms with the minimal carve-out, 0.88 with the driver's default, 0.95 at 72 %, 1.43 at 86 %)
#define VS_SOLVE_LAUNCH(NT, MINB)                                                                                                 \
    do {                                                                                                                          \
        VS_CUDA(ctx, cudaFuncSetAttribute(k_solve_pairs<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        cudaFuncAttributes fa;                                                                                                    \
        VS_CUDA(ctx, cudaFuncGetAttributes(&fa, k_solve_pairs<NT, MINB>));                                                        \
        const size_t need = (size_t)MINB * (smem + fa.sharedSizeBytes + 1024);                                                    \
        const int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));                                                      \
        VS_CUDA(ctx, cudaFuncSetAttribute(k_solve_pairs<NT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,                \
                                          pct > 100 ? 100 : pct));                                                                \
        VS_LAUNCH_BEGIN(ctx, VSK_SOLVE);                                                                                          \
        k_solve_pairs<NT, MINB><<<a.n_pairs, NT, smem, ctx->stream>>>(g, args);                                                   \
    } while (0)
    // (1024 threads for a single pair — the per-frame API — were measured too: 980 against 1046-1078 frames/s)
    if (threads == 1024) VS_SOLVE_LAUNCH(1024, 1);
    else if (threads == 512) VS_SOLVE_LAUNCH(512, 1);
    else VS_SOLVE_LAUNCH(256, 3);
#undef VS_SOLVE_LAUNCH
    VS_LAUNCH_CHECK(ctx);
    return VS_OK;
}
