// vs_phasecorr.cu — the phase-correlation initialiser of VideoAligner::AlignNextFrame
// (alignment.cpp:225-229, 369-388: cv::phaseCorrelate of the level-2 images of the previous and the current frame,
// accepted when its response exceeds phase_correlate_threshold).
//
// cv::phaseCorrelate (OpenCV imgproc/src/phasecorr.cpp — not under /root/reference) is restated as a direct,
// separable DFT in f64:  M, N = getOptimalDFTSize(rows), (cols), zero padding on the right / bottom;
// P = F1 conj(F2); C = P |P| / (|P|^2 + FLT_EPSILON); R = unnormalised inverse DFT of C; fftShift; first maximum;
// 5x5 weighted centroid clamped to the array; response = window sum / (M N); shift = (N/2, M/2) - centroid.
// Every output element is one thread's serial sum in ascending index order (no FMA, no tree reduction), which is the
// canonical order of the CPU restatement the parity tests compare with: results are bit-identical, and the solver
// that starts from them keeps its bit-exact keypoint selections.  Sizes on this path are 5-smooth but small
// (480x270 at 1080p, 960x540 at 4K): four dense passes of M x N x (M + N) / 2 multiply-adds on the FP64 pipe,
// twiddles and the shared operand of a pass in shared memory, 2-4 outputs per thread so that the shared-memory
// wavefronts and the FP64 issue rate balance.  Default off upstream, so not on the benchmarked path.
#include "vs_internal.h"

#include <float.h>
#include <math.h>

namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_FWD_ROWS = 8;     // image rows per CTA of the forward row pass (they share every twiddle load)
constexpr int PC_INV_ROWS = 4;     // surface rows per CTA of the inverse row pass
constexpr int PC_COL_M = 4;        // outputs per thread of a column pass (they share every load of the input column)

// forward row pass: F[slot][r][k] = sum_{n<w} x[r][n] tw_N[(k n) mod N], k < N/2+1 (real input, Hermitian half)
__global__ void __launch_bounds__(PC_THREADS)
k_pc_rows_fwd(const uint8_t* __restrict__ img_base, size_t slot_bytes, int pitch, int w, int h,
              const int32_t* __restrict__ slots, const double2* __restrict__ twN, int N, int Kh,
              double2* __restrict__ F)
{
    extern __shared__ double2 pc_smem[];
    double2* const tw = pc_smem;
    double* const x = reinterpret_cast<double*>(pc_smem + N);          // [PC_FWD_ROWS][w]
    const int slot = slots[blockIdx.y];
    const int r0 = blockIdx.x * PC_FWD_ROWS;
    const uint8_t* img = img_base + (size_t)slot * slot_bytes;
    for (int j = threadIdx.x; j < N; j += PC_THREADS) tw[j] = twN[j];
    for (int i = threadIdx.x; i < PC_FWD_ROWS * w; i += PC_THREADS) {
        const int rr = i / w, n = i - rr * w;
        const int r = min(r0 + rr, h - 1);                              // rows past the image repeat the last one (not stored)
        x[i] = (double)img[(size_t)r * pitch + n];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kh; k += PC_THREADS) {
        double re[PC_FWD_ROWS], im[PC_FWD_ROWS];
#pragma unroll
        for (int q = 0; q < PC_FWD_ROWS; q++) { re[q] = 0.0; im[q] = 0.0; }
        int j = 0;
        for (int n = 0; n < w; n++) {
            const double2 t = tw[j];
#pragma unroll
            for (int q = 0; q < PC_FWD_ROWS; q++) {
                const double v = x[q * w + n];
                re[q] = re[q] + v * t.x;
                im[q] = im[q] + v * t.y;
            }
            j += k; if (j >= N) j -= N;
        }
#pragma unroll
        for (int q = 0; q < PC_FWD_ROWS; q++)
            if (r0 + q < h) F[((size_t)slot * h + r0 + q) * Kh + k] = make_double2(re[q], im[q]);
    }
}

// column pass: out[b][m][k] = sum_{r<rows} in[b][r][k] tw_M[(m r) mod M]   (CONJ: conjugate twiddles, the inverse)
// in / out are indexed by slots[b] when `slots` is given (forward: per frame), else by b (inverse: per pair)
template <bool CONJ>
__global__ void __launch_bounds__(PC_THREADS)
k_pc_cols(const double2* __restrict__ in, size_t in_stride, int rows, const int32_t* __restrict__ slots,
          const double2* __restrict__ twM, int M, int Kh, double2* __restrict__ out, size_t out_stride)
{
    extern __shared__ double2 pc_smem[];
    double2* const tw = pc_smem;
    for (int j = threadIdx.x; j < M; j += PC_THREADS) tw[j] = twM[j];
    __syncthreads();
    const int b = slots ? slots[blockIdx.z] : blockIdx.z;
    const int k = blockIdx.x * 32 + (threadIdx.x & 31);
    const int m0 = (blockIdx.y * (PC_THREADS / 32) + (threadIdx.x >> 5)) * PC_COL_M;
    if (k >= Kh || m0 >= M) return;
    const double2* col = in + (size_t)b * in_stride + k;
    double re[PC_COL_M], im[PC_COL_M];
    int j[PC_COL_M], step[PC_COL_M];
#pragma unroll
    for (int q = 0; q < PC_COL_M; q++) { re[q] = 0.0; im[q] = 0.0; j[q] = 0; step[q] = min(m0 + q, M - 1); }
    for (int r = 0; r < rows; r++) {
        const double2 a = __ldg(col + (size_t)r * Kh);
#pragma unroll
        for (int q = 0; q < PC_COL_M; q++) {
            const double2 t = tw[j[q]];
            if (CONJ) {
                const double t1 = a.x * t.x, t2 = a.y * t.y, t3 = a.y * t.x, t4 = a.x * t.y;
                re[q] = re[q] + (t1 + t2);
                im[q] = im[q] + (t3 - t4);
            } else {
                const double t1 = a.x * t.x, t2 = a.y * t.y, t3 = a.x * t.y, t4 = a.y * t.x;
                re[q] = re[q] + (t1 - t2);
                im[q] = im[q] + (t3 + t4);
            }
            j[q] += step[q]; if (j[q] >= M) j[q] -= M;
        }
    }
#pragma unroll
    for (int q = 0; q < PC_COL_M; q++)
        if (m0 + q < M) out[(size_t)b * out_stride + (size_t)(m0 + q) * Kh + k] = make_double2(re[q], im[q]);
}

// cross-power spectrum of a pair: previous frame x conj(current frame), normalised (mulSpectrums conjB, magSpectrums,
// divSpectrums with OpenCV's FLT_EPSILON guard)
__global__ void __launch_bounds__(PC_THREADS)
k_pc_cross(const double2* __restrict__ spec, size_t spec_stride, const vs_pair* __restrict__ pairs, size_t count,
           double2* __restrict__ C)
{
    const size_t i = (size_t)blockIdx.x * PC_THREADS + threadIdx.x;
    if (i >= count) return;
    const vs_pair pr = pairs[blockIdx.y];
    // template = even frame, keyframe = odd frame; invert <=> the current frame is the template (alignment.cpp:690-693)
    const int prev = pr.invert ? pr.keyframe_slot : pr.template_slot;
    const int curr = pr.invert ? pr.template_slot : pr.keyframe_slot;
    const double2 a = spec[(size_t)prev * spec_stride + i], b = spec[(size_t)curr * spec_stride + i];
    const double pr_ = a.x * b.x + a.y * b.y, pi_ = a.y * b.x - a.x * b.y;
    const double mag = sqrt(pr_ * pr_ + pi_ * pi_);
    const double den = mag * mag + (double)FLT_EPSILON;
    C[(size_t)blockIdx.y * count + i] = make_double2((pr_ * mag) / den, (pi_ * mag) / den);
}

// inverse row pass to the real correlation surface: R[p][r][n] = Re D[0] + sum_{k=1}^{(N-1)/2} 2 Re(D[k] conj tw[(k n) mod N])
// (+ the Nyquist term for even N)
__global__ void __launch_bounds__(PC_THREADS)
k_pc_rows_inv(const double2* __restrict__ D, const double2* __restrict__ twN, int M, int N, int Kh, double* __restrict__ R)
{
    extern __shared__ double2 pc_smem[];
    double2* const tw = pc_smem;
    double2* const d = pc_smem + N;                                     // [PC_INV_ROWS][Kh]
    const int r0 = blockIdx.x * PC_INV_ROWS;
    const size_t pair = blockIdx.y;
    for (int j = threadIdx.x; j < N; j += PC_THREADS) tw[j] = twN[j];
    for (int i = threadIdx.x; i < PC_INV_ROWS * Kh; i += PC_THREADS) {
        const int rr = i / Kh, k = i - rr * Kh;
        d[i] = D[(pair * M + min(r0 + rr, M - 1)) * Kh + k];
    }
    __syncthreads();
    const int kfull = (N - 1) / 2;
    for (int n = threadIdx.x; n < N; n += PC_THREADS) {
        double acc[PC_INV_ROWS];
#pragma unroll
        for (int q = 0; q < PC_INV_ROWS; q++) acc[q] = d[q * Kh].x;
        int j = 0;
        for (int k = 1; k <= kfull; k++) {
            j += n; if (j >= N) j -= N;
            const double2 t = tw[j];
#pragma unroll
            for (int q = 0; q < PC_INV_ROWS; q++) {
                const double2 v = d[q * Kh + k];
                const double s = v.x * t.x + v.y * t.y;
                acc[q] = acc[q] + 2.0 * s;
            }
        }
        if ((N & 1) == 0) {
            j += n; if (j >= N) j -= N;
#pragma unroll
            for (int q = 0; q < PC_INV_ROWS; q++) acc[q] = acc[q] + d[q * Kh + N / 2].x * tw[j].x;
        }
#pragma unroll
        for (int q = 0; q < PC_INV_ROWS; q++)
            if (r0 + q < M) R[(pair * M + r0 + q) * N + n] = acc[q];
    }
}

// first maximum of the shifted surface (minMaxLoc), 5x5 weighted centroid, response, and the seed of the solver
__global__ void __launch_bounds__(1024)
k_pc_peak(const double* __restrict__ R, int M, int N, const vs_pair* __restrict__ pairs, double threshold, float scale,
          double* __restrict__ out_phase, double* __restrict__ out_init)
{
    __shared__ double s_val[32];
    __shared__ int s_pos[32];
    const size_t pair = blockIdx.x;
    const double* S = R + pair * (size_t)M * N;
    const int hy = M / 2, hx = N / 2, total = M * N;
    auto shifted = [&](int y, int x) {   // fftShift: element (r, n) is shown at ((r + M/2) % M, (n + N/2) % N)
        int r = y - hy; if (r < 0) r += M;
        int n = x - hx; if (n < 0) n += N;
        return S[(size_t)r * N + n];
    };
    double best = -INFINITY;
    int bpos = 0x7fffffff;
    for (int p = threadIdx.x; p < total; p += blockDim.x) {
        const int y = p / N, x = p - y * N;
        const double v = shifted(y, x);
        if (v > best) { best = v; bpos = p; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, o);
        const int op = __shfl_down_sync(0xffffffffu, bpos, o);
        if (ov > best || (ov == best && op < bpos)) { best = ov; bpos = op; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_pos[threadIdx.x >> 5] = bpos; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int i = 1; i < (int)(blockDim.x >> 5); i++)
        if (s_val[i] > best || (s_val[i] == best && s_pos[i] < bpos)) { best = s_val[i]; bpos = s_pos[i]; }
    const int py = bpos / N, px = bpos - py * N;
    int minr = py - 2, maxr = py + 2, minc = px - 2, maxc = px + 2;
    if (minr < 0) minr = 0;
    if (minc < 0) minc = 0;
    if (maxr > M - 1) maxr = M - 1;
    if (maxc > N - 1) maxc = N - 1;
    double cx = 0.0, cy = 0.0, sum = 0.0;
    for (int y = minr; y <= maxr; y++)
        for (int x = minc; x <= maxc; x++) {
            const double v = shifted(y, x);
            cx = cx + (double)x * v;
            cy = cy + (double)y * v;
            sum = sum + v;
        }
    const double response = sum / (double)total;
    sum = sum + DBL_EPSILON;
    cx = cx / sum; cy = cy / sum;
    const double sx = (double)N / 2.0 - cx, sy = (double)M / 2.0 - cy;
    if (out_phase) { out_phase[pair * 3] = sx; out_phase[pair * 3 + 1] = sy; out_phase[pair * 3 + 2] = response; }
    if (out_init) {
        double tx = 0.0, ty = 0.0;
        if (response > threshold) {                       // alignment.cpp:379-387
            tx = sx * scale; ty = sy * scale;
            if (!pairs[pair].invert) { tx = -tx; ty = -ty; }   // the current frame is the keyframe
        }
        out_init[pair * 2] = tx; out_init[pair * 2 + 1] = ty;
    }
}

template <typename K>
int pc_smem_attr(vs_ctx* ctx, K kernel, size_t bytes)
{
    if (bytes > 200 * 1024) return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "phase correlation: image too wide for the shared-memory tables");
    if (bytes > 48 * 1024) VS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return VS_OK;
}

}  // namespace

int vs_optimal_dft_size(int n)
{
    for (;; n++) {
        int m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return n;
    }
}

void vs_phase_twiddles(int n, double* out2)
{
    for (int j = 0; j < n; j++) {
        const double a = 2.0 * M_PI * (double)j / (double)n;
        out2[2 * j] = cos(a); out2[2 * j + 1] = -sin(a);
    }
}

int vsk_phase_forward(vs_ctx* ctx, const VsPhasePlan& p, const uint8_t* d_img_base, size_t slot_bytes,
                      const int32_t* d_slots, int nslots)
{
    if (nslots == 0) return VS_OK;
    const double2* twN = reinterpret_cast<const double2*>(p.d_tw);
    const double2* twM = twN + p.N;
    double2* F = reinterpret_cast<double2*>(p.d_rows);
    double2* G = reinterpret_cast<double2*>(p.d_spec);
    const size_t smem_a = (size_t)p.N * 16 + (size_t)PC_FWD_ROWS * p.w * 8;
    int r = pc_smem_attr(ctx, k_pc_rows_fwd, smem_a);
    if (r != VS_OK) return r;
    r = pc_smem_attr(ctx, k_pc_cols<false>, (size_t)p.M * 16);
    if (r != VS_OK) return r;
    k_pc_rows_fwd<<<dim3((p.h + PC_FWD_ROWS - 1) / PC_FWD_ROWS, nslots), PC_THREADS, smem_a, ctx->stream>>>(
        d_img_base, slot_bytes, p.pitch, p.w, p.h, d_slots, twN, p.N, p.Kh, F);
    const int mrows = (PC_THREADS / 32) * PC_COL_M;
    k_pc_cols<false><<<dim3((p.Kh + 31) / 32, (p.M + mrows - 1) / mrows, nslots), PC_THREADS, (size_t)p.M * 16, ctx->stream>>>(
        F, (size_t)p.h * p.Kh, p.h, d_slots, twM, p.M, p.Kh, G, (size_t)p.M * p.Kh);
    VS_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return VS_OK;
}

int vsk_phase_pairs(vs_ctx* ctx, const VsPhasePlan& p, const vs_pair* d_pairs, int n, double threshold, float scale,
                    double* d_phase, double* d_init)
{
    if (n == 0) return VS_OK;
    const double2* twN = reinterpret_cast<const double2*>(p.d_tw);
    const double2* twM = twN + p.N;
    const double2* G = reinterpret_cast<const double2*>(p.d_spec);
    double2* C = reinterpret_cast<double2*>(p.d_cross);
    double2* D = reinterpret_cast<double2*>(p.d_inv);
    const size_t count = (size_t)p.M * p.Kh;
    const size_t smem_d = (size_t)p.N * 16 + (size_t)PC_INV_ROWS * p.Kh * 16;
    int r = pc_smem_attr(ctx, k_pc_cols<true>, (size_t)p.M * 16);
    if (r != VS_OK) return r;
    r = pc_smem_attr(ctx, k_pc_rows_inv, smem_d);
    if (r != VS_OK) return r;
    k_pc_cross<<<dim3((unsigned)((count + PC_THREADS - 1) / PC_THREADS), n), PC_THREADS, 0, ctx->stream>>>(G, count, d_pairs, count, C);
    const int mrows = (PC_THREADS / 32) * PC_COL_M;
    k_pc_cols<true><<<dim3((p.Kh + 31) / 32, (p.M + mrows - 1) / mrows, n), PC_THREADS, (size_t)p.M * 16, ctx->stream>>>(
        C, count, p.M, nullptr, twM, p.M, p.Kh, D, count);
    k_pc_rows_inv<<<dim3((p.M + PC_INV_ROWS - 1) / PC_INV_ROWS, n), PC_THREADS, smem_d, ctx->stream>>>(D, twN, p.M, p.N, p.Kh, p.d_surf);
    k_pc_peak<<<n, 1024, 0, ctx->stream>>>(p.d_surf, p.M, p.N, d_pairs, threshold, scale, d_phase, d_init);
    VS_CUDA(ctx, cudaGetLastError());
    ctx->launches += 4;
    return VS_OK;
}
