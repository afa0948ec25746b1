// vs_phasecorr.cu — the phase-correlation initialiser of VideoAligner::AlignNextFrame
// (alignment.cpp:225-229, 369-388: cv::phaseCorrelate of the level-2 images of the previous and the current frame,
// accepted when its response exceeds phase_correlate_threshold).
//
// cv::phaseCorrelate (OpenCV imgproc/src/phasecorr.cpp — not under /root/reference) is restated as a separable
// DFT in f64, every 1-D transform in two stages (L = L1 L2, L (L1 + L2) multiply-adds instead of L^2):  M, N = getOptimalDFTSize(rows), (cols), zero padding on the right / bottom;
// P = F1 conj(F2); C = P |P| / (|P|^2 + FLT_EPSILON); R = unnormalised inverse DFT of C; fftShift; first maximum;
// 5x5 weighted centroid clamped to the array; response = window sum / (M N); shift = (N/2, M/2) - centroid.
// Every stage output is one thread's serial sum in ascending index order (no FMA, no tree reduction), which is the
// canonical order of the CPU restatement the parity tests compare with: results are bit-identical, and the solver
// that starts from them keeps its bit-exact keypoint selections.  Sizes on this path are 5-smooth and small
// (480 x 270 = (20 x 24) x (15 x 18) at 1080p, 960 x 540 at 4K): four passes (rows, columns, inverse columns, inverse
// rows), each one launch of k_pc_dft with the transforms of a CTA, both stages and the twiddles in shared memory.
// Default off upstream, so not on the benchmarked path.
#include "vs_internal.h"

#include <float.h>
#include <math.h>

#define VS_TRY(expr) do { int _r = (expr); if (_r != VS_OK) return _r; } while (0)

namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_SMEM_TARGET = 96 * 1024;    // two CTAs per SM when the transform length allows it
constexpr int PC_SMEM_LIMIT = 200 * 1024;

enum { PC_ROWS_U8 = 0, PC_COLS = 1, PC_ROWS_HERM = 2 };

// One launch = one separable pass: every CTA evaluates T 1-D transforms of length L = L1 L2 that lie next to each other
// (T image rows, or T adjacent columns so that a column pass still moves 16 T contiguous bytes per row), in two stages
// through shared memory, in exactly the order of the CPU restatement the parity tests compare with:
//     A[k1][n2]     = sum_{n1 < L1} x[L2 n1 + n2] W^(L2 n1 k1)      (terms beyond the valid length skipped)
//     B[k1][n2]     = A[k1][n2] W^(n2 k1)
//     X[k1 + L1 k2] = sum_{n2 < L2} B[k1][n2] W^(L1 n2 k2)
// every sum one thread's serial sum in ascending index order, every product and sum individually rounded (-fmad=false).
struct PcPass {
    const void* in;
    void* out;
    const int32_t* slots;      // batch index -> buffer index (per frame), or null (per pair)
    const double2* tw;         // L twiddles (cos, -sin)
    size_t in_batch, out_batch;   // elements between the buffers of two batch entries
    int L, L1, L2;
    int valid;                 // input length (zero padded to L)
    int kout;                  // outputs wanted (k < kout)
    int count;                 // transforms per batch entry (rows, or columns)
    int T;                     // transforms per CTA
    int pitch;                 // PC_ROWS_U8: bytes between image rows; otherwise Kh (complex elements between rows)
    int outw;                  // elements between output rows
};

template <int MODE, bool CONJ>
__global__ void __launch_bounds__(PC_THREADS)
k_pc_dft(const PcPass a)
{
    extern __shared__ double2 pc_smem[];
    const int L = a.L, L1 = a.L1, L2 = a.L2, T = a.T;
    double2* const tw = pc_smem;
    double2* const B = pc_smem + L;                         // [T][L2][L1] (rows) or [L2][L1][T] (columns)
    double2* const X = B + (size_t)T * L;                   // complex input; doubles for PC_ROWS_U8
    double* const Xr = reinterpret_cast<double*>(X);
    const int t0 = blockIdx.x * T;
    const int nt = min(T, a.count - t0);
    const size_t b = a.slots ? (size_t)a.slots[blockIdx.y] : (size_t)blockIdx.y;
    auto at = [&](int t, int i) { return MODE == PC_COLS ? i * T + t : t * L + i; };

    for (int j = threadIdx.x; j < L; j += PC_THREADS) tw[j] = a.tw[j];
    if (MODE == PC_ROWS_U8) {
        const uint8_t* img = static_cast<const uint8_t*>(a.in) + b * a.in_batch;
        for (int i = threadIdx.x; i < nt * a.valid; i += PC_THREADS) {
            const int t = i / a.valid, n = i - t * a.valid;
            Xr[t * L + n] = (double)img[(size_t)(t0 + t) * a.pitch + n];
        }
    } else if (MODE == PC_COLS) {
        const double2* in = static_cast<const double2*>(a.in) + b * a.in_batch + t0;
        for (int i = threadIdx.x; i < T * a.valid; i += PC_THREADS) {
            const int n = i / T, t = i - n * T;
            if (t < nt) X[i] = __ldg(in + (size_t)n * a.pitch + t);
        }
    } else {
        // the row completed by its Hermitian half: Y[k] = conj(Y[L - k]) for k > L / 2
        const double2* in = static_cast<const double2*>(a.in) + b * a.in_batch + (size_t)t0 * a.pitch;
        for (int i = threadIdx.x; i < nt * L; i += PC_THREADS) {
            const int t = i / L, k = i - t * L;
            double2 v;
            if (k < a.pitch) v = __ldg(in + (size_t)t * a.pitch + k);
            else { v = __ldg(in + (size_t)t * a.pitch + (L - k)); v.y = -v.y; }
            X[i] = v;
        }
    }
    __syncthreads();

    const int n_a = (MODE == PC_COLS ? T : nt) * L;
    for (int o = threadIdx.x; o < n_a; o += PC_THREADS) {
        int t, rem;
        if (MODE == PC_COLS) { rem = o / T; t = o - rem * T; if (t >= nt) continue; }
        else { t = o / L; rem = o - t * L; }
        const int k1 = rem / L2, n2 = rem - k1 * L2;
        const int step = L2 * k1;                            // < L
        int terms = a.valid > n2 ? (a.valid - n2 + L2 - 1) / L2 : 0;   // n = L2 n1 + n2 < valid
        terms = min(terms, L1);
        double re = 0.0, im = 0.0;
        int j = 0;
        for (int n1 = 0; n1 < terms; n1++) {
            const double2 w = tw[j];
            if (MODE == PC_ROWS_U8) {
                const double x = Xr[t * L + L2 * n1 + n2];
                re = re + x * w.x;
                im = CONJ ? im - x * w.y : im + x * w.y;
            } else {
                const double2 v = X[at(t, L2 * n1 + n2)];
                const double t1 = v.x * w.x, t2 = v.y * w.y, t3 = v.x * w.y, t4 = v.y * w.x;
                if (CONJ) { re = re + (t1 + t2); im = im + (t4 - t3); }
                else      { re = re + (t1 - t2); im = im + (t3 + t4); }
            }
            j += step; if (j >= L) j -= L;
        }
        const double2 w = tw[n2 * k1];                       // < L
        const double t1 = re * w.x, t2 = im * w.y, t3 = re * w.y, t4 = im * w.x;
        B[at(t, n2 * L1 + k1)] = CONJ ? make_double2(t1 + t2, t4 - t3) : make_double2(t1 - t2, t3 + t4);
    }
    __syncthreads();

    const int n_b = (MODE == PC_COLS ? T : nt) * a.kout;
    for (int o = threadIdx.x; o < n_b; o += PC_THREADS) {
        int t, k;
        if (MODE == PC_COLS) { k = o / T; t = o - k * T; if (t >= nt) continue; }
        else { t = o / a.kout; k = o - t * a.kout; }
        const int k2 = k / L1, k1 = k - k2 * L1;
        const int step = L1 * k2;                            // < L
        double re = 0.0, im = 0.0;
        int j = 0;
        for (int n2 = 0; n2 < L2; n2++) {
            const double2 w = tw[j];
            const double2 v = B[at(t, n2 * L1 + k1)];
            const double t1 = v.x * w.x, t2 = v.y * w.y, t3 = v.x * w.y, t4 = v.y * w.x;
            if (CONJ) { re = re + (t1 + t2); im = im + (t4 - t3); }
            else      { re = re + (t1 - t2); im = im + (t3 + t4); }
            j += step; if (j >= L) j -= L;
        }
        if (MODE == PC_ROWS_HERM)
            static_cast<double*>(a.out)[b * a.out_batch + (size_t)(t0 + t) * a.outw + k] = re;
        else if (MODE == PC_COLS)
            static_cast<double2*>(a.out)[b * a.out_batch + (size_t)k * a.outw + t0 + t] = make_double2(re, im);
        else
            static_cast<double2*>(a.out)[b * a.out_batch + (size_t)(t0 + t) * a.outw + k] = make_double2(re, im);
    }
}

// L = L1 L2 with L1 the largest divisor of L whose square does not exceed L
void pc_split(int L, int& L1, int& L2)
{
    L1 = 1;
    for (int d = 1; d * d <= L; d++)
        if (L % d == 0) L1 = d;
    L2 = L / L1;
}

template <int MODE, bool CONJ>
int pc_launch(vs_ctx* ctx, PcPass a, int batches)
{
    pc_split(a.L, a.L1, a.L2);
    const size_t per_t = (size_t)a.L * (MODE == PC_ROWS_U8 ? 24 : 32), fixed = (size_t)a.L * 16;
    if (fixed + per_t > (size_t)PC_SMEM_LIMIT)
        return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "phase correlation: image too large for the shared-memory tables");
    int T = 8;
    while (T > 1 && fixed + per_t * T > (size_t)PC_SMEM_TARGET) T >>= 1;
    a.T = T;
    const size_t smem = fixed + per_t * T;
    if (smem > 48 * 1024) VS_CUDA(ctx, cudaFuncSetAttribute(k_pc_dft<MODE, CONJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pc_dft<MODE, CONJ><<<dim3((a.count + T - 1) / T, batches), PC_THREADS, smem, ctx->stream>>>(a);
    ctx->launches += 1;
    return VS_OK;
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
    PcPass rows{};   // F[slot][r][k] = row transforms of the real image, k < N/2+1 (Hermitian half)
    rows.in = d_img_base; rows.out = p.d_rows; rows.slots = d_slots; rows.tw = twN;
    rows.in_batch = slot_bytes; rows.out_batch = (size_t)p.h * p.Kh;
    rows.L = p.N; rows.valid = p.w; rows.kout = p.Kh; rows.count = p.h; rows.pitch = p.pitch; rows.outw = p.Kh;
    VS_TRY((pc_launch<PC_ROWS_U8, false>(ctx, rows, nslots)));
    PcPass cols{};   // G[slot][m][k] = column transforms of F, the h rows zero padded to M
    cols.in = p.d_rows; cols.out = p.d_spec; cols.slots = d_slots; cols.tw = twM;
    cols.in_batch = (size_t)p.h * p.Kh; cols.out_batch = (size_t)p.M * p.Kh;
    cols.L = p.M; cols.valid = p.h; cols.kout = p.M; cols.count = p.Kh; cols.pitch = p.Kh; cols.outw = p.Kh;
    VS_TRY((pc_launch<PC_COLS, false>(ctx, cols, nslots)));
    VS_CUDA(ctx, cudaGetLastError());
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
    const size_t count = (size_t)p.M * p.Kh;
    k_pc_cross<<<dim3((unsigned)((count + PC_THREADS - 1) / PC_THREADS), n), PC_THREADS, 0, ctx->stream>>>(G, count, d_pairs, count, C);
    PcPass cols{};   // inverse along the columns (conjugate twiddles)
    cols.in = p.d_cross; cols.out = p.d_inv; cols.slots = nullptr; cols.tw = twM;
    cols.in_batch = count; cols.out_batch = count;
    cols.L = p.M; cols.valid = p.M; cols.kout = p.M; cols.count = p.Kh; cols.pitch = p.Kh; cols.outw = p.Kh;
    VS_TRY((pc_launch<PC_COLS, true>(ctx, cols, n)));
    PcPass rows{};   // inverse along the rows to the real correlation surface
    rows.in = p.d_inv; rows.out = p.d_surf; rows.slots = nullptr; rows.tw = twN;
    rows.in_batch = count; rows.out_batch = (size_t)p.M * p.N;
    rows.L = p.N; rows.valid = p.N; rows.kout = p.N; rows.count = p.M; rows.pitch = p.Kh; rows.outw = p.N;
    VS_TRY((pc_launch<PC_ROWS_HERM, true>(ctx, rows, n)));
    k_pc_peak<<<n, 1024, 0, ctx->stream>>>(p.d_surf, p.M, p.N, d_pairs, threshold, scale, d_phase, d_init);
    VS_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return VS_OK;
}
