// vs_internal.h — device-pointer launchers behind the C ABI.  Everything here takes
// DEVICE pointers and only enqueues work on ctx->stream.
#pragma once
#include "vs_common.cuh"

// Dense image descriptor (device memory).  stride/batch_stride in elements.
struct VsDevImg {
    void* data;
    int w, h;
    int64_t stride;
    int batch;
    int64_t batch_stride;
};

static inline VsDevImg vs_dev_img(const vs_img* im)
{
    VsDevImg d;
    d.data = im->data; d.w = im->width; d.h = im->height; d.stride = im->stride;
    d.batch = im->batch < 1 ? 1 : im->batch; d.batch_stride = im->batch_stride;
    return d;
}

// ---- dense kernels (vs_kernels_dense.cu)
int vsk_bgr2gray(vs_ctx*, const VsDevImg& bgr, const VsDevImg& gray);
int vsk_pyr_down(vs_ctx*, const VsDevImg& in, const VsDevImg& out);
// fused BGR -> gray + first pyramid level (one pass over the frame); *fused = false when the geometry needs the two
// separate kernels (width not a multiple of 16, unaligned rows, level 1 not the floor half of level 0)
int vsk_ingest_bgr_gray_l1(vs_ctx*, const VsDevImg& bgr, const VsDevImg& gray0, const VsDevImg& gray1, bool* fused);
int vsk_grad_xy(vs_ctx*, const VsDevImg& in, const VsDevImg& gx, const VsDevImg& gy);
int vsk_image_warp(vs_ctx*, const VsDevImg& in, const float* d_params4, const VsDevImg& out);

// inverse-map coefficients of one BGR warp, computed on the host exactly like
// cv::warpAffine does (f64 inverse of the forward 2x3 matrix)
struct VsWarpCoef {
    double i00, i01, i02, i10, i11, i12;
};
void vs_warp_coef_from_forward(const double* M6, VsWarpCoef* out);
// imgproc.cpp:458-466 — centre-based similarity -> forward 2x3 matrix
void vs_forward_matrix_from_transform(const double* T4, int cols, int rows, double* M6);
// d_coef: device array of `src.batch` VsWarpCoef
// d_tab: optional scratch of vs_warp_rows_tab_ints(dst.w, dst.h) int32 per image; with it the two bilinear modes (constant
// border) run the row-group kernel whenever the source can be described by a tensor map (vs_warp_rows_usable)
int vsk_bgr_warp(vs_ctx*, const VsDevImg& src, const VsWarpCoef* d_coef, const VsDevImg& dst,
                 int dst_x0, int dst_y0, int mode, int border, int32_t* d_tab = nullptr);
bool vs_warp_rows_usable(const VsDevImg& src, int mode, int border);
// same, but image b of the batch reads slot d_slots[b] of `src` (src.batch_stride apart)
int vsk_bgr_warp_slots(vs_ctx*, const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef,
                       const VsDevImg& dst, int dst_x0, int dst_y0, int mode, int border);

// planar form for NV12 frames: src / dst are 1-channel (Y) or 2-channel (interleaved UV) u8 images, widths in pixels,
// strides in bytes; cv-exact bilinear, constant border
int vsk_plane_warp_slots(vs_ctx*, const VsDevImg& src, int channels, const int32_t* d_slots, const VsWarpCoef* d_coef,
                         const VsDevImg& dst, int dst_x0, int dst_y0);

// row-group form (the production kernel): tensor map with box {120, 28, 1} (160 pixels x 28 rows) over the same view;
// d_tab is scratch for the per-launch fixed-point tables, vs_warp_rows_tab_ints(dst.w, dst.h) int32 per image
int vsk_bgr_warp_slots_rows(vs_ctx*, const void* tensor_map, const VsDevImg& src, const int32_t* d_slots,
                            const VsWarpCoef* d_coef, const VsDevImg& dst, int dst_x0, int dst_y0, int32_t* d_tab,
                            int mode = VS_WARP_CV_EXACT_BILINEAR);
constexpr int VS_WARP_ROWS_BOX_WORDS = 120, VS_WARP_ROWS_BOX_ROWS = 28, VS_WARP_ROWS_TILE_W = 128, VS_WARP_ROWS_TILE_H = 24;
constexpr int VS_WARP_LZ_BOX_ROWS = 32;     // the Lanczos-2 warp reads one row above and two below: a taller box
static inline size_t vs_warp_rows_tab_ints(int dw, int dh)
{
    const size_t tx = (dw + VS_WARP_ROWS_TILE_W - 1) / VS_WARP_ROWS_TILE_W, ty = (dh + VS_WARP_ROWS_TILE_H - 1) / VS_WARP_ROWS_TILE_H;
    return 2 * tx * VS_WARP_ROWS_TILE_W + 2 * ty * VS_WARP_ROWS_TILE_H + 4 * tx * ty;
}
// Lanczos-2 mode: row-group kernel on a TMA box of VS_WARP_LZ_BOX_ROWS rows (tensor_map) or direct loads (null)
int vsk_bgr_warp_lz(vs_ctx*, const void* tensor_map, const VsDevImg& src, const int32_t* d_slots, const VsWarpCoef* d_coef,
                    const VsDevImg& dst, int dst_x0, int dst_y0, int border, int32_t* d_tab);
// cuTensorMapEncodeTiled resolved through the runtime (PFN_cuTensorMapEncodeTiled), or null
void* vs_tensor_map_encoder();

// ---- sparse kernels (vs_kernels_sparse.cu)
int vsk_grad_argmax(vs_ctx*, const VsDevImg& gx, const VsDevImg& gy, int tile,
                    uint16_t* d_lmx, uint16_t* d_lmy);
int vsk_sparse_jac(vs_ctx*, const VsDevImg& gx, const VsDevImg& gy, const uint16_t* d_lmx,
                   const uint16_t* d_lmy, int tw, int th, float* d_jx, float* d_jy);
int vsk_sparse_warpdiff(vs_ctx*, const VsDevImg& tmpl, const VsDevImg& key, const uint16_t* d_lm,
                        int tw, int th, float A, float B, float TX, float TY, uint16_t* d_out);
int vsk_sparse_ica(vs_ctx*, const VsDevImg& tmpl, const VsDevImg& key, const uint16_t* d_selx, int kx,
                   const uint16_t* d_sely, int ky, const float* d_jx, const float* d_jy,
                   float A, float B, float TX, float TY, double* d_out4);

// ---- fused clip pipeline (vs_clip.cu uses these from vs_kernels_sparse.cu)
struct VsLevel {
    int w, h, pitch;          // gray level geometry; pitch in bytes
    int tile, tw, th, ntiles;
    uint32_t img_off;         // byte offset of this level inside a slot's pyramid block
    uint32_t tile_off;        // tile offset of this level inside a per-axis feature array
};

constexpr int VS_CLK_STRIDE = 8 + 8 * VS_MAX_LEVELS;   // solver debug clocks per pair: 8 totals, then 8 per level

struct VsClipGeom {
    int levels;
    int total_tiles;          // sum of ntiles over levels
    int max_tiles;            // max ntiles over levels
    size_t pyr_slot_bytes;    // bytes of one slot's pyramid block
    VsLevel lv[VS_MAX_LEVELS];
};

// one parameter set of a batched sweep (vs_clip_align_sweep): job j of the launch aligns pair j % sweep_pairs with
// parameter set j / sweep_pairs; scratch and outputs are indexed by job
struct VsSweepSet {
    double threshold, max_displacement;
    float fraction;
    int32_t max_iters;
    int32_t use_seed;         // take the phase-correlation seed of the pair (init_T), else start from the identity
    int32_t pad;
};

struct VsSolveArgs {
    const uint8_t* pyr;       // slot-major pyramid store
    const uint32_t* kp;       // [slot][axis][total_tiles] packed (y<<16 | x)
    const float4* jac;        // [slot][axis][total_tiles]
    const vs_pair* pairs;     // device array
    int n_pairs;
    double threshold;
    float fraction;
    int max_iters;
    double max_displacement;
    double* out_T;            // 4 per pair
    int32_t* out_status;      // 1 per pair
    int32_t* out_iters;       // levels per pair
    uint16_t* dbg_warpdiff;   // [pair][axis][total_tiles] or null
    uint32_t* dbg_order;      // [pair][axis][total_tiles] or null
    int32_t* dbg_count;       // [pair][axis][levels] or null
    long long* dbg_clock;     // [pair][VS_CLK_STRIDE] cycles per phase, totals then per level (debug taps only; zeroed by the caller), or null
    uint16_t* pos_scratch;    // [pair][4][max_tiles] selection scratch in global memory, or null (shared memory)
    float* res_scratch;       // [pair][2][max_tiles] signed residual of every tile from the warp-diff pass, or null
    uint4* patch_scratch = nullptr;   // [pair][2][max_tiles] the 4x4 keyframe window of every tile as gathered by the warp-diff pass, or null
    uint8_t* tb_scratch = nullptr;    // [pair][2][max_tiles] the template byte of every tile (with patch_scratch)
    uint32_t* key_scratch = nullptr;  // [pair][2][max_tiles] the selection keys when a level does not fit shared memory, or null
    int force_threads;        // 0 = CTA size by pair count; 256 when several launches must be resident together
    const double* init_T = nullptr;   // [pair][2] initial (TX, TY) at the coarsest level (phase-correlation seed), or null
    const VsSweepSet* sweep = nullptr;   // device array of parameter sets, or null: n_pairs = sets * sweep_pairs jobs
    int sweep_pairs = 0;
};

// ---- phase-correlation initialiser (vs_phasecorr.cu; alignment.cpp:225-229, 369-388)
struct VsPhasePlan {
    int w = 0, h = 0, pitch = 0;      // the level-2 image
    int M = 0, N = 0, Kh = 0;         // DFT size (getOptimalDFTSize of h, w) and N/2+1
    double* d_tw = nullptr;           // N + M twiddles (cos, -sin), interleaved
    double* d_rows = nullptr;         // [slot][h][Kh] complex: row transforms
    double* d_spec = nullptr;         // [slot][M][Kh] complex: spectrum of the padded image
    double* d_cross = nullptr;        // [pair][M][Kh] complex: normalised cross-power spectrum
    double* d_inv = nullptr;          // [pair][M][Kh] complex: after the inverse column pass
    double* d_surf = nullptr;         // [pair][M][N]: correlation surface (unshifted)
};
int vs_optimal_dft_size(int n);
void vs_phase_twiddles(int n, double* out2);
// spectra of the listed slots (images at d_img_base + slot * slot_bytes, rows plan.pitch bytes apart)
int vsk_phase_forward(vs_ctx*, const VsPhasePlan& p, const uint8_t* d_img_base, size_t slot_bytes,
                      const int32_t* d_slots, int nslots);
// d_phase [n][3] = shift x, shift y, response (or null); d_init [n][2] = the seed (TX, TY) of the solver (or null)
int vsk_phase_pairs(vs_ctx*, const VsPhasePlan& p, const vs_pair* d_pairs, int n, double threshold, float scale,
                    double* d_phase, double* d_init);

int vsk_keyframe_features(vs_ctx*, const VsClipGeom& g, const uint8_t* d_pyr, const int32_t* d_slots,
                          int n_slots, uint32_t* d_kp, float4* d_jac);
// selection keys of a level beyond this many bytes (2 axes x 4 B per tile: 4K-class clips) may live in VsSolveArgs::key_scratch
constexpr size_t VS_SOLVE_BIG_KEYS = 100 * 1024;
int vsk_solve_pairs(vs_ctx*, const VsClipGeom& g, const VsSolveArgs& a);
int vsk_debug_invert4(vs_ctx*, const double* d_H, int n, double* d_quad, double* d_serial, double* d_cond);
