// vs_clip.cu — device-resident frame store and the batched alignment / warp pipeline.
//
// HBM layout of a clip (all sized at creation, nothing is allocated on the hot path):
//   bgr      [capacity][h][bgr_pitch]                     interleaved BGR, pitch = align(3w,128)
//   pyr      [capacity][pyr_slot_bytes]                   gray pyramid, level l at lv[l].img_off,
//                                                         rows lv[l].pitch = align(w_l,128) bytes apart
//   kp       [capacity][2 axes][total_tiles] u32          keypoint (y<<16|x) per tile, levels concatenated
//   jac      [capacity][2 axes][total_tiles] float4       Jacobian per keypoint (reference channel order)
//   pairs / out_T / out_status / out_iters               per vs_clip_align call, max_pairs entries
//   dbg_*    optional (VS_CLIP_DEBUG_TAPS)               warpdiff and selection order per pair
//   pos_scratch / res_scratch / warp_tab                  per-pair selection lists and warp-diff residuals, per-launch warp tables
#include "vs_internal.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct vs_clip {
    // TMA descriptor of the BGR store viewed as u32 [slot][row][pitch/4], box {120 words = 160 pixels, 28 rows, 1}: the
    // cv-exact warp fetches a tile's source box with one cp.async.bulk.tensor (first member: 64-byte aligned)
    CUtensorMap bgr_map_rows;
    CUtensorMap bgr_map_lz;           // the same view with the taller box of the Lanczos-2 warp
    bool bgr_map_rows_ok = false, bgr_map_lz_ok = false;
    int32_t* d_warp_tab = nullptr;    // fixed-point column / row tables of the row-group warp, capacity images
    vs_ctx* ctx = nullptr;
    int w = 0, h = 0, capacity = 0, max_pairs = 0, flags = 0;
    vs_align_params params;
    VsClipGeom g;
    size_t bgr_pitch = 0, bgr_slot_bytes = 0;
    uint8_t* d_bgr = nullptr;
    // VS_CLIP_NV12: the Y plane of a frame IS level 0 of its pyramid (uploaded there, never rewritten); the interleaved
    // UV plane ((w/2) x (h/2) pairs) lives here.  d_bgr is not allocated.
    bool nv12 = false;
    size_t uv_pitch = 0, uv_slot_bytes = 0;
    uint8_t* d_uv = nullptr;
    uint8_t* d_pyr = nullptr;
    uint32_t* d_kp = nullptr;
    float4* d_jac = nullptr;
    vs_pair* d_pairs = nullptr;
    double* d_T = nullptr;
    int32_t* d_status = nullptr;
    int32_t* d_iters = nullptr;
    int32_t* d_slots = nullptr;       // max(capacity, max_pairs) ints
    VsWarpCoef* d_coef = nullptr;     // capacity entries
    uint16_t* d_dbg_wd = nullptr;
    uint32_t* d_dbg_order = nullptr;
    int32_t* d_dbg_count = nullptr;
    long long* d_dbg_clock = nullptr;   // [max_pairs][VS_CLK_STRIDE]
    uint16_t* d_pos_scratch = nullptr; // [max_pairs][4][max_tiles] candidate lists of the parallel selection
    uint4* d_patch_scratch = nullptr;  // [max_pairs][2][max_tiles] 4x4 keyframe windows of the warp-diff pass (patch cache of the solver)
    uint8_t* d_tb_scratch = nullptr;   // [max_pairs][2][max_tiles] template bytes of the warp-diff pass
    uint32_t* d_key_scratch = nullptr; // [max_pairs][2][max_tiles] selection keys of 4K-class and larger clips (vsk_solve_pairs decides per launch)
    float* d_res_scratch = nullptr;    // [max_pairs][2][max_tiles] warp-diff residuals reused by the first Gauss-Newton iteration
    uint8_t* d_warp_out = nullptr;    // staging for VS_MEM_HOST warps, grown on demand
    size_t warp_out_bytes = 0;
    // asynchronous transfer pipeline (vs_clip_upload_async / vs_clip_warp_to_host_async)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_upload = nullptr;
    cudaEvent_t ev_bgr_read = nullptr;   // recorded after every kernel that reads the BGR store (ingest, warps)
    bool bgr_read_recorded = false;
    static const int kOutRing = 3;
    uint8_t* d_out_ring[kOutRing] = {nullptr, nullptr, nullptr};
    size_t out_ring_bytes[kOutRing] = {0, 0, 0};
    cudaEvent_t ev_warp[kOutRing] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_down[kOutRing] = {nullptr, nullptr, nullptr};
    bool out_ring_used[kOutRing] = {false, false, false};
    VsWarpCoef* d_coef_ring = nullptr;  // kOutRing * capacity
    int32_t* d_slot_ring = nullptr;     // kOutRing * capacity
    int out_ring_next = 0;
    int last_pairs = 0;
    // asynchronous solver lanes (vs_clip_align_async / vs_clip_align_wait)
    static const int kLanes = VS_CLIP_SOLVER_LANES;
    cudaStream_t solve_stream[kLanes] = {};
    cudaEvent_t ev_ready[kLanes] = {}, ev_solved[kLanes] = {};
    double* h_T = nullptr;            // pinned, max_pairs * 4
    int32_t* h_status = nullptr;      // pinned, max_pairs
    int32_t* h_iters = nullptr;       // pinned, max_pairs * levels (only when a caller asks for iteration counts)
    VsSweepSet* d_sweep = nullptr;    // parameter sets of vs_clip_align_sweep
    size_t sweep_capacity = 0;
    int lane_base[kLanes] = {}, lane_n[kLanes] = {};
    // phase-correlation initialiser (params.phase_correlate): allocated on first use
    VsPhasePlan pc;
    bool pc_ready = false;
    std::vector<char> pc_valid;       // per slot: the spectrum matches the pyramid in the slot
    int32_t* d_pc_slots = nullptr;    // capacity
    double* d_pc_init = nullptr;      // [max_pairs][2] seed (TX, TY) of the solver
    double* d_pc_phase = nullptr;     // [max_pairs][3] shift x, shift y, response
};

namespace {

#define VS_TRY(expr) do { int _r = (expr); if (_r != VS_OK) return _r; } while (0)

template <typename T>
int dev_alloc(vs_ctx* ctx, T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) return VS_OK;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess)
        return vs_set_error(ctx, VS_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    return VS_OK;
}

void free_all(vs_clip* c)
{
    cudaFree(c->d_bgr); cudaFree(c->d_uv); cudaFree(c->d_pyr); cudaFree(c->d_kp); cudaFree(c->d_jac); cudaFree(c->d_pairs);
    cudaFree(c->d_T); cudaFree(c->d_status); cudaFree(c->d_iters); cudaFree(c->d_slots); cudaFree(c->d_coef);
    for (int i = 0; i < vs_clip::kOutRing; i++) {
        cudaFree(c->d_out_ring[i]);
        if (c->ev_warp[i]) cudaEventDestroy(c->ev_warp[i]);
        if (c->ev_down[i]) cudaEventDestroy(c->ev_down[i]);
    }
    cudaFree(c->d_coef_ring); cudaFree(c->d_slot_ring);
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
    if (c->ev_bgr_read) cudaEventDestroy(c->ev_bgr_read);
    if (c->up_stream) cudaStreamDestroy(c->up_stream);
    if (c->down_stream) cudaStreamDestroy(c->down_stream);
    cudaFree(c->d_pos_scratch); cudaFree(c->d_res_scratch); cudaFree(c->d_patch_scratch); cudaFree(c->d_tb_scratch); cudaFree(c->d_key_scratch); cudaFree(c->d_dbg_wd); cudaFree(c->d_dbg_order); cudaFree(c->d_dbg_count); cudaFree(c->d_dbg_clock); cudaFree(c->d_warp_out);
    cudaFree(c->d_warp_tab);
    cudaFree(c->d_sweep);
    cudaFree(c->pc.d_tw); cudaFree(c->pc.d_rows); cudaFree(c->pc.d_spec); cudaFree(c->pc.d_cross); cudaFree(c->pc.d_inv);
    cudaFree(c->pc.d_surf); cudaFree(c->d_pc_slots); cudaFree(c->d_pc_init); cudaFree(c->d_pc_phase);
    for (int l = 0; l < vs_clip::kLanes; l++) {
        if (c->solve_stream[l]) { cudaStreamSynchronize(c->solve_stream[l]); cudaStreamDestroy(c->solve_stream[l]); }
        if (c->ev_ready[l]) cudaEventDestroy(c->ev_ready[l]);
        if (c->ev_solved[l]) cudaEventDestroy(c->ev_solved[l]);
    }
    if (c->h_T) cudaFreeHost(c->h_T);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_iters) cudaFreeHost(c->h_iters);
}

PFN_cuTensorMapEncodeTiled tensor_map_encoder()
{
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(vs_tensor_map_encoder());
}

void build_bgr_tensor_map(vs_clip* c)
{
    c->bgr_map_rows_ok = c->bgr_map_lz_ok = false;
    // per-launch fixed-point tables of the row-group warps
    if (cudaMalloc((void**)&c->d_warp_tab, vs_warp_rows_tab_ints(c->w, c->h) * c->capacity * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        c->d_warp_tab = nullptr;
        return;
    }
    PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc || c->bgr_pitch % 16 != 0 || c->bgr_slot_bytes % 16 != 0) return;
    if ((int)(c->bgr_pitch / 4) < VS_WARP_ROWS_BOX_WORDS) return;
    const cuuint64_t dims[3] = {(cuuint64_t)(c->bgr_pitch / 4), (cuuint64_t)c->h, (cuuint64_t)c->capacity};
    const cuuint64_t strides[2] = {(cuuint64_t)c->bgr_pitch, (cuuint64_t)c->bgr_slot_bytes};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (c->h >= VS_WARP_ROWS_BOX_ROWS) {
        const cuuint32_t box_rows[3] = {(cuuint32_t)VS_WARP_ROWS_BOX_WORDS, (cuuint32_t)VS_WARP_ROWS_BOX_ROWS, 1};
        CUresult r = enc(&c->bgr_map_rows, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, c->d_bgr, dims, strides, box_rows, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->bgr_map_rows_ok = (r == CUDA_SUCCESS);
    }
    if (c->h >= VS_WARP_LZ_BOX_ROWS) {
        const cuuint32_t box_lz[3] = {(cuuint32_t)VS_WARP_ROWS_BOX_WORDS, (cuuint32_t)VS_WARP_LZ_BOX_ROWS, 1};
        CUresult r = enc(&c->bgr_map_lz, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, c->d_bgr, dims, strides, box_lz, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->bgr_map_lz_ok = (r == CUDA_SUCCESS);
    }
}

// the BGR warp of a clip: row-group kernel for the mode the stabiliser uses (cv-exact, constant border); the tiled kernel
// when the clip is too small for its TMA box; generic kernels for the other modes / borders
int clip_warp_launch(vs_clip* c, const int32_t* d_slots, const VsWarpCoef* d_coef, const VsDevImg& dst, int crop, int mode, int border, int n)
{
    if (c->nv12) {
        // dst: the Y planes of n dense NV12 frames (dst.batch_stride = frame bytes); the UV plane follows each Y plane.
        // d_coef: n coefficient sets of the Y planes, then (c->capacity entries later) those of the UV planes.
        if (mode != VS_WARP_CV_EXACT_BILINEAR || border != VS_BORDER_CONSTANT0)
            return vs_set_error(c->ctx, VS_ERR_UNSUPPORTED, "clip_warp: NV12 clips warp with the cv-exact bilinear mode and the constant border");
        const VsLevel& L0 = c->g.lv[0];
        VsDevImg ysrc{c->d_pyr + L0.img_off, c->w, c->h, (int64_t)L0.pitch, n, (int64_t)c->g.pyr_slot_bytes};
        VsDevImg usrc{c->d_uv, c->w / 2, c->h / 2, (int64_t)c->uv_pitch, n, (int64_t)c->uv_slot_bytes};
        VsDevImg udst{(uint8_t*)dst.data + (size_t)dst.w * dst.h, dst.w / 2, dst.h / 2, dst.stride, n, dst.batch_stride};
        int r = vsk_plane_warp_slots(c->ctx, ysrc, 1, d_slots, d_coef, dst, crop, crop);
        if (r == VS_OK) r = vsk_plane_warp_slots(c->ctx, usrc, 2, d_slots, d_coef + c->capacity, udst, crop / 2, crop / 2);
        return r;
    }
    VsDevImg src{c->d_bgr, c->w, c->h, (int64_t)c->bgr_pitch, n, (int64_t)c->bgr_slot_bytes};
    if ((mode == VS_WARP_CV_EXACT_BILINEAR || mode == VS_WARP_FLOAT_BILINEAR) && border == VS_BORDER_CONSTANT0 &&
        c->bgr_map_rows_ok && n <= c->capacity && dst.w <= c->w && dst.h <= c->h)
        return vsk_bgr_warp_slots_rows(c->ctx, &c->bgr_map_rows, src, d_slots, d_coef, dst, crop, crop, c->d_warp_tab, mode);
    if (mode == VS_WARP_LANCZOS2) {
        if (!c->d_warp_tab || n > c->capacity || dst.w > c->w || dst.h > c->h)
            return vs_set_error(c->ctx, VS_ERR_NOMEM, "clip_warp: no table scratch for the Lanczos-2 warp");
        return vsk_bgr_warp_lz(c->ctx, c->bgr_map_lz_ok ? &c->bgr_map_lz : nullptr, src, d_slots, d_coef, dst, crop, crop, border,
                               c->d_warp_tab);
    }
    return vsk_bgr_warp_slots(c->ctx, src, d_slots, d_coef, dst, crop, crop, mode, border);
}

// An asynchronous upload overwrites BGR frames only: it has to wait for the kernels that READ the BGR store (the ingest of
// a slot's previous frame, the warps), not for the pyramid levels, keyframe features and solves enqueued since.
void note_bgr_read(vs_clip* c)
{
    if (!c->ev_bgr_read) return;
    cudaEventRecord(c->ev_bgr_read, c->ctx->stream);
    c->bgr_read_recorded = true;
}

// bytes of one dense output frame of ow x oh pixels
size_t clip_frame_bytes(const vs_clip* c, int ow, int oh)
{
    return c->nv12 ? (size_t)ow * oh * 3 / 2 : (size_t)ow * oh * 3;
}

// Inverse-map coefficients of n corrections: [0, n) for the BGR frame / the Y plane, and for NV12 clips
// [capacity, capacity + n) for the UV plane, which is warped as a (w/2) x (h/2) image by the same similarity with half
// the translation.
void clip_warp_coefs(const vs_clip* c, const double* transforms, int n, std::vector<VsWarpCoef>& coef)
{
    coef.resize(c->nv12 ? (size_t)c->capacity + n : (size_t)n);
    for (int i = 0; i < n; i++) {
        double M[6];
        vs_forward_matrix_from_transform(transforms + 4 * i, c->w, c->h, M);
        vs_warp_coef_from_forward(M, &coef[i]);
        if (c->nv12) {
            const double* T = transforms + 4 * i;
            const double Tuv[4] = {T[0], T[1], T[2] * 0.5, T[3] * 0.5};
            vs_forward_matrix_from_transform(Tuv, c->w / 2, c->h / 2, M);
            vs_warp_coef_from_forward(M, &coef[(size_t)c->capacity + i]);
        }
    }
}

// n frames into slots [slot0, slot0 + n) on `stream`
int clip_enqueue_upload(vs_clip* c, int slot0, int n, const uint8_t* frames, int64_t row_stride, int64_t frame_stride,
                        cudaMemcpyKind kind, cudaStream_t stream)
{
    vs_ctx* ctx = c->ctx;
    if (c->nv12) {
        const VsLevel& L0 = c->g.lv[0];
        if (n > 1 && row_stride > 0 && frame_stride % row_stride == 0 && c->g.pyr_slot_bytes % (size_t)L0.pitch == 0) {
            // the Y planes of the whole range in one 3-D copy (slices = frames), the UV planes in another: a copy per
            // plane per frame costs the DMA engine a few microseconds each, a quarter of a 16-frame sub-chunk's transfer
            cudaMemcpy3DParms py = {};
            py.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(frames), (size_t)row_stride, (size_t)c->w, (size_t)(frame_stride / row_stride));
            py.dstPtr = make_cudaPitchedPtr(c->d_pyr + (size_t)slot0 * c->g.pyr_slot_bytes + L0.img_off, (size_t)L0.pitch, (size_t)c->w,
                                            c->g.pyr_slot_bytes / (size_t)L0.pitch);
            py.extent = make_cudaExtent((size_t)c->w, (size_t)c->h, (size_t)n);
            py.kind = kind;
            VS_CUDA(ctx, cudaMemcpy3DAsync(&py, stream));
            cudaMemcpy3DParms pu = {};
            pu.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(frames) + (size_t)row_stride * c->h, (size_t)row_stride, (size_t)c->w,
                                            (size_t)(frame_stride / row_stride));
            pu.dstPtr = make_cudaPitchedPtr(c->d_uv + (size_t)slot0 * c->uv_slot_bytes, c->uv_pitch, (size_t)c->w, (size_t)c->h / 2);
            pu.extent = make_cudaExtent((size_t)c->w, (size_t)c->h / 2, (size_t)n);
            pu.kind = kind;
            VS_CUDA(ctx, cudaMemcpy3DAsync(&pu, stream));
            return VS_OK;
        }
        for (int i = 0; i < n; i++) {
            const uint8_t* f = frames + (size_t)frame_stride * i;
            VS_CUDA(ctx, cudaMemcpy2DAsync(c->d_pyr + (size_t)(slot0 + i) * c->g.pyr_slot_bytes + L0.img_off, (size_t)L0.pitch,
                                           f, (size_t)row_stride, (size_t)c->w, (size_t)c->h, kind, stream));
            VS_CUDA(ctx, cudaMemcpy2DAsync(c->d_uv + (size_t)(slot0 + i) * c->uv_slot_bytes, c->uv_pitch,
                                           f + (size_t)row_stride * c->h, (size_t)row_stride, (size_t)c->w, (size_t)c->h / 2, kind, stream));
        }
        return VS_OK;
    }
    if (frame_stride == row_stride * c->h) {
        // frames are back to back: one 2-D copy for the whole range
        VS_CUDA(ctx, cudaMemcpy2DAsync(c->d_bgr + (size_t)slot0 * c->bgr_slot_bytes, c->bgr_pitch, frames, (size_t)row_stride,
                                       (size_t)c->w * 3, (size_t)c->h * n, kind, stream));
    } else {
        for (int i = 0; i < n; i++)
            VS_CUDA(ctx, cudaMemcpy2DAsync(c->d_bgr + (size_t)(slot0 + i) * c->bgr_slot_bytes, c->bgr_pitch,
                                           frames + (size_t)frame_stride * i, (size_t)row_stride, (size_t)c->w * 3, c->h,
                                           kind, stream));
    }
    return VS_OK;
}

constexpr int kPhaseLevel = 2;   // alignment.hpp:69

int phase_prepare(vs_clip* c)
{
    if (c->pc_ready) return VS_OK;
    vs_ctx* ctx = c->ctx;
    const VsClipGeom& g = c->g;
    if (g.levels <= kPhaseLevel)
        return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "phase_correlate needs pyramid level %d, this clip has %d levels", kPhaseLevel, g.levels);
    VsPhasePlan& p = c->pc;
    const VsLevel& L = g.lv[kPhaseLevel];
    p.w = L.w; p.h = L.h; p.pitch = L.pitch;
    p.M = vs_optimal_dft_size(L.h); p.N = vs_optimal_dft_size(L.w); p.Kh = p.N / 2 + 1;
    const size_t spec = (size_t)p.M * p.Kh * 2;
    int r = dev_alloc(ctx, &p.d_tw, (size_t)(p.N + p.M) * 2);
    if (r == VS_OK) r = dev_alloc(ctx, &p.d_rows, (size_t)c->capacity * p.h * p.Kh * 2);
    if (r == VS_OK) r = dev_alloc(ctx, &p.d_spec, (size_t)c->capacity * spec);
    if (r == VS_OK) r = dev_alloc(ctx, &p.d_cross, (size_t)c->max_pairs * spec);
    if (r == VS_OK) r = dev_alloc(ctx, &p.d_inv, (size_t)c->max_pairs * spec);
    if (r == VS_OK) r = dev_alloc(ctx, &p.d_surf, (size_t)c->max_pairs * p.M * p.N);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pc_slots, (size_t)c->capacity);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pc_init, (size_t)c->max_pairs * 2);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pc_phase, (size_t)c->max_pairs * 3);
    if (r != VS_OK) return r;
    std::vector<double> tw((size_t)(p.N + p.M) * 2);
    vs_phase_twiddles(p.N, tw.data());
    vs_phase_twiddles(p.M, tw.data() + (size_t)p.N * 2);
    VS_CUDA(ctx, cudaMemcpyAsync(p.d_tw, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // tw is a local
    c->pc_valid.assign(c->capacity, 0);
    c->pc_ready = true;
    return VS_OK;
}

// Enqueues, on the context stream, the spectra of the frames of `pairs` that do not have one yet and the phase
// correlation of every pair; d_pairs is the device copy of `pairs`, [base, base + n) the pairs' scratch range.
// *init receives the device array of seeds for the solver.
int phase_seed(vs_clip* c, const vs_pair* pairs, const vs_pair* d_pairs, int n, int base, const double** init)
{
    vs_ctx* ctx = c->ctx;
    VS_TRY(phase_prepare(c));
    std::vector<int32_t> todo;
    for (int i = 0; i < n; i++)
        for (int s : {pairs[i].template_slot, pairs[i].keyframe_slot})
            if (!c->pc_valid[s]) { c->pc_valid[s] = 1; todo.push_back(s); }
    if (!todo.empty()) {
        // pageable source: the copy is staged before the call returns
        VS_CUDA(ctx, cudaMemcpyAsync(c->d_pc_slots, todo.data(), todo.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        VS_TRY(vsk_phase_forward(ctx, c->pc, c->d_pyr + c->g.lv[kPhaseLevel].img_off, c->g.pyr_slot_bytes, c->d_pc_slots, (int)todo.size()));
    }
    // alignment.cpp:380: the seed is expressed at the coarsest level
    const float scale = (float)(1 << kPhaseLevel) / float(1 << c->g.levels);
    VsPhasePlan p = c->pc;
    const size_t spec = (size_t)p.M * p.Kh * 2;
    p.d_cross += (size_t)base * spec; p.d_inv += (size_t)base * spec; p.d_surf += (size_t)base * p.M * p.N;
    VS_TRY(vsk_phase_pairs(ctx, p, d_pairs, n, c->params.phase_correlate_threshold, scale,
                           c->d_pc_phase + (size_t)base * 3, c->d_pc_init + (size_t)base * 2));
    *init = c->d_pc_init + (size_t)base * 2;
    return VS_OK;
}

}  // namespace

extern "C" {

int vs_clip_create(vs_ctx* ctx, int width, int height, int capacity, int max_pairs,
                   const vs_align_params* params, int flags, vs_clip** out)
{
    if (!ctx || !out) return VS_ERR_INVALID;
    *out = nullptr;
    VS_REQUIRE(ctx, width > 0 && height > 0 && capacity > 0 && max_pairs >= 0, "clip_create: bad geometry");
    VS_REQUIRE(ctx, width <= 65535 && height <= 65535, "clip_create: keypoint coordinates must fit in u16");
    vs_align_params P;
    if (params) P = *params; else vs_align_params_default(&P);
    VS_REQUIRE(ctx, P.max_iters >= 1, "clip_create: max_iters must be >= 1");
    VS_REQUIRE(ctx, P.smallest_fraction >= 0.f && P.smallest_fraction <= 1.f, "clip_create: smallest_fraction must be in [0, 1]");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));

    vs_clip* c = new vs_clip();
    c->ctx = ctx; c->w = width; c->h = height; c->capacity = capacity; c->max_pairs = max_pairs;
    c->flags = flags; c->params = P;

    // level count exactly as ComputePyramid (alignment.cpp:164-169)
    VsClipGeom& g = c->g;
    memset(&g, 0, sizeof(g));
    {
        int w = width, h = height, levels = 0;
        do { levels++; w /= 2; h /= 2; } while (w >= P.pyramid_min_width && h >= P.pyramid_min_height);
        if (levels > VS_MAX_LEVELS) { delete c; return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "more than %d pyramid levels", VS_MAX_LEVELS); }
        g.levels = levels;
    }
    {
        int w = width, h = height;
        size_t off = 0; int toff = 0;
        for (int l = 0; l < g.levels; l++) {
            if (l > 0) { w /= 2; h /= 2; }
            if (w < 1 || h < 1) { delete c; return vs_set_error(ctx, VS_ERR_INVALID, "image too small for its pyramid"); }
            VsLevel& L = g.lv[l];
            L.w = w; L.h = h; L.pitch = (int)vs_align_up((size_t)w, 128);
            L.tile = vs_grad_argmax_tile_size(w, h);
            L.tw = w / L.tile; L.th = h / L.tile; L.ntiles = L.tw * L.th;
            L.img_off = (uint32_t)off; L.tile_off = (uint32_t)toff;
            off += (size_t)L.pitch * h;
            toff += L.ntiles;
            g.max_tiles = L.ntiles > g.max_tiles ? L.ntiles : g.max_tiles;
            if (L.ntiles < 1) { delete c; return vs_set_error(ctx, VS_ERR_INVALID, "pyramid level %d has no tiles", l); }
            // the solver packs a selected keypoint as tile column | tile row << 10 | offsets
            if (L.tw > 1023 || L.th > 1023) { delete c; return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "pyramid level %d has more than 1023 tiles along an axis", l); }
        }
        g.total_tiles = toff;
        g.pyr_slot_bytes = vs_align_up(off, 256);
        // NV12 clips upload the Y planes of a run of slots as the slices of one 3-D copy: whole level-0 rows per slot
        if (flags & VS_CLIP_NV12) g.pyr_slot_bytes = vs_align_up(off, (size_t)2 * g.lv[0].pitch);
    }
    // the solver keeps a level's selection keys (8 B per tile) in shared memory when they fit (up to ~24 k tiles: 4K has
    // 20 736) and in a global scratch slice per pair otherwise (8K: 82 944); a key holds the tile index in 17 bits
    const bool keys_global = (size_t)2 * g.max_tiles * 4 + (size_t)g.max_tiles + 64 > 216 * 1024;
    if (g.max_tiles > 131071 || (size_t)g.max_tiles + 64 > 200 * 1024) {
        const int mt = g.max_tiles;
        delete c;
        return vs_set_error(ctx, VS_ERR_UNSUPPORTED, "%d tiles on a level exceeds the selection's capacity", mt);
    }
    c->nv12 = (flags & VS_CLIP_NV12) != 0;
    if (c->nv12 && ((width | height) & 1)) {
        delete c;
        return vs_set_error(ctx, VS_ERR_INVALID, "clip_create: NV12 frames have even width and height");
    }
    c->bgr_pitch = vs_align_up((size_t)width * 3, 128);
    c->bgr_slot_bytes = c->nv12 ? 0 : c->bgr_pitch * height;
    c->uv_pitch = vs_align_up((size_t)width, 128);
    c->uv_slot_bytes = c->nv12 ? c->uv_pitch * (height / 2) : 0;

    int r = VS_OK;
    size_t feat = (size_t)capacity * 2 * g.total_tiles;
    int nslots = capacity > max_pairs ? capacity : max_pairs;
    if (r == VS_OK && !c->nv12) r = dev_alloc(ctx, &c->d_bgr, c->bgr_slot_bytes * capacity);
    if (r == VS_OK && c->nv12) r = dev_alloc(ctx, &c->d_uv, c->uv_slot_bytes * capacity);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pyr, g.pyr_slot_bytes * capacity);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_kp, feat);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_jac, feat);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pairs, (size_t)max_pairs);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_T, (size_t)max_pairs * 4);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_status, (size_t)max_pairs);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_iters, (size_t)max_pairs * g.levels);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_slots, (size_t)nslots);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_coef, (size_t)capacity * 2);   // NV12: a second set for the UV planes
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_pos_scratch, (size_t)max_pairs * 4 * g.max_tiles);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_res_scratch, (size_t)max_pairs * 2 * g.max_tiles);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_patch_scratch, (size_t)max_pairs * 2 * g.max_tiles);
    if (r == VS_OK) r = dev_alloc(ctx, &c->d_tb_scratch, (size_t)max_pairs * 2 * g.max_tiles);
    if (r == VS_OK && (keys_global || (size_t)2 * g.max_tiles * 4 > VS_SOLVE_BIG_KEYS))
        r = dev_alloc(ctx, &c->d_key_scratch, (size_t)max_pairs * 2 * g.max_tiles);
    if (r == VS_OK && (flags & VS_CLIP_DEBUG_TAPS)) {
        r = dev_alloc(ctx, &c->d_dbg_wd, (size_t)max_pairs * 2 * g.total_tiles);
        if (r == VS_OK) r = dev_alloc(ctx, &c->d_dbg_order, (size_t)max_pairs * 2 * g.total_tiles);
        if (r == VS_OK) r = dev_alloc(ctx, &c->d_dbg_count, (size_t)max_pairs * 2 * g.levels);
        if (r == VS_OK) r = dev_alloc(ctx, &c->d_dbg_clock, (size_t)max_pairs * VS_CLK_STRIDE);
    }
    if (r != VS_OK) { free_all(c); delete c; return r; }
    // padding bytes of the pyramid rows are never read as pixels, but keep them defined
    cudaMemsetAsync(c->d_pyr, 0, g.pyr_slot_bytes * capacity, ctx->stream);
    if (c->nv12) cudaMemsetAsync(c->d_uv, 0, c->uv_slot_bytes * capacity, ctx->stream);
    else {
        cudaMemsetAsync(c->d_bgr, 0, c->bgr_slot_bytes * capacity, ctx->stream);
        build_bgr_tensor_map(c);
    }
    *out = c;
    return VS_OK;
}

int vs_clip_destroy(vs_clip* c)
{
    if (!c) return VS_OK;
    cudaSetDevice(c->ctx->device);
    if (c->up_stream) cudaStreamSynchronize(c->up_stream);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->down_stream) cudaStreamSynchronize(c->down_stream);
    free_all(c);
    delete c;
    return VS_OK;
}

int vs_clip_set_params(vs_clip* c, const vs_align_params* params)
{
    if (!c || !params) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, params->max_iters >= 1, "clip_set_params: max_iters must be >= 1");
    VS_REQUIRE(ctx, params->smallest_fraction >= 0.f && params->smallest_fraction <= 1.f,
               "clip_set_params: smallest_fraction must be in [0, 1]");
    const int mw = c->params.pyramid_min_width, mh = c->params.pyramid_min_height;
    c->params = *params;
    c->params.pyramid_min_width = mw; c->params.pyramid_min_height = mh;
    return VS_OK;
}

int vs_clip_levels(const vs_clip* c) { return c ? c->g.levels : 0; }

int vs_clip_level_info(const vs_clip* c, int level, int* w, int* h, int* tile, int* tw, int* th)
{
    if (!c || level < 0 || level >= c->g.levels) return VS_ERR_INVALID;
    const VsLevel& L = c->g.lv[level];
    if (w) *w = L.w; if (h) *h = L.h; if (tile) *tile = L.tile; if (tw) *tw = L.tw; if (th) *th = L.th;
    return VS_OK;
}

int vs_clip_upload(vs_clip* c, int slot0, int n, const uint8_t* bgr, int64_t row_stride, int64_t frame_stride, int mem)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot0 >= 0 && n >= 0 && slot0 + n <= c->capacity, "clip_upload: slot range out of bounds");
    VS_REQUIRE(ctx, n == 0 || bgr, "clip_upload: source is NULL");
    VS_REQUIRE(ctx, row_stride >= (int64_t)c->w * (c->nv12 ? 1 : 3), "clip_upload: row_stride smaller than a row");
    VS_REQUIRE(ctx, !c->nv12 || n <= 1 || frame_stride >= row_stride * (c->h + c->h / 2), "clip_upload: frame_stride smaller than an NV12 frame");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaMemcpyKind kind = mem == VS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    return clip_enqueue_upload(c, slot0, n, bgr, row_stride, frame_stride, kind, ctx->stream);
}

int vs_clip_build_pyramids(vs_clip* c, int slot0, int n)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot0 >= 0 && n >= 0 && slot0 + n <= c->capacity, "clip_build_pyramids: slot range out of bounds");
    if (n == 0) return VS_OK;
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    const VsClipGeom& g = c->g;
    VsDevImg bgr{c->nv12 ? nullptr : c->d_bgr + (size_t)slot0 * c->bgr_slot_bytes, c->w, c->h, (int64_t)c->bgr_pitch, n, (int64_t)c->bgr_slot_bytes};
    uint8_t* pyr0 = c->d_pyr + (size_t)slot0 * g.pyr_slot_bytes;
    VsDevImg prev{pyr0 + g.lv[0].img_off, g.lv[0].w, g.lv[0].h, g.lv[0].pitch, n, (int64_t)g.pyr_slot_bytes};
    if (c->pc_ready) for (int i = 0; i < n; i++) c->pc_valid[slot0 + i] = 0;   // the spectra follow the pyramids
    if (c->nv12) {
        // level 0 is the Y plane the upload put there
        for (int l = 1; l < g.levels; l++) {
            VsDevImg cur{pyr0 + g.lv[l].img_off, g.lv[l].w, g.lv[l].h, g.lv[l].pitch, n, (int64_t)g.pyr_slot_bytes};
            VS_TRY(vsk_pyr_down(ctx, prev, cur));
            if (l == 1) note_bgr_read(c);
            prev = cur;
        }
        return VS_OK;
    }
    // gray + level 1 in one pass over the BGR frame when the geometry allows it (every 16-aligned width: 720p, 1080p, 4K ...)
    int l0 = 1;
    if (g.levels > 1) {
        VsDevImg lv1{pyr0 + g.lv[1].img_off, g.lv[1].w, g.lv[1].h, g.lv[1].pitch, n, (int64_t)g.pyr_slot_bytes};
        bool fused = false;
        VS_TRY(vsk_ingest_bgr_gray_l1(ctx, bgr, prev, lv1, &fused));
        if (fused) { prev = lv1; l0 = 2; }
    }
    if (l0 == 1) VS_TRY(vsk_bgr2gray(ctx, bgr, prev));
    note_bgr_read(c);
    for (int l = l0; l < g.levels; l++) {
        VsDevImg cur{pyr0 + g.lv[l].img_off, g.lv[l].w, g.lv[l].h, g.lv[l].pitch, n, (int64_t)g.pyr_slot_bytes};
        VS_TRY(vsk_pyr_down(ctx, prev, cur));
        prev = cur;
    }
    return VS_OK;
}

int vs_clip_build_keyframes(vs_clip* c, const int32_t* slots, int n)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, n >= 0 && n <= c->capacity && (n == 0 || slots), "clip_build_keyframes: bad slot list");
    if (n == 0) return VS_OK;
    for (int i = 0; i < n; i++) VS_REQUIRE(ctx, slots[i] >= 0 && slots[i] < c->capacity, "clip_build_keyframes: slot out of range");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_slots, slots, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    return vsk_keyframe_features(ctx, c->g, c->d_pyr, c->d_slots, n, c->d_kp, c->d_jac);
}

int vs_clip_align(vs_clip* c, const vs_pair* pairs, int n, double* out_T, int32_t* out_status, int32_t* out_iters, int mem)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, n >= 0 && n <= c->max_pairs, "clip_align: more pairs than max_pairs");
    VS_REQUIRE(ctx, mem == VS_MEM_HOST || mem == VS_MEM_DEVICE, "clip_align: bad mem");
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, pairs && out_T && out_status, "clip_align: NULL pointer");
    for (int i = 0; i < n; i++)
        VS_REQUIRE(ctx, pairs[i].template_slot >= 0 && pairs[i].template_slot < c->capacity &&
                        pairs[i].keyframe_slot >= 0 && pairs[i].keyframe_slot < c->capacity, "clip_align: slot out of range");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_pairs, pairs, (size_t)n * sizeof(vs_pair), cudaMemcpyHostToDevice, ctx->stream));
    const bool dev_out = mem == VS_MEM_DEVICE;
    VsSolveArgs a;
    a.pyr = c->d_pyr; a.kp = c->d_kp; a.jac = c->d_jac; a.pairs = c->d_pairs; a.n_pairs = n;
    a.threshold = c->params.threshold; a.fraction = c->params.smallest_fraction;
    a.max_iters = c->params.max_iters; a.max_displacement = c->params.max_displacement;
    a.out_T = dev_out ? out_T : c->d_T;
    a.out_status = dev_out ? out_status : c->d_status;
    a.out_iters = dev_out ? out_iters : (out_iters ? c->d_iters : nullptr);
    a.dbg_warpdiff = c->d_dbg_wd; a.dbg_order = c->d_dbg_order; a.dbg_count = c->d_dbg_count;
    a.pos_scratch = c->d_pos_scratch;
    a.res_scratch = c->d_res_scratch;
    a.patch_scratch = c->d_patch_scratch; a.tb_scratch = c->d_tb_scratch; a.key_scratch = c->d_key_scratch;
    a.force_threads = 0;
    a.dbg_clock = c->d_dbg_clock;
    if (c->params.phase_correlate) VS_TRY(phase_seed(c, pairs, c->d_pairs, n, 0, &a.init_T));
    if (c->d_dbg_clock)
        VS_CUDA(ctx, cudaMemsetAsync(c->d_dbg_clock, 0, (size_t)c->max_pairs * VS_CLK_STRIDE * sizeof(long long), ctx->stream));
    if (c->d_dbg_count)   // -1 marks levels a pair never reached
        VS_CUDA(ctx, cudaMemsetAsync(c->d_dbg_count, 0xFF, (size_t)c->max_pairs * 2 * c->g.levels * sizeof(int32_t), ctx->stream));
    VS_TRY(vsk_solve_pairs(ctx, c->g, a));
    c->last_pairs = n;
    if (!dev_out) {
        // results land in page-locked memory and are handed over after a stream synchronisation: a device-to-host copy
        // into pageable memory is a blocking driver call, and host threads driving other clips queue up behind it
        if (!c->h_T) {
            VS_CUDA(ctx, cudaMallocHost((void**)&c->h_T, (size_t)c->max_pairs * 4 * sizeof(double)));
            VS_CUDA(ctx, cudaMallocHost((void**)&c->h_status, (size_t)c->max_pairs * sizeof(int32_t)));
        }
        if (out_iters && !c->h_iters)
            VS_CUDA(ctx, cudaMallocHost((void**)&c->h_iters, (size_t)c->max_pairs * c->g.levels * sizeof(int32_t)));
        VS_CUDA(ctx, cudaMemcpyAsync(c->h_T, c->d_T, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaMemcpyAsync(c->h_status, c->d_status, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_iters)
            VS_CUDA(ctx, cudaMemcpyAsync(c->h_iters, c->d_iters, (size_t)n * c->g.levels * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(out_T, c->h_T, (size_t)n * 4 * sizeof(double));
        memcpy(out_status, c->h_status, (size_t)n * sizeof(int32_t));
        if (out_iters) memcpy(out_iters, c->h_iters, (size_t)n * c->g.levels * sizeof(int32_t));
    }
    return VS_OK;
}

// Batched parameter sweep (the reference sweeps VideoAlignerParams with one VideoStabilizer per
// worker thread and parameter combination, grid_search_align.cpp:134-210): every pair is solved once per parameter set in
// ONE launch.  Job (set s, pair p) uses scratch / output index s * n_pairs + p.
int vs_clip_align_sweep(vs_clip* c, const vs_pair* pairs, int n_pairs, const vs_sweep_params* sets, int n_sets,
                        double* out_T, int32_t* out_status, int mem)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, n_pairs >= 0 && n_sets >= 0 && (long long)n_pairs * n_sets <= c->max_pairs,
               "clip_align_sweep: pairs x parameter sets exceeds max_pairs");
    VS_REQUIRE(ctx, mem == VS_MEM_HOST || mem == VS_MEM_DEVICE, "clip_align_sweep: bad mem");
    const int jobs = n_pairs * n_sets;
    if (jobs == 0) return VS_OK;
    VS_REQUIRE(ctx, pairs && sets && out_T && out_status, "clip_align_sweep: NULL pointer");
    for (int i = 0; i < n_pairs; i++)
        VS_REQUIRE(ctx, pairs[i].template_slot >= 0 && pairs[i].template_slot < c->capacity &&
                        pairs[i].keyframe_slot >= 0 && pairs[i].keyframe_slot < c->capacity, "clip_align_sweep: slot out of range");
    std::vector<VsSweepSet> hs(n_sets);
    bool any_seed = false;
    for (int i = 0; i < n_sets; i++) {
        VS_REQUIRE(ctx, sets[i].max_iters >= 1, "clip_align_sweep: max_iters must be >= 1");
        VS_REQUIRE(ctx, sets[i].smallest_fraction >= 0.f && sets[i].smallest_fraction <= 1.f,
                   "clip_align_sweep: smallest_fraction must be in [0, 1]");
        hs[i].threshold = sets[i].threshold; hs[i].max_displacement = sets[i].max_displacement;
        hs[i].fraction = sets[i].smallest_fraction; hs[i].max_iters = sets[i].max_iters;
        hs[i].use_seed = sets[i].phase_correlate ? 1 : 0; hs[i].pad = 0;
        any_seed = any_seed || sets[i].phase_correlate;
    }
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)n_sets > c->sweep_capacity) {
        VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(c->d_sweep); c->d_sweep = nullptr; c->sweep_capacity = 0;
        VS_TRY(dev_alloc(ctx, &c->d_sweep, (size_t)n_sets));
        c->sweep_capacity = (size_t)n_sets;
    }
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_pairs, pairs, (size_t)n_pairs * sizeof(vs_pair), cudaMemcpyHostToDevice, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_sweep, hs.data(), (size_t)n_sets * sizeof(VsSweepSet), cudaMemcpyHostToDevice, ctx->stream));
    const bool dev_out = mem == VS_MEM_DEVICE;
    VsSolveArgs a;
    a.pyr = c->d_pyr; a.kp = c->d_kp; a.jac = c->d_jac; a.pairs = c->d_pairs; a.n_pairs = jobs;
    a.threshold = c->params.threshold; a.fraction = c->params.smallest_fraction;
    a.max_iters = c->params.max_iters; a.max_displacement = c->params.max_displacement;
    a.out_T = dev_out ? out_T : c->d_T;
    a.out_status = dev_out ? out_status : c->d_status;
    a.out_iters = nullptr;
    a.dbg_warpdiff = nullptr; a.dbg_order = nullptr; a.dbg_count = nullptr; a.dbg_clock = nullptr;
    a.pos_scratch = c->d_pos_scratch;
    a.res_scratch = c->d_res_scratch;
    a.patch_scratch = c->d_patch_scratch; a.tb_scratch = c->d_tb_scratch; a.key_scratch = c->d_key_scratch;
    a.force_threads = 0;
    a.sweep = c->d_sweep;
    a.sweep_pairs = n_pairs;
    if (any_seed) VS_TRY(phase_seed(c, pairs, c->d_pairs, n_pairs, 0, &a.init_T));   // alignment.cpp:369-388, once per pair
    VS_TRY(vsk_solve_pairs(ctx, c->g, a));
    c->last_pairs = jobs;
    if (!dev_out) {
        if (!c->h_T) {
            VS_CUDA(ctx, cudaMallocHost((void**)&c->h_T, (size_t)c->max_pairs * 4 * sizeof(double)));
            VS_CUDA(ctx, cudaMallocHost((void**)&c->h_status, (size_t)c->max_pairs * sizeof(int32_t)));
        }
        VS_CUDA(ctx, cudaMemcpyAsync(c->h_T, c->d_T, (size_t)jobs * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaMemcpyAsync(c->h_status, c->d_status, (size_t)jobs * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(out_T, c->h_T, (size_t)jobs * 4 * sizeof(double));
        memcpy(out_status, c->h_status, (size_t)jobs * sizeof(int32_t));
    }
    return VS_OK;
}

int vs_clip_align_async(vs_clip* c, const vs_pair* pairs, int n, int base, int lane)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, lane >= 0 && lane < vs_clip::kLanes, "clip_align_async: lane out of range");
    VS_REQUIRE(ctx, n >= 0 && base >= 0 && base + n <= c->max_pairs, "clip_align_async: pair range exceeds max_pairs");
    c->lane_base[lane] = base; c->lane_n[lane] = n;
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, pairs, "clip_align_async: NULL pointer");
    for (int i = 0; i < n; i++)
        VS_REQUIRE(ctx, pairs[i].template_slot >= 0 && pairs[i].template_slot < c->capacity &&
                        pairs[i].keyframe_slot >= 0 && pairs[i].keyframe_slot < c->capacity, "clip_align_async: slot out of range");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!c->solve_stream[lane]) {
        VS_CUDA(ctx, cudaStreamCreateWithFlags(&c->solve_stream[lane], cudaStreamNonBlocking));
        VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_ready[lane], cudaEventDisableTiming));
        VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_solved[lane], cudaEventDisableTiming));
    }
    if (!c->h_T) {
        VS_CUDA(ctx, cudaMallocHost((void**)&c->h_T, (size_t)c->max_pairs * 4 * sizeof(double)));
        VS_CUDA(ctx, cudaMallocHost((void**)&c->h_status, (size_t)c->max_pairs * sizeof(int32_t)));
    }
    cudaStream_t s = c->solve_stream[lane];
    // the solve follows everything enqueued on the context stream so far (pyramids, keyframe features, the pair list)
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_pairs + base, pairs, (size_t)n * sizeof(vs_pair), cudaMemcpyHostToDevice, ctx->stream));
    const double* init_T = nullptr;
    if (c->params.phase_correlate) VS_TRY(phase_seed(c, pairs, c->d_pairs + base, n, base, &init_T));
    VS_CUDA(ctx, cudaEventRecord(c->ev_ready[lane], ctx->stream));
    VS_CUDA(ctx, cudaStreamWaitEvent(s, c->ev_ready[lane], 0));
    VsSolveArgs a;
    a.pyr = c->d_pyr; a.kp = c->d_kp; a.jac = c->d_jac; a.pairs = c->d_pairs + base; a.n_pairs = n;
    a.threshold = c->params.threshold; a.fraction = c->params.smallest_fraction;
    a.max_iters = c->params.max_iters; a.max_displacement = c->params.max_displacement;
    a.out_T = c->d_T + (size_t)base * 4;
    a.out_status = c->d_status + base;
    a.out_iters = nullptr;
    a.dbg_warpdiff = nullptr; a.dbg_order = nullptr; a.dbg_count = nullptr; a.dbg_clock = nullptr;
    a.pos_scratch = c->d_pos_scratch ? c->d_pos_scratch + (size_t)base * 4 * c->g.max_tiles : nullptr;
    a.res_scratch = c->d_res_scratch ? c->d_res_scratch + (size_t)base * 2 * c->g.max_tiles : nullptr;
    a.patch_scratch = c->d_patch_scratch ? c->d_patch_scratch + (size_t)base * 2 * c->g.max_tiles : nullptr;
    a.tb_scratch = c->d_tb_scratch ? c->d_tb_scratch + (size_t)base * 2 * c->g.max_tiles : nullptr;
    a.key_scratch = c->d_key_scratch ? c->d_key_scratch + (size_t)base * 2 * c->g.max_tiles : nullptr;
    // all lanes must be resident on the GPU together: an SM per pair (512 threads, registers uncapped: the shortest
    // latency per pair) when every pair of every lane gets one, else three 256-thread CTAs per SM
    a.force_threads = c->max_pairs <= ctx->sm_count ? 512 : 256;
    a.init_T = init_T;
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = s;
    int r = vsk_solve_pairs(ctx, c->g, a);
    ctx->stream = main_stream;
    VS_TRY(r);
    VS_CUDA(ctx, cudaMemcpyAsync(c->h_T + (size_t)base * 4, c->d_T + (size_t)base * 4, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    VS_CUDA(ctx, cudaMemcpyAsync(c->h_status + base, c->d_status + base, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    VS_CUDA(ctx, cudaEventRecord(c->ev_solved[lane], s));
    c->last_pairs = base + n;
    return VS_OK;
}

int vs_clip_align_wait(vs_clip* c, int lane, double* out_T, int32_t* out_status)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, lane >= 0 && lane < vs_clip::kLanes, "clip_align_wait: lane out of range");
    const int n = c->lane_n[lane], base = c->lane_base[lane];
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, out_T && out_status && c->ev_solved[lane], "clip_align_wait: nothing in flight on this lane / NULL pointer");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaEventSynchronize(c->ev_solved[lane]));
    memcpy(out_T, c->h_T + (size_t)base * 4, (size_t)n * 4 * sizeof(double));
    memcpy(out_status, c->h_status + base, (size_t)n * sizeof(int32_t));
    c->lane_n[lane] = 0;
    return VS_OK;
}

int vs_clip_warp(vs_clip* c, const int32_t* slots, int n, const double* transforms, int mode, int border, int crop,
                 uint8_t* out, int64_t out_frame_stride, int mem)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, n >= 0 && n <= c->capacity, "clip_warp: more frames than capacity");
    VS_REQUIRE(ctx, mem == VS_MEM_HOST || mem == VS_MEM_DEVICE, "clip_warp: bad mem");
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, slots && transforms && out, "clip_warp: NULL pointer");
    VS_REQUIRE(ctx, crop >= 0 && 2 * crop < c->w && 2 * crop < c->h, "clip_warp: crop too large");
    VS_REQUIRE(ctx, !c->nv12 || (crop & 1) == 0, "clip_warp: NV12 clips crop by an even number of pixels");
    const int ow = c->w - 2 * crop, oh = c->h - 2 * crop;
    const size_t frame_bytes = clip_frame_bytes(c, ow, oh);
    const int64_t out_row = c->nv12 ? ow : (int64_t)ow * 3;
    VS_REQUIRE(ctx, out_frame_stride >= (int64_t)frame_bytes, "clip_warp: out_frame_stride smaller than a frame");
    for (int i = 0; i < n; i++) VS_REQUIRE(ctx, slots[i] >= 0 && slots[i] < c->capacity, "clip_warp: slot out of range");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));

    std::vector<VsWarpCoef> coef;
    clip_warp_coefs(c, transforms, n, coef);
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_coef, coef.data(), coef.size() * sizeof(VsWarpCoef), cudaMemcpyHostToDevice, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(c->d_slots, slots, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));

    uint8_t* d_out = out;
    int64_t d_stride = out_frame_stride;
    if (mem == VS_MEM_HOST) {
        size_t need = frame_bytes * n;
        if (need > c->warp_out_bytes) {
            VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(c->d_warp_out); c->d_warp_out = nullptr; c->warp_out_bytes = 0;
            VS_TRY(dev_alloc(ctx, &c->d_warp_out, need));
            c->warp_out_bytes = need;
        }
        d_out = c->d_warp_out;
        d_stride = (int64_t)frame_bytes;
    }
    VsDevImg dst{d_out, ow, oh, out_row, n, d_stride};
    VS_TRY(clip_warp_launch(c, c->d_slots, c->d_coef, dst, crop, mode, border, n));
    note_bgr_read(c);
    if (mem == VS_MEM_HOST) {
        VS_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)out_frame_stride, d_out, (size_t)d_stride, frame_bytes, n,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return VS_OK;
}

// ------------------------------------------------------------------ asynchronous transfers
static int ensure_async(vs_clip* c)
{
    vs_ctx* ctx = c->ctx;
    if (c->up_stream) return VS_OK;
    VS_CUDA(ctx, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    VS_CUDA(ctx, cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming));
    VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_bgr_read, cudaEventDisableTiming));
    // everything enqueued before the first asynchronous upload counts as a reader
    VS_CUDA(ctx, cudaEventRecord(c->ev_bgr_read, ctx->stream));
    c->bgr_read_recorded = true;
    for (int i = 0; i < vs_clip::kOutRing; i++) {
        VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_warp[i], cudaEventDisableTiming));
        VS_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_down[i], cudaEventDisableTiming));
    }
    VS_TRY(dev_alloc(ctx, &c->d_coef_ring, (size_t)vs_clip::kOutRing * c->capacity * 2));
    VS_TRY(dev_alloc(ctx, &c->d_slot_ring, (size_t)vs_clip::kOutRing * c->capacity));
    return VS_OK;
}

int vs_clip_upload_async(vs_clip* c, int slot0, int n, const uint8_t* bgr, int64_t row_stride, int64_t frame_stride)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot0 >= 0 && n >= 0 && slot0 + n <= c->capacity, "clip_upload_async: slot range out of bounds");
    VS_REQUIRE(ctx, n == 0 || bgr, "clip_upload_async: source is NULL");
    VS_REQUIRE(ctx, row_stride >= (int64_t)c->w * (c->nv12 ? 1 : 3), "clip_upload_async: row_stride smaller than a row");
    VS_REQUIRE(ctx, !c->nv12 || n <= 1 || frame_stride >= row_stride * (c->h + c->h / 2), "clip_upload_async: frame_stride smaller than an NV12 frame");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_TRY(ensure_async(c));
    if (n == 0) return VS_OK;
    // the slots being overwritten may still be read by the ingest / warps enqueued earlier (nothing else reads BGR frames;
    // the Y plane of an NV12 frame is also read by the keyframe features and the solves of its pairs, which have finished
    // by the time its warp, whose transform they decide, is enqueued)
    if (c->bgr_read_recorded) VS_CUDA(ctx, cudaStreamWaitEvent(c->up_stream, c->ev_bgr_read, 0));
    return clip_enqueue_upload(c, slot0, n, bgr, row_stride, frame_stride, cudaMemcpyHostToDevice, c->up_stream);
}

int vs_clip_wait_uploads(vs_clip* c)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    if (!c->up_stream) return VS_OK;
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaEventRecord(c->ev_upload, c->up_stream));
    VS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_upload, 0));
    return VS_OK;
}

int vs_clip_warp_to_host_async(vs_clip* c, const int32_t* slots, int n, const double* transforms, int mode, int border, int crop,
                               uint8_t* out, int64_t out_frame_stride)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, n >= 0 && n <= c->capacity, "clip_warp_to_host_async: more frames than capacity");
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, slots && transforms && out, "clip_warp_to_host_async: NULL pointer");
    VS_REQUIRE(ctx, crop >= 0 && 2 * crop < c->w && 2 * crop < c->h, "clip_warp_to_host_async: crop too large");
    VS_REQUIRE(ctx, !c->nv12 || (crop & 1) == 0, "clip_warp_to_host_async: NV12 clips crop by an even number of pixels");
    const int ow = c->w - 2 * crop, oh = c->h - 2 * crop;
    const size_t frame_bytes = clip_frame_bytes(c, ow, oh);
    VS_REQUIRE(ctx, out_frame_stride >= (int64_t)frame_bytes, "clip_warp_to_host_async: out_frame_stride smaller than a frame");
    for (int i = 0; i < n; i++) VS_REQUIRE(ctx, slots[i] >= 0 && slots[i] < c->capacity, "clip_warp_to_host_async: slot out of range");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_TRY(ensure_async(c));

    const int b = c->out_ring_next;
    c->out_ring_next = (b + 1) % vs_clip::kOutRing;
    // this staging buffer (and its coefficient / slot arrays) may still be draining to the host
    if (c->out_ring_used[b]) VS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_down[b], 0));
    const size_t need = frame_bytes * n;
    if (need > c->out_ring_bytes[b]) {
        if (c->out_ring_used[b]) VS_CUDA(ctx, cudaEventSynchronize(c->ev_down[b]));
        cudaFree(c->d_out_ring[b]); c->d_out_ring[b] = nullptr; c->out_ring_bytes[b] = 0;
        VS_TRY(dev_alloc(ctx, &c->d_out_ring[b], need));
        c->out_ring_bytes[b] = need;
    }
    std::vector<VsWarpCoef> coef;
    clip_warp_coefs(c, transforms, n, coef);
    VsWarpCoef* d_coef = c->d_coef_ring + (size_t)b * c->capacity * 2;
    int32_t* d_slots = c->d_slot_ring + (size_t)b * c->capacity;
    // pageable sources: both copies complete (staged) before the call returns
    VS_CUDA(ctx, cudaMemcpyAsync(d_coef, coef.data(), coef.size() * sizeof(VsWarpCoef), cudaMemcpyHostToDevice, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(d_slots, slots, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    VsDevImg dst{c->d_out_ring[b], ow, oh, c->nv12 ? (int64_t)ow : (int64_t)ow * 3, n, (int64_t)frame_bytes};
    VS_TRY(clip_warp_launch(c, d_slots, d_coef, dst, crop, mode, border, n));
    note_bgr_read(c);
    VS_CUDA(ctx, cudaEventRecord(c->ev_warp[b], ctx->stream));
    VS_CUDA(ctx, cudaStreamWaitEvent(c->down_stream, c->ev_warp[b], 0));
    if (out_frame_stride == (int64_t)frame_bytes)
        VS_CUDA(ctx, cudaMemcpyAsync(out, c->d_out_ring[b], need, cudaMemcpyDeviceToHost, c->down_stream));
    else
        VS_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)out_frame_stride, c->d_out_ring[b], frame_bytes, frame_bytes, n,
                                       cudaMemcpyDeviceToHost, c->down_stream));
    VS_CUDA(ctx, cudaEventRecord(c->ev_down[b], c->down_stream));
    c->out_ring_used[b] = true;
    return VS_OK;
}

int vs_clip_sync_transfers(vs_clip* c)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    if (!c->up_stream) return VS_OK;
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaStreamSynchronize(c->up_stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(c->down_stream));
    return VS_OK;
}

// ------------------------------------------------------------------ inspection taps

int vs_clip_get_bgr(vs_clip* c, int slot, uint8_t* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot >= 0 && slot < c->capacity && out, "clip_get_bgr: bad arguments");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (c->nv12) {
        const VsLevel& L0 = c->g.lv[0];
        VS_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)c->w, c->d_pyr + (size_t)slot * c->g.pyr_slot_bytes + L0.img_off, (size_t)L0.pitch,
                                       (size_t)c->w, c->h, cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaMemcpy2DAsync(out + (size_t)c->w * c->h, (size_t)c->w, c->d_uv + (size_t)slot * c->uv_slot_bytes, c->uv_pitch,
                                       (size_t)c->w, c->h / 2, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        VS_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)c->w * 3, c->d_bgr + (size_t)slot * c->bgr_slot_bytes, c->bgr_pitch,
                                       (size_t)c->w * 3, c->h, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

int vs_clip_get_gray(vs_clip* c, int slot, int level, uint8_t* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot >= 0 && slot < c->capacity && level >= 0 && level < c->g.levels && out, "clip_get_gray: bad arguments");
    const VsLevel& L = c->g.lv[level];
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpy2DAsync(out, L.w, c->d_pyr + (size_t)slot * c->g.pyr_slot_bytes + L.img_off, L.pitch, L.w, L.h,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

int vs_clip_get_keypoints(vs_clip* c, int slot, int level, int axis, uint16_t* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot >= 0 && slot < c->capacity && level >= 0 && level < c->g.levels && (axis == 0 || axis == 1) && out,
               "clip_get_keypoints: bad arguments");
    const VsLevel& L = c->g.lv[level];
    std::vector<uint32_t> tmp(L.ntiles);
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(tmp.data(), c->d_kp + ((size_t)slot * 2 + axis) * c->g.total_tiles + L.tile_off,
                                 (size_t)L.ntiles * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < L.ntiles; t++) { out[t] = (uint16_t)(tmp[t] & 0xffff); out[L.ntiles + t] = (uint16_t)(tmp[t] >> 16); }
    return VS_OK;
}

int vs_clip_get_jacobians(vs_clip* c, int slot, int level, int axis, float* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, slot >= 0 && slot < c->capacity && level >= 0 && level < c->g.levels && (axis == 0 || axis == 1) && out,
               "clip_get_jacobians: bad arguments");
    const VsLevel& L = c->g.lv[level];
    std::vector<float4> tmp(L.ntiles);
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(tmp.data(), c->d_jac + ((size_t)slot * 2 + axis) * c->g.total_tiles + L.tile_off,
                                 (size_t)L.ntiles * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < L.ntiles; t++) {
        out[t] = tmp[t].x; out[L.ntiles + t] = tmp[t].y; out[2 * L.ntiles + t] = tmp[t].z; out[3 * L.ntiles + t] = tmp[t].w;
    }
    return VS_OK;
}

int vs_clip_get_warpdiff(vs_clip* c, int pair, int level, int axis, uint16_t* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, c->d_dbg_wd, "clip_get_warpdiff: clip was created without VS_CLIP_DEBUG_TAPS");
    VS_REQUIRE(ctx, pair >= 0 && pair < c->last_pairs && level >= 0 && level < c->g.levels && (axis == 0 || axis == 1) && out,
               "clip_get_warpdiff: bad arguments");
    const VsLevel& L = c->g.lv[level];
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(out, c->d_dbg_wd + ((size_t)pair * 2 + axis) * c->g.total_tiles + L.tile_off,
                                 (size_t)L.ntiles * 2, cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

int vs_clip_get_phase(vs_clip* c, int pair, double* out3)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, c->pc_ready && c->params.phase_correlate, "clip_get_phase: the last align call did not use phase_correlate");
    VS_REQUIRE(ctx, pair >= 0 && pair < c->last_pairs && out3, "clip_get_phase: bad arguments");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(out3, c->d_pc_phase + (size_t)pair * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

int vs_clip_get_solver_cycles(vs_clip* c, int pair, long long* out6)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, c->d_dbg_clock, "clip_get_solver_cycles: clip was created without VS_CLIP_DEBUG_TAPS");
    VS_REQUIRE(ctx, pair >= 0 && pair < c->last_pairs && out6, "clip_get_solver_cycles: bad arguments");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    long long t[8];
    VS_CUDA(ctx, cudaMemcpyAsync(t, c->d_dbg_clock + (size_t)pair * VS_CLK_STRIDE, 8 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    t[3] += t[6];   // the documented six: SVD beside the first iteration + the gathers of the later iterations
    for (int i = 0; i < 6; i++) out6[i] = t[i];
    return VS_OK;
}

int vs_clip_get_solver_level_cycles(vs_clip* c, int pair, long long* out)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, c->d_dbg_clock, "clip_get_solver_level_cycles: clip was created without VS_CLIP_DEBUG_TAPS");
    VS_REQUIRE(ctx, pair >= 0 && pair < c->last_pairs && out, "clip_get_solver_level_cycles: bad arguments");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(out, c->d_dbg_clock + (size_t)pair * VS_CLK_STRIDE + 8, (size_t)c->g.levels * 8 * sizeof(long long),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

int vs_clip_get_selected(vs_clip* c, int pair, int level, int axis, uint32_t* out_order, int* out_k)
{
    if (!c) return VS_ERR_INVALID;
    vs_ctx* ctx = c->ctx;
    VS_REQUIRE(ctx, c->d_dbg_order, "clip_get_selected: clip was created without VS_CLIP_DEBUG_TAPS");
    VS_REQUIRE(ctx, pair >= 0 && pair < c->last_pairs && level >= 0 && level < c->g.levels && (axis == 0 || axis == 1) &&
                    out_order && out_k, "clip_get_selected: bad arguments");
    const VsLevel& L = c->g.lv[level];
    int k = 0;
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    VS_CUDA(ctx, cudaMemcpyAsync(&k, c->d_dbg_count + ((size_t)pair * 2 + axis) * c->g.levels + level, sizeof(int),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (k < 0 || k > L.ntiles) return vs_set_error(ctx, VS_ERR_INVALID, "clip_get_selected: level %d was not reached by this pair", level);
    if (k) {
        VS_CUDA(ctx, cudaMemcpyAsync(out_order, c->d_dbg_order + ((size_t)pair * 2 + axis) * c->g.total_tiles + L.tile_off,
                                     (size_t)k * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out_k = k;
    return VS_OK;
}

}  // extern "C"
