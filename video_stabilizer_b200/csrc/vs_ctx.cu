// vs_ctx.cu — context, error text, device memory helpers of the C ABI (include/vstab.h).
#include "vs_internal.h"

#include <stdarg.h>
#include <string.h>

static thread_local std::string g_create_error;

int vs_set_error(vs_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf; else g_create_error = buf;
    return code;
}

void vs_scratch_reset(vs_ctx* ctx) { ctx->scratch_used = 0; }

// Bump allocation out of one grow-only block.  Growing frees the old block, so it is only
// legal before anything was handed out in the current operation: callers reserve the total
// first (vs_scratch_reserve) and then carve.
static int vs_scratch_reserve(vs_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return VS_OK;
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr; ctx->scratch_bytes = 0;
    size_t want = vs_align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&ctx->scratch, want);
    if (e != cudaSuccess) return vs_set_error(ctx, VS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    ctx->scratch_bytes = want;
    return VS_OK;
}

void* vs_scratch_alloc(vs_ctx* ctx, size_t bytes)
{
    size_t off = vs_align_up(ctx->scratch_used, 256);
    if (off + bytes > ctx->scratch_bytes) {
        if (ctx->scratch_used != 0) {
            vs_set_error(ctx, VS_ERR_NOMEM, "scratch exhausted (reserve first)");
            return nullptr;
        }
        if (vs_scratch_reserve(ctx, bytes) != VS_OK) return nullptr;
        off = 0;
    }
    ctx->scratch_used = off + bytes;
    return (char*)ctx->scratch + off;
}

// ---------------------------------------------------------------- per-kernel timing
static cudaEvent_t prof_event(VsProfiler* p)
{
    if (!p->pool.empty()) { cudaEvent_t e = p->pool.back(); p->pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void vs_prof_begin(vs_ctx* ctx, int id)
{
    VsProfiler* p = ctx->prof;
    p->open_id = id;
    p->open_a = prof_event(p);
    cudaEventRecord(p->open_a, ctx->stream);
}

void vs_prof_end(vs_ctx* ctx)
{
    VsProfiler* p = ctx->prof;
    if (p->open_id < 0) return;
    cudaEvent_t b = prof_event(p);
    cudaEventRecord(b, ctx->stream);
    p->pending.push_back({p->open_a, b, p->open_id});
    p->open_id = -1;
}

static void prof_resolve(vs_ctx* ctx)
{
    VsProfiler* p = ctx->prof;
    if (!p) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : p->pending) {
        float ms = 0;
        cudaEventSynchronize(s.b);      // launches on another stream of the owner (the clip's solver lanes)
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { p->ms[s.id] += ms; p->n[s.id]++; }
        p->pool.push_back(s.a); p->pool.push_back(s.b);
    }
    p->pending.clear();
}

static void prof_destroy(vs_ctx* ctx)
{
    VsProfiler* p = ctx->prof;
    if (!p) return;
    prof_resolve(ctx);
    for (cudaEvent_t e : p->pool) cudaEventDestroy(e);
    delete p;
    ctx->prof = nullptr;
}

extern "C" {

int vs_ctx_profile_enable(vs_ctx* ctx, int enable)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (enable && !ctx->prof) ctx->prof = new VsProfiler();
    if (!enable) prof_destroy(ctx);
    return VS_OK;
}

int vs_ctx_profile_reset(vs_ctx* ctx)
{
    if (!ctx) return VS_ERR_INVALID;
    if (!ctx->prof) return VS_OK;
    prof_resolve(ctx);
    for (int i = 0; i < VSK_COUNT; i++) { ctx->prof->ms[i] = 0; ctx->prof->n[i] = 0; }
    return VS_OK;
}

int vs_ctx_profile_read(vs_ctx* ctx, int kernel, int64_t* launches, double* total_ms)
{
    if (!ctx || kernel < 0 || kernel >= VSK_COUNT) return VS_ERR_INVALID;
    if (!ctx->prof) { if (launches) *launches = 0; if (total_ms) *total_ms = 0; return VS_OK; }
    prof_resolve(ctx);
    if (launches) *launches = ctx->prof->n[kernel];
    if (total_ms) *total_ms = ctx->prof->ms[kernel];
    return VS_OK;
}

int vs_ctx_profile_timeline(vs_ctx* ctx, double* out3, int capacity)
{
    if (!ctx || !ctx->prof) return 0;
    VsProfiler* p = ctx->prof;
    cudaStreamSynchronize(ctx->stream);
    int n = 0;
    cudaEvent_t base = p->pending.empty() ? nullptr : p->pending.front().a;
    for (auto& s : p->pending) {
        cudaEventSynchronize(s.b);
        float t0 = 0, t1 = 0;
        if (n < capacity && cudaEventElapsedTime(&t0, base, s.a) == cudaSuccess && cudaEventElapsedTime(&t1, base, s.b) == cudaSuccess) {
            out3[3 * n] = s.id; out3[3 * n + 1] = t0; out3[3 * n + 2] = t1;
            n++;
        }
    }
    cudaGetLastError();
    return n;
}

const char* vs_kernel_name(int kernel)
{
    static const char* names[VSK_COUNT] = {"bgr2gray", "pyr_down", "grad_xy", "image_warp", "bgr_warp", "grad_argmax",
                                           "sparse_jac", "sparse_warpdiff", "sparse_ica", "keyframe_features", "solve_pairs",
                                           "ingest_bgr_gray_l1"};
    return kernel >= 0 && kernel < VSK_COUNT ? names[kernel] : "";
}

int vs_abi_version(void) { return VS_ABI_VERSION; }

int vs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vs_ctx_create(int device, vs_ctx** out)
{
    if (!out) return vs_set_error(nullptr, VS_ERR_INVALID, "vs_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return vs_set_error(nullptr, VS_ERR_CUDA, "no CUDA device available (%s); libvstab has no CPU fallback",
                            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return vs_set_error(nullptr, VS_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return vs_set_error(nullptr, VS_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    vs_ctx* ctx = new vs_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        int r = vs_set_error(nullptr, VS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return r;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return VS_OK;
}

int vs_ctx_destroy(vs_ctx* ctx)
{
    if (!ctx) return VS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    prof_destroy(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return VS_OK;
}

int vs_ctx_set_stream(vs_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return VS_OK;
}

int vs_ctx_synchronize(vs_ctx* ctx)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VS_OK;
}

const char* vs_last_error(const vs_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int64_t vs_ctx_launch_count(const vs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vs_dev_alloc(vs_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return VS_ERR_INVALID;
    *out = nullptr;
    if (bytes == 0) return VS_OK;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) return vs_set_error(ctx, VS_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return VS_OK;
}

int vs_dev_free(vs_ctx* ctx, void* p)
{
    if (!ctx) return VS_ERR_INVALID;
    if (p) VS_CUDA(ctx, cudaFree(p));
    return VS_OK;
}

int vs_host_alloc_pinned(vs_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return VS_ERR_INVALID;
    *out = nullptr;
    if (bytes == 0) return VS_OK;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) return vs_set_error(ctx, VS_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return VS_OK;
}

int vs_host_free_pinned(vs_ctx* ctx, void* p)
{
    if (!ctx) return VS_ERR_INVALID;
    if (p) VS_CUDA(ctx, cudaFreeHost(p));
    return VS_OK;
}

int vs_pinned_alloc(size_t bytes, void** out)
{
    if (!out) return VS_ERR_INVALID;
    *out = nullptr;
    if (bytes == 0) return VS_OK;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return vs_set_error(nullptr, VS_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return VS_OK;
}

int vs_pinned_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return VS_ERR_CUDA; }
    return VS_OK;
}

int vs_host_is_pinned(const void* p)
{
    if (!p) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

int vs_memcpy_h2d(vs_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return VS_ERR_INVALID;
    if (bytes) VS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return VS_OK;
}

int vs_memcpy_d2h(vs_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return VS_ERR_INVALID;
    if (bytes) VS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return VS_OK;
}

void vs_align_params_default(vs_align_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->phase_correlate = 0;
    p->phase_correlate_threshold = 0.5;
    p->threshold = 0.02;
    p->smallest_fraction = 0.8f;
    p->max_iters = 64;
    p->pyramid_min_width = 20;
    p->pyramid_min_height = 20;
    p->max_displacement = 10.0;
}

}  // extern "C"

// shared with vs_ops.cu
int vs_scratch_reserve_public(vs_ctx* ctx, size_t bytes) { return vs_scratch_reserve(ctx, bytes); }
