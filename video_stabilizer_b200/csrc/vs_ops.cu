// vs_ops.cu — the single-operator half of the C ABI: one entry point per Halide pipeline of
// the reference (imgproc.cpp wrappers), for host or device pointers.  With VS_MEM_HOST the
// call stages the images in device scratch (256-byte pitched), runs the kernel and copies
// the result back before returning; with VS_MEM_DEVICE it only enqueues the kernel.
#include "vs_internal.h"

#include <vector>

int vs_scratch_reserve_public(vs_ctx* ctx, size_t bytes);

namespace {

struct Stage {
    vs_ctx* ctx;
    size_t need = 0;
    explicit Stage(vs_ctx* c) : ctx(c) { vs_scratch_reset(c); }
    static size_t pitch_of(int w, size_t px_bytes) { return vs_align_up((size_t)w * px_bytes, 256); }
    void want_img(const vs_img* im, size_t px_bytes)
    {
        need += vs_align_up(pitch_of(im->width, px_bytes) * (size_t)im->height * (size_t)(im->batch < 1 ? 1 : im->batch), 256) + 256;
    }
    void want_raw(size_t bytes) { need += vs_align_up(bytes, 256) + 256; }
    int reserve() { return vs_scratch_reserve_public(ctx, need); }

    // device image with the host image's geometry; elem = bytes per element of `stride`
    int img(const vs_img* im, size_t px_bytes, size_t elem, bool copy_in, VsDevImg* out)
    {
        int batch = im->batch < 1 ? 1 : im->batch;
        size_t pitch = pitch_of(im->width, px_bytes);
        size_t per = pitch * (size_t)im->height;
        void* d = vs_scratch_alloc(ctx, per * batch);
        if (!d && per != 0) return VS_ERR_NOMEM;
        out->data = d; out->w = im->width; out->h = im->height;
        out->stride = (int64_t)(pitch / elem); out->batch = batch; out->batch_stride = (int64_t)(per / elem);
        if (copy_in && im->width > 0 && im->height > 0) {
            for (int b = 0; b < batch; b++)
                VS_CUDA(ctx, cudaMemcpy2DAsync((char*)d + per * b, pitch,
                                               (const char*)im->data + (size_t)im->batch_stride * elem * b,
                                               (size_t)im->stride * elem, (size_t)im->width * px_bytes, im->height,
                                               cudaMemcpyHostToDevice, ctx->stream));
        }
        return VS_OK;
    }
    int img_out(const VsDevImg& dv, const vs_img* im, size_t px_bytes, size_t elem)
    {
        if (im->width <= 0 || im->height <= 0) return VS_OK;
        for (int b = 0; b < dv.batch; b++)
            VS_CUDA(ctx, cudaMemcpy2DAsync((char*)im->data + (size_t)im->batch_stride * elem * b, (size_t)im->stride * elem,
                                           (const char*)dv.data + (size_t)dv.batch_stride * elem * b, (size_t)dv.stride * elem,
                                           (size_t)im->width * px_bytes, im->height, cudaMemcpyDeviceToHost, ctx->stream));
        return VS_OK;
    }
    int raw(const void* host, size_t bytes, bool copy_in, void** out)
    {
        void* d = vs_scratch_alloc(ctx, bytes ? bytes : 1);
        if (!d) return VS_ERR_NOMEM;
        *out = d;
        if (copy_in && bytes) VS_CUDA(ctx, cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return VS_OK;
    }
    int raw_out(void* host, const void* dev, size_t bytes)
    {
        if (bytes) VS_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return VS_OK;
    }
    int finish() { VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return VS_OK; }
};

#define VS_TRY(expr) do { int _r = (expr); if (_r != VS_OK) return _r; } while (0)

int check_img(vs_ctx* ctx, const vs_img* im, const char* what)
{
    if (!im) return vs_set_error(ctx, VS_ERR_INVALID, "%s: image descriptor is NULL", what);
    if (im->width < 0 || im->height < 0) return vs_set_error(ctx, VS_ERR_INVALID, "%s: negative extent", what);
    if (im->width > 0 && im->height > 0 && !im->data) return vs_set_error(ctx, VS_ERR_INVALID, "%s: data is NULL", what);
    return VS_OK;
}

int check_mem(vs_ctx* ctx, int mem)
{
    if (mem != VS_MEM_HOST && mem != VS_MEM_DEVICE) return vs_set_error(ctx, VS_ERR_INVALID, "mem must be VS_MEM_HOST or VS_MEM_DEVICE");
    return VS_OK;
}

}  // namespace

extern "C" {

int vs_bgr2gray_u8(vs_ctx* ctx, const vs_img* bgr, const vs_img* gray, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, bgr, "bgr2gray input")); VS_TRY(check_img(ctx, gray, "bgr2gray output"));
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == VS_MEM_DEVICE) return vsk_bgr2gray(ctx, vs_dev_img(bgr), vs_dev_img(gray));
    Stage st(ctx);
    st.want_img(bgr, 3); st.want_img(gray, 1);
    VS_TRY(st.reserve());
    VsDevImg din, dout;
    VS_TRY(st.img(bgr, 3, 1, true, &din));
    VS_TRY(st.img(gray, 1, 1, false, &dout));
    VS_TRY(vsk_bgr2gray(ctx, din, dout));
    VS_TRY(st.img_out(dout, gray, 1, 1));
    return st.finish();
}

int vs_ingest_bgr_u8(vs_ctx* ctx, const vs_img* bgr, const vs_img* gray0, const vs_img* gray1, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, bgr, "ingest input"));
    VS_TRY(check_img(ctx, gray0, "ingest gray level 0")); VS_TRY(check_img(ctx, gray1, "ingest gray level 1"));
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    auto run = [&](const VsDevImg& din, const VsDevImg& d0, const VsDevImg& d1) -> int {
        bool fused = false;
        VS_TRY(vsk_ingest_bgr_gray_l1(ctx, din, d0, d1, &fused));
        if (fused) return VS_OK;
        VS_TRY(vsk_bgr2gray(ctx, din, d0));
        return vsk_pyr_down(ctx, d0, d1);
    };
    if (mem == VS_MEM_DEVICE) return run(vs_dev_img(bgr), vs_dev_img(gray0), vs_dev_img(gray1));
    Stage st(ctx);
    st.want_img(bgr, 3); st.want_img(gray0, 1); st.want_img(gray1, 1);
    VS_TRY(st.reserve());
    VsDevImg din, d0, d1;
    VS_TRY(st.img(bgr, 3, 1, true, &din));
    VS_TRY(st.img(gray0, 1, 1, false, &d0));
    VS_TRY(st.img(gray1, 1, 1, false, &d1));
    VS_TRY(run(din, d0, d1));
    VS_TRY(st.img_out(d0, gray0, 1, 1));
    VS_TRY(st.img_out(d1, gray1, 1, 1));
    return st.finish();
}

int vs_phase_correlate_u8(vs_ctx* ctx, const vs_img* src1, const vs_img* src2, double* out3, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, src1, "phase_correlate src1")); VS_TRY(check_img(ctx, src2, "phase_correlate src2"));
    VS_REQUIRE(ctx, out3, "phase_correlate: out3 is NULL");
    VS_REQUIRE(ctx, src1->width == src2->width && src1->height == src2->height && src1->width > 0 && src1->height > 0,
               "phase_correlate: the images must have one non-empty size");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int w = src1->width, h = src1->height;
    VsPhasePlan p;
    p.w = w; p.h = h; p.pitch = w;
    p.M = vs_optimal_dft_size(h); p.N = vs_optimal_dft_size(w); p.Kh = p.N / 2 + 1;
    const size_t spec = (size_t)p.M * p.Kh * 2 * sizeof(double);
    // one allocation: twiddles | row transforms x2 | spectra x2 | cross | inverse | surface | result | images | slots, pair
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += vs_align_up(bytes, 256); return o; };
    const size_t o_tw = take((size_t)(p.N + p.M) * 16), o_rows = take((size_t)2 * h * p.Kh * 16), o_spec = take(2 * spec),
                 o_cross = take(spec), o_inv = take(spec), o_surf = take((size_t)p.M * p.N * 8), o_res = take(3 * 8),
                 o_img = take((size_t)2 * w * h), o_idx = take(64);
    uint8_t* base = nullptr;
    VS_CUDA(ctx, cudaMalloc((void**)&base, off));
    p.d_tw = (double*)(base + o_tw); p.d_rows = (double*)(base + o_rows); p.d_spec = (double*)(base + o_spec);
    p.d_cross = (double*)(base + o_cross); p.d_inv = (double*)(base + o_inv); p.d_surf = (double*)(base + o_surf);
    std::vector<double> tw((size_t)(p.N + p.M) * 2);
    vs_phase_twiddles(p.N, tw.data());
    vs_phase_twiddles(p.M, tw.data() + (size_t)p.N * 2);
    struct { int32_t slots[2]; vs_pair pair; } idx = {{0, 1}, {0, 1, 0}};   // previous = template slot 0, current = keyframe slot 1
    const cudaMemcpyKind kind = mem == VS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaError_t e = cudaMemcpyAsync(p.d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_idx, &idx, sizeof(idx), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(base + o_img, w, src1->data, (size_t)src1->stride, w, h, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(base + o_img + (size_t)w * h, w, src2->data, (size_t)src2->stride, w, h, kind, ctx->stream);
    int r = e == cudaSuccess ? VS_OK : vs_set_error(ctx, VS_ERR_CUDA, "phase_correlate: copy failed: %s", cudaGetErrorString(e));
    if (r == VS_OK) r = vsk_phase_forward(ctx, p, base + o_img, (size_t)w * h, (const int32_t*)(base + o_idx), 2);
    if (r == VS_OK) r = vsk_phase_pairs(ctx, p, (const vs_pair*)(base + o_idx + 8), 1, 0.0, 1.0f, (double*)(base + o_res), nullptr);
    if (r == VS_OK) {
        e = cudaMemcpyAsync(out3, base + o_res, 3 * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) r = vs_set_error(ctx, VS_ERR_CUDA, "phase_correlate: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(base);
    return r;
}

int vs_pyr_down_u8(vs_ctx* ctx, const vs_img* in, const vs_img* out, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, in, "pyr_down input")); VS_TRY(check_img(ctx, out, "pyr_down output"));
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == VS_MEM_DEVICE) return vsk_pyr_down(ctx, vs_dev_img(in), vs_dev_img(out));
    Stage st(ctx);
    st.want_img(in, 1); st.want_img(out, 1);
    VS_TRY(st.reserve());
    VsDevImg din, dout;
    VS_TRY(st.img(in, 1, 1, true, &din));
    VS_TRY(st.img(out, 1, 1, false, &dout));
    VS_TRY(vsk_pyr_down(ctx, din, dout));
    VS_TRY(st.img_out(dout, out, 1, 1));
    return st.finish();
}

int vs_grad_xy_u8_f32(vs_ctx* ctx, const vs_img* in, const vs_img* gx, const vs_img* gy, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, in, "grad_xy input"));
    VS_TRY(check_img(ctx, gx, "grad_xy grad_x")); VS_TRY(check_img(ctx, gy, "grad_xy grad_y"));
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == VS_MEM_DEVICE) return vsk_grad_xy(ctx, vs_dev_img(in), vs_dev_img(gx), vs_dev_img(gy));
    Stage st(ctx);
    st.want_img(in, 1); st.want_img(gx, 4); st.want_img(gy, 4);
    VS_TRY(st.reserve());
    VsDevImg din, dgx, dgy;
    VS_TRY(st.img(in, 1, 1, true, &din));
    VS_TRY(st.img(gx, 4, 4, false, &dgx));
    VS_TRY(st.img(gy, 4, 4, false, &dgy));
    VS_TRY(vsk_grad_xy(ctx, din, dgx, dgy));
    VS_TRY(st.img_out(dgx, gx, 4, 4));
    VS_TRY(st.img_out(dgy, gy, 4, 4));
    return st.finish();
}

// imgproc.cpp:151-162
int vs_grad_argmax_tile_size(int width, int height)
{
    int tile = 2;
    for (int i = 4; i <= 20; i += 2) {
        int tx = width / i, ty = height / i;
        if (tx * ty < 1000) break;
        tile = i;
    }
    return tile;
}

int vs_grad_argmax_f32_u16(vs_ctx* ctx, const vs_img* gx, const vs_img* gy, int tile,
                           uint16_t* lmx, uint16_t* lmy, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, gx, "grad_argmax grad_x")); VS_TRY(check_img(ctx, gy, "grad_argmax grad_y"));
    VS_REQUIRE(ctx, tile >= 1, "grad_argmax: tile must be >= 1");
    VS_REQUIRE(ctx, gx->batch <= 1 && gy->batch <= 1, "grad_argmax: batch not supported");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    int tw = gx->width / tile, th = gy->height / tile;
    size_t n = (size_t)tw * th * 2;
    VS_REQUIRE(ctx, n == 0 || (lmx && lmy), "grad_argmax: output is NULL");
    if (mem == VS_MEM_DEVICE) return vsk_grad_argmax(ctx, vs_dev_img(gx), vs_dev_img(gy), tile, lmx, lmy);
    Stage st(ctx);
    st.want_img(gx, 4); st.want_img(gy, 4); st.want_raw(n * 2); st.want_raw(n * 2);
    VS_TRY(st.reserve());
    VsDevImg dgx, dgy; void *dx, *dy;
    VS_TRY(st.img(gx, 4, 4, true, &dgx));
    VS_TRY(st.img(gy, 4, 4, true, &dgy));
    VS_TRY(st.raw(nullptr, n * 2, false, &dx));
    VS_TRY(st.raw(nullptr, n * 2, false, &dy));
    VS_TRY(vsk_grad_argmax(ctx, dgx, dgy, tile, (uint16_t*)dx, (uint16_t*)dy));
    VS_TRY(st.raw_out(lmx, dx, n * 2));
    VS_TRY(st.raw_out(lmy, dy, n * 2));
    return st.finish();
}

int vs_sparse_jac_f32(vs_ctx* ctx, const vs_img* gx, const vs_img* gy, const uint16_t* lmx, const uint16_t* lmy,
                      int tw, int th, float* out_x, float* out_y, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, gx, "sparse_jac grad_x")); VS_TRY(check_img(ctx, gy, "sparse_jac grad_y"));
    VS_REQUIRE(ctx, tw >= 0 && th >= 0, "sparse_jac: negative tile extent");
    VS_REQUIRE(ctx, gx->width > 0 && gx->height > 0, "sparse_jac: empty gradient image");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)tw * th;
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, lmx && lmy && out_x && out_y, "sparse_jac: NULL pointer");
    if (mem == VS_MEM_DEVICE) return vsk_sparse_jac(ctx, vs_dev_img(gx), vs_dev_img(gy), lmx, lmy, tw, th, out_x, out_y);
    Stage st(ctx);
    st.want_img(gx, 4); st.want_img(gy, 4);
    st.want_raw(n * 4); st.want_raw(n * 4); st.want_raw(n * 16); st.want_raw(n * 16);
    VS_TRY(st.reserve());
    VsDevImg dgx, dgy; void *dlx, *dly, *dox, *doy;
    VS_TRY(st.img(gx, 4, 4, true, &dgx));
    VS_TRY(st.img(gy, 4, 4, true, &dgy));
    VS_TRY(st.raw(lmx, n * 4, true, &dlx));
    VS_TRY(st.raw(lmy, n * 4, true, &dly));
    VS_TRY(st.raw(nullptr, n * 16, false, &dox));
    VS_TRY(st.raw(nullptr, n * 16, false, &doy));
    VS_TRY(vsk_sparse_jac(ctx, dgx, dgy, (const uint16_t*)dlx, (const uint16_t*)dly, tw, th, (float*)dox, (float*)doy));
    VS_TRY(st.raw_out(out_x, dox, n * 16));
    VS_TRY(st.raw_out(out_y, doy, n * 16));
    return st.finish();
}

int vs_sparse_warpdiff_u8_u16(vs_ctx* ctx, const vs_img* tmpl, const vs_img* key, const uint16_t* lm, int tw, int th,
                              float A, float B, float TX, float TY, uint16_t* out, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, tmpl, "warpdiff template")); VS_TRY(check_img(ctx, key, "warpdiff keyframe"));
    VS_REQUIRE(ctx, tw >= 0 && th >= 0, "warpdiff: negative tile extent");
    VS_REQUIRE(ctx, key->width > 0 && key->height > 0, "warpdiff: empty image");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)tw * th;
    if (n == 0) return VS_OK;
    VS_REQUIRE(ctx, lm && out, "warpdiff: NULL pointer");
    if (mem == VS_MEM_DEVICE) return vsk_sparse_warpdiff(ctx, vs_dev_img(tmpl), vs_dev_img(key), lm, tw, th, A, B, TX, TY, out);
    Stage st(ctx);
    st.want_img(tmpl, 1); st.want_img(key, 1); st.want_raw(n * 4); st.want_raw(n * 2);
    VS_TRY(st.reserve());
    VsDevImg dt, dk; void *dlm, *dout;
    VS_TRY(st.img(tmpl, 1, 1, true, &dt));
    VS_TRY(st.img(key, 1, 1, true, &dk));
    VS_TRY(st.raw(lm, n * 4, true, &dlm));
    VS_TRY(st.raw(nullptr, n * 2, false, &dout));
    VS_TRY(vsk_sparse_warpdiff(ctx, dt, dk, (const uint16_t*)dlm, tw, th, A, B, TX, TY, (uint16_t*)dout));
    VS_TRY(st.raw_out(out, dout, n * 2));
    return st.finish();
}

int vs_sparse_ica_f64(vs_ctx* ctx, const vs_img* tmpl, const vs_img* key, const uint16_t* selx, int kx,
                      const uint16_t* sely, int ky, const float* jx, const float* jy,
                      float A, float B, float TX, float TY, double* out4, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, tmpl, "ica template")); VS_TRY(check_img(ctx, key, "ica keyframe"));
    VS_REQUIRE(ctx, kx >= 0 && ky >= 0 && out4, "ica: bad arguments");
    VS_REQUIRE(ctx, key->width > 0 && key->height > 0, "ica: empty image");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == VS_MEM_DEVICE) return vsk_sparse_ica(ctx, vs_dev_img(tmpl), vs_dev_img(key), selx, kx, sely, ky, jx, jy, A, B, TX, TY, out4);
    Stage st(ctx);
    st.want_img(tmpl, 1); st.want_img(key, 1);
    st.want_raw((size_t)kx * 4); st.want_raw((size_t)ky * 4); st.want_raw((size_t)kx * 16); st.want_raw((size_t)ky * 16); st.want_raw(32);
    VS_TRY(st.reserve());
    VsDevImg dt, dk; void *dsx, *dsy, *djx, *djy, *dout;
    VS_TRY(st.img(tmpl, 1, 1, true, &dt));
    VS_TRY(st.img(key, 1, 1, true, &dk));
    VS_TRY(st.raw(selx, (size_t)kx * 4, true, &dsx));
    VS_TRY(st.raw(sely, (size_t)ky * 4, true, &dsy));
    VS_TRY(st.raw(jx, (size_t)kx * 16, true, &djx));
    VS_TRY(st.raw(jy, (size_t)ky * 16, true, &djy));
    VS_TRY(st.raw(nullptr, 32, false, &dout));
    VS_TRY(vsk_sparse_ica(ctx, dt, dk, (const uint16_t*)dsx, kx, (const uint16_t*)dsy, ky, (const float*)djx, (const float*)djy,
                          A, B, TX, TY, (double*)dout));
    VS_TRY(st.raw_out(out4, dout, 32));
    return st.finish();
}

int vs_image_warp_u8_f32(vs_ctx* ctx, const vs_img* in, const float* params4, const vs_img* out, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, in, "image_warp input")); VS_TRY(check_img(ctx, out, "image_warp output"));
    VS_REQUIRE(ctx, params4, "image_warp: params is NULL");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == VS_MEM_DEVICE) return vsk_image_warp(ctx, vs_dev_img(in), params4, vs_dev_img(out));
    int batch = in->batch < 1 ? 1 : in->batch;
    Stage st(ctx);
    st.want_img(in, 1); st.want_img(out, 4); st.want_raw((size_t)batch * 16);
    VS_TRY(st.reserve());
    VsDevImg din, dout; void* dp;
    VS_TRY(st.img(in, 1, 1, true, &din));
    VS_TRY(st.img(out, 4, 4, false, &dout));
    VS_TRY(st.raw(params4, (size_t)batch * 16, true, &dp));
    VS_TRY(vsk_image_warp(ctx, din, (const float*)dp, dout));
    VS_TRY(st.img_out(dout, out, 4, 4));
    return st.finish();
}

int vs_bgr_warp_u8(vs_ctx* ctx, const vs_img* src, const double* M6, const vs_img* dst,
                   int dst_x0, int dst_y0, int mode, int border, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, src, "bgr_warp source")); VS_TRY(check_img(ctx, dst, "bgr_warp destination"));
    VS_REQUIRE(ctx, M6, "bgr_warp: matrix is NULL");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    int batch = src->batch < 1 ? 1 : src->batch;
    std::vector<VsWarpCoef> coef(batch);
    for (int b = 0; b < batch; b++) vs_warp_coef_from_forward(M6 + 6 * b, &coef[b]);
    Stage st(ctx);
    st.want_raw(coef.size() * sizeof(VsWarpCoef));
    // fixed-point tables of the row-group kernel (cv-exact mode, constant border)
    const size_t tab_bytes = mode == VS_WARP_LANCZOS2 || border == VS_BORDER_CONSTANT0
                                 ? vs_warp_rows_tab_ints(dst->width, dst->height) * sizeof(int32_t) * batch : 0;
    st.want_raw(tab_bytes);
    if (mem == VS_MEM_HOST) { st.want_img(src, 3); st.want_img(dst, 3); }
    VS_TRY(st.reserve());
    void *dcoef, *dtab = nullptr;
    // pageable source: the copy is complete when cudaMemcpyAsync returns, so `coef` may go away
    VS_TRY(st.raw(coef.data(), coef.size() * sizeof(VsWarpCoef), true, &dcoef));
    if (tab_bytes) VS_TRY(st.raw(nullptr, tab_bytes, false, &dtab));
    if (mem == VS_MEM_DEVICE) {
        int r = vsk_bgr_warp(ctx, vs_dev_img(src), (const VsWarpCoef*)dcoef, vs_dev_img(dst), dst_x0, dst_y0, mode, border, (int32_t*)dtab);
        return r;
    }
    VsDevImg dsrc, ddst;
    VS_TRY(st.img(src, 3, 1, true, &dsrc));
    VS_TRY(st.img(dst, 3, 1, false, &ddst));
    VS_TRY(vsk_bgr_warp(ctx, dsrc, (const VsWarpCoef*)dcoef, ddst, dst_x0, dst_y0, mode, border, (int32_t*)dtab));
    VS_TRY(st.img_out(ddst, dst, 3, 1));
    return st.finish();
}

int vs_plane_warp_u8(vs_ctx* ctx, const vs_img* src, int channels, const double* M6, const vs_img* dst,
                     int dst_x0, int dst_y0, int mem)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_TRY(check_mem(ctx, mem)); VS_TRY(check_img(ctx, src, "plane_warp source")); VS_TRY(check_img(ctx, dst, "plane_warp destination"));
    VS_REQUIRE(ctx, M6, "plane_warp: matrix is NULL");
    VS_REQUIRE(ctx, channels == 1 || channels == 2, "plane_warp: 1 or 2 channels");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    int batch = src->batch < 1 ? 1 : src->batch;
    std::vector<VsWarpCoef> coef(batch);
    for (int b = 0; b < batch; b++) vs_warp_coef_from_forward(M6 + 6 * b, &coef[b]);
    Stage st(ctx);
    st.want_raw(coef.size() * sizeof(VsWarpCoef));
    if (mem == VS_MEM_HOST) { st.want_img(src, channels); st.want_img(dst, channels); }
    VS_TRY(st.reserve());
    void* dcoef;
    VS_TRY(st.raw(coef.data(), coef.size() * sizeof(VsWarpCoef), true, &dcoef));
    if (mem == VS_MEM_DEVICE)
        return vsk_plane_warp_slots(ctx, vs_dev_img(src), channels, nullptr, (const VsWarpCoef*)dcoef, vs_dev_img(dst), dst_x0, dst_y0);
    VsDevImg dsrc, ddst;
    VS_TRY(st.img(src, channels, 1, true, &dsrc));
    VS_TRY(st.img(dst, channels, 1, false, &ddst));
    VS_TRY(vsk_plane_warp_slots(ctx, dsrc, channels, nullptr, (const VsWarpCoef*)dcoef, ddst, dst_x0, dst_y0));
    VS_TRY(st.img_out(ddst, dst, channels, 1));
    return st.finish();
}

int vs_debug_invert4(vs_ctx* ctx, const double* H, int n, double* out_quad, double* out_serial, double* out_cond)
{
    if (!ctx) return VS_ERR_INVALID;
    VS_REQUIRE(ctx, H && out_quad && out_serial && out_cond && n >= 0, "debug_invert4: bad arguments");
    VS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t mb = (size_t)n * 16 * sizeof(double);
    Stage st(ctx);
    st.want_raw(mb); st.want_raw(mb); st.want_raw(mb); st.want_raw((size_t)n * sizeof(double));
    VS_TRY(st.reserve());
    void *dH, *dq, *ds, *dc;
    VS_TRY(st.raw(H, mb, true, &dH));
    VS_TRY(st.raw(nullptr, mb, false, &dq));
    VS_TRY(st.raw(nullptr, mb, false, &ds));
    VS_TRY(st.raw(nullptr, (size_t)n * sizeof(double), false, &dc));
    VS_TRY(vsk_debug_invert4(ctx, (const double*)dH, n, (double*)dq, (double*)ds, (double*)dc));
    VS_TRY(st.raw_out(out_quad, dq, mb));
    VS_TRY(st.raw_out(out_serial, ds, mb));
    VS_TRY(st.raw_out(out_cond, dc, (size_t)n * sizeof(double)));
    return st.finish();
}

}  // extern "C"
