// tma_probe.cu — stand-alone check of the 3-D tensor-map box load used by k_bgr_warp_cv_tma
// (u32 [slot][row][word] tensor, box {96, 20, 1}, negative / out-of-range coordinates zero-filled).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; run on a B200.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int BOXW = 96, BOXH = 20;

__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, uint32_t* out, int variant)
{
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ __align__(8) unsigned long long bar_mem;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_mem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(BOXW * BOXH * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        if (!done && ++spins > (1u << 22)) { if (threadIdx.x == 0) out[BOXW * BOXH] = 0xdeadbeefu; return; }
    }
    for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = sm[i];
    if (threadIdx.x == 0) out[BOXW * BOXH] = spins;
}

int main(int argc, char** argv)
{
    const int pitchw = 480, h = 360, slots = 3;   // 640x360 BGR, pitch 1920 bytes
    std::vector<uint32_t> host((size_t)pitchw * h * slots);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint32_t)i * 2654435761u;
    uint32_t *d, *dout;
    cudaMalloc(&d, host.size() * 4);
    cudaMalloc(&dout, (BOXW * BOXH + 1) * 4);
    cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    auto enc = (PFN_cuTensorMapEncodeTiled)p;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)pitchw, (cuuint64_t)h, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)pitchw * 4, (cuuint64_t)pitchw * 4 * h};
    const cuuint32_t box[3] = {BOXW, BOXH, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int cases[1][3] = {{atoi(argv[1]), atoi(argv[2]), atoi(argv[3])}};
    int bad_total = 0;
    for (auto& c : cases) {
        cudaMemset(dout, 0, (BOXW * BOXH + 1) * 4);
        k<<<1, 128, 28160>>>(map, c[0], c[1], c[2], dout, 0);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case (%d,%d,%d): %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 2; }
        std::vector<uint32_t> out(BOXW * BOXH + 1);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < BOXH; y++)
            for (int x = 0; x < BOXW; x++) {
                const int gx = c[0] + x, gy = c[1] + y;
                const uint32_t want = (gx < 0 || gx >= pitchw || gy < 0 || gy >= h) ? 0u : host[((size_t)c[2] * h + gy) * pitchw + gx];
                if (out[y * BOXW + x] != want) bad++;
            }
        printf("case (%d,%d,%d): %s, flag/spins=%u, mismatches=%d\n", c[0], c[1], c[2], cudaGetErrorString(e), out[BOXW * BOXH], bad);
        bad_total += bad;
    }
    return bad_total ? 1 : 0;
}
