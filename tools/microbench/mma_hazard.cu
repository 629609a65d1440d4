// Microbenchmark: cost of the attention kernel's tensor-memory dependencies inside the tcgen05 pipe.  One thread issues,
// per "unit", 8 TS MMAs (PV: A = P read from an S buffer, D = O) and 4 SS MMAs (QK^T: D = an S buffer), 512 clk of
// tensor work, in three arrangements:
//   0  PV reads buffer (u % 3), QK writes buffer ((u + 1) % 3): no dependency between neighbours
//   1  PV reads buffer (u % 3), then QK writes THE SAME buffer (write-after-read, the three-buffer rotation)
//   2  QK writes buffer (u % 3), then PV reads THE SAME buffer (read-after-write; never happens in the kernel without a softmax between)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../video-summarization_b200/csrc -I../../include -o mma_hazard mma_hazard.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vsum_tc05.cuh"
using namespace vsum;
constexpr int UNITS = 600;

template <int MODE, int PER_COMMIT>
__global__ void __launch_bounds__(128, 1) k(float *clk_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ (uint32_t)(i * 2654435761u >> 20);
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (warp == 1) { tc::tmem_alloc(slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        constexpr uint32_t IDESC_QK = tc::make_idesc(1, 128, 128, 0, 0), IDESC_PV = tc::make_idesc(1, 128, 64, 0, 1);
        const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
        const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem + 32 * 1024), 16, 1024);
        long long t0 = 0, t1 = 0;
        uint32_t ph = 0;
        for (int pass = 0; pass < 2; ++pass) {
            t0 = clock64();
            for (int u = 0; u < UNITS; ++u) {
                const uint32_t rb = tmem + (uint32_t)(u % 3) * 128, wb = tmem + (uint32_t)((MODE == 0 ? u + 1 : u) % 3) * 128;
                if (tc::elect_one()) {
                    if (MODE == 2) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) tc::mma_f16_ss(wb, a_desc + (uint64_t)(k2 * 2), b_desc + (uint64_t)(k2 * 2), IDESC_QK, k2 != 0);
                    }
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) tc::mma_f16_ts(tmem + 384 + (u & 1) * 64, rb + k2 * 8, b_desc + (uint64_t)(k2 * 128), IDESC_PV, 1);
                    if (MODE != 2) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) tc::mma_f16_ss(wb, a_desc + (uint64_t)(k2 * 2), b_desc + (uint64_t)(k2 * 2), IDESC_QK, k2 != 0);
                    }
                    if ((u + 1) % PER_COMMIT == 0) tc::mma_commit(bar);
                }
                __syncwarp();
                if ((u + 1) % PER_COMMIT == 0) { tc::mbar_wait(bar, ph); ph ^= 1; }
            }
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) clk_out[0] = (float)(t1 - t0) / UNITS;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int MODE, int PER_COMMIT>
static void run(const char *what, float *d_clk) {
    auto kern = k<MODE, PER_COMMIT>;
    const int smem = 96 * 1024 + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<148, 128, smem>>>(d_clk);
    cudaError_t e = cudaDeviceSynchronize();
    float clk = 0;
    cudaMemcpy(&clk, d_clk, 4, cudaMemcpyDeviceToHost);
    printf("%-64s commit+wait every %3d units: %7.1f clk per unit (512 of tensor work) %s\n", what, PER_COMMIT, clk, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    float *d; cudaMalloc(&d, 4);
    run<0, 100>("PV reads buffer u, QK writes buffer u+1 (independent)", d);
    run<1, 100>("PV reads buffer u, QK overwrites buffer u (write after read)", d);
    run<2, 100>("QK writes buffer u, PV reads buffer u (read after write)", d);
    run<0, 1>("PV reads buffer u, QK writes buffer u+1 (independent)", d);
    run<1, 1>("PV reads buffer u, QK overwrites buffer u (write after read)", d);
    run<2, 1>("QK writes buffer u, PV reads buffer u (read after write)", d);
    return 0;
}
