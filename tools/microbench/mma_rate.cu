// Microbenchmark: sustained issue period of tcgen05.mma (kind::f16, bf16 operands, K = 16) for the tile shapes the
// attention kernels use, with the A operand in shared memory (SS) or in tensor memory (TS), and the rate of
// tcgen05.st / tcgen05.ld of 32 columns.  One CTA per SM, one elected thread issues; data are arbitrary bit patterns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../video-summarization_b200/csrc -o mma_rate mma_rate.cu && ./mma_rate
// Expectation from the data path: an SS MMA reads (M + N) x 32 bytes of operands from shared memory (128 B/clk/SM) and
// needs M x N x 32 / 8192 clk of tensor pipe; whichever is larger sets the period.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vsum_tc05.cuh"

using namespace vsum;

constexpr int REPS = 512;          // MMAs per measurement = REPS x 8

// MODE 0: SS, A and B K-major.  MODE 1: SS, B MN-major (the PV shape).  MODE 2: TS (A in TMEM), B MN-major.  MODE 3: TS, B K-major.
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) mma_kernel(float *clk_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ (uint32_t)(i * 2654435761u >> 20);   // small finite bf16 pairs
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (warp == 1) { tc::tmem_alloc(slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        constexpr uint32_t IDESC = tc::make_idesc(1, 128, N, 0, (MODE == 1 || MODE == 2) ? 1 : 0);
        const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);                 // [128 x 64] K-major tile(s)
        const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem + 32 * 1024), 16, 1024);   // up to [256 x 64]
        long long t0 = 0, t1 = 0;
        for (int pass = 0; pass < 2; ++pass) {            // pass 0 warms up
            t0 = clock64();
            if (tc::elect_one()) {
                for (int r = 0; r < REPS; ++r) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // k & 3: the four 32-byte K steps of a 128-byte row; k >> 2: a second tile 16 KB further on
                        const uint64_t ad = a_desc + (uint64_t)((k & 3) * 2 + (k >> 2) * 1024);
                        const uint64_t bd = (MODE == 1 || MODE == 2) ? b_desc + (uint64_t)(k * 128) : b_desc + (uint64_t)((k & 3) * 2);
                        if (MODE >= 2) tc::mma_f16_ts(tmem, tmem + 256 + k * 8, bd, IDESC, 1);
                        else tc::mma_f16_ss(tmem, ad, bd, IDESC, 1);
                    }
                }
                tc::mma_commit(bar);
            }
            __syncwarp();
            tc::mbar_wait(bar, pass & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) clk_out[0] = (float)(t1 - t0) / (REPS * 8);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

// ST = true: tcgen05.st of 32 columns per warp back to back; false: tcgen05.ld.  Four warps (all lane quarters).
template <bool ST>
__global__ void __launch_bounds__(128, 1) tmem_rw_kernel(float *clk_out, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 32 + i;
    tc::tmem_st32(base, v);
    tc::tmem_wait_st();
    const long long t0 = clock64();
    for (int r = 0; r < REPS; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (ST) tc::tmem_st32(base + k * 32, v);
            else tc::tmem_ld32(base + k * 32, v);
        }
        if (ST) tc::tmem_wait_st(); else tc::tmem_wait_ld();
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk_out[0] = (float)(t1 - t0) / (REPS * 8);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(slot, 512); }
}

template <int MODE, int N>
static void run(const char *what, float *d_clk) {
    auto kern = mma_kernel<MODE, N>;
    const int smem = 96 * 1024 + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<148, 128, smem>>>(d_clk);
    cudaError_t e = cudaDeviceSynchronize();
    float clk = 0;
    cudaMemcpy(&clk, d_clk, 4, cudaMemcpyDeviceToHost);
    const double pipe = 128.0 * N * 32 / 8192, smem_clk = (MODE >= 2 ? N : 128 + N) * 32.0 / 128;
    printf("%-44s %7.1f clk / MMA   (tensor pipe %5.1f, shared-memory operand reads %5.1f)  %s\n", what, clk, pipe, smem_clk,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float *d_clk; uint32_t *sink;
    cudaMalloc(&d_clk, 4); cudaMalloc(&sink, 148 * 128 * 4);
    run<0, 128>("SS  M128 N128 K16 (QK^T)", d_clk);
    run<0, 256>("SS  M128 N256 K16 (GEMM)", d_clk);
    run<0, 64>("SS  M128 N64  K16, B K-major", d_clk);
    run<1, 64>("SS  M128 N64  K16, B MN-major (PV)", d_clk);
    run<2, 64>("TS  M128 N64  K16, A in TMEM, B MN-major (PV)", d_clk);
    run<3, 128>("TS  M128 N128 K16, A in TMEM (Q in TMEM)", d_clk);
    run<3, 256>("TS  M128 N256 K16, A in TMEM", d_clk);
    float clk;
    tmem_rw_kernel<true><<<148, 128>>>(d_clk, sink);
    cudaDeviceSynchronize(); cudaMemcpy(&clk, d_clk, 4, cudaMemcpyDeviceToHost);
    printf("tcgen05.st 32x32b.x32, 4 warps                %7.1f clk / instruction / warp\n", clk);
    tmem_rw_kernel<false><<<148, 128>>>(d_clk, sink);
    cudaDeviceSynchronize(); cudaMemcpy(&clk, d_clk, 4, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld 32x32b.x32, 4 warps                %7.1f clk / instruction / warp\n", clk);
    return 0;
}
