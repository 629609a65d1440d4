// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them
// (one CTA per SM, 512 TMEM columns; warp w touches lane quarter w % 4).  Reports bytes per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rate tmem_rate.cu && ./tmem_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
          "r"(taddr) : "memory");
}

// MODE 0: loads only (4 in flight, then wait); 1: stores only; 2: one load + wait per iteration (latency-bound, as a softmax warp sees it)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t *out, long long *clk, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32], acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    tmem_st32(base, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)(((it + warp) * 32) & 511);
        if (MODE == 0) {
            uint32_t a[32], b[32];
            tmem_ld32(base + col, a);
            tmem_ld32(base + ((col + 128) & 511), b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += a[0] ^ a[31] ^ b[5] ^ b[17];
        } else if (MODE == 2) {
            uint32_t a[32];
            tmem_ld32(base + col, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += a[0] ^ a[31];
        } else {
            tmem_st32(base + col, r);
            tmem_st32(base + ((col + 128) & 511), r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char *name, int warps, uint32_t *d, long long *dc) {
    const int iters = 4000;
    k<MODE><<<148, warps * 32>>>(d, dc, 10);
    k<MODE><<<148, warps * 32>>>(d, dc, iters);
    long long h[148];
    cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    const double bytes = (double)iters * warps * 32 * 32 * 4 * (MODE == 2 ? 1 : 2);
    printf("%-44s %2d warps/SM: %8.1f clk per iteration, %7.1f B per clk per SM\n", name, warps, avg / iters, bytes / avg);
}

int main() {
    uint32_t *d; long long *dc;
    cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 148 * 8);
    for (int w : {4, 8, 16, 32}) run<0>("tcgen05.ld x32, two in flight", w, d, dc);
    for (int w : {4, 8, 16, 32}) run<2>("tcgen05.ld x32, one at a time", w, d, dc);
    for (int w : {4, 8, 16}) run<1>("tcgen05.st x32, two in flight", w, d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
