// Microbenchmark: issue rate of MUFU.EX2 for f32 vs packed bf16x2 / f16x2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(uint32_t *out, int iters) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = 0x3c003c00u + threadIdx.x + i;   // small positive halves / floats
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
            if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
            if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(uint32_t *d, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 4, 512>>>(d, 10);
    cudaEventRecord(a);
    k<MODE><<<148 * 4, 512>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    uint32_t *d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    const int iters = 20000;
    const double instr = 148.0 * 4 * 512 * 8.0 * iters;     // thread-level MUFU instructions
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters);
    printf("ex2.f32     : %.3f ms  %.1f Ginstr/s (thread-level)  %.2f per clk per SM @1.9GHz\n", t0, instr / t0 / 1e6, instr / (t0 * 1e-3) / 148 / 1.9e9);
    printf("ex2.bf16x2  : %.3f ms  %.1f Ginstr/s  %.2f per clk per SM  (x2 results)\n", t1, instr / t1 / 1e6, instr / (t1 * 1e-3) / 148 / 1.9e9);
    printf("ex2.f16x2   : %.3f ms  %.1f Ginstr/s  %.2f per clk per SM  (x2 results)\n", t2, instr / t2 / 1e6, instr / (t2 * 1e-3) / 148 / 1.9e9);
    return 0;
}
