// Microbenchmark: does the fp32x2 -> bf16x2 pack (F2FP.BF16.F32.PACK_AB) share a pipe with MUFU.EX2 on sm_100a?
// Modes: 0 = ex2 only, 1 = cvt.rn.bf16x2.f32 only, 2 = both interleaved 2:1 (the softmax mix), 3 = ex2 + PRMT truncation,
//        4 = ex2 + FFMA2, 5 = cvt.rn.f16x2.f32 only
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pack_rate pack_rate.cu && ./pack_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(uint32_t *out, int iters) {
    float a[8]; uint32_t p[4];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) p[i] = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (MODE == 1 || MODE == 2) { uint32_t t; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[2 * i + 1]), "f"(a[2 * i])); p[i] ^= t; }
            if (MODE == 5) { uint32_t t; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[2 * i + 1]), "f"(a[2 * i])); p[i] ^= t; }
            if (MODE == 3) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(t) : "r"(__float_as_uint(a[2 * i])), "r"(__float_as_uint(a[2 * i + 1]))); p[i] ^= t; }
            if (MODE == 4) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(*reinterpret_cast<uint64_t *>(&a[2 * i])) : "l"(*reinterpret_cast<uint64_t *>(&a[2 * i]))); }
        }
        if (MODE == 1 || MODE == 5) { for (int i = 0; i < 8; ++i) a[i] += 1.0f; }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 4; ++i) acc ^= p[i];
    for (int i = 0; i < 8; ++i) acc ^= __float_as_uint(a[i]);
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
    const double groups = 148.0 * 4 * 512 * (double)iters;      // thread-level loop iterations (8 exps + 4 packs each)
    const char *names[6] = {"8 ex2", "4 cvt.bf16x2 (+8 fadd)", "8 ex2 + 4 cvt.bf16x2", "8 ex2 + 4 prmt", "8 ex2 + 4 ffma2", "4 cvt.f16x2 (+8 fadd)"};
    float t[6] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters)};
    for (int m = 0; m < 6; ++m)
        printf("%-26s: %8.3f ms  %.2f clk per thread-iteration per SMSP-lane-group (SM clocks per 32-thread iteration: %.2f)\n", names[m], t[m],
               0.0, t[m] * 1e-3 * 1.9e9 / (groups / 148 / 32) * 4);
    return 0;
}
