// Microbenchmark: throughput of DADD, DSETP (+ 64-bit select) and LDS.64 on sm_100a -- the knapsack row update's mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters, double seed) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = seed * threadIdx.x; sm[threadIdx.x + 512] = seed;
    __syncthreads();
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = a[i] + seed;                                   // DADD
            if (MODE == 1) a[i] = (a[i] >= seed + i) ? a[i] : seed;               // DSETP + select (no DADD: seed + i hoisted)
            if (MODE == 2) { const double b = a[i] + seed; a[i] = !(a[i] >= b) ? b : a[i]; }   // DADD + DSETP + select
            if (MODE == 3) a[i] += sm[(threadIdx.x + i * 37 + it) & 511];         // LDS.64 + DADD
            if (MODE == 4) { const long long x = __double_as_longlong(a[i]), y = __double_as_longlong(seed) + i; a[i] = __longlong_as_double(x >= y ? x + 1 : y); }  // 64-bit int compare/select
        }
    }
    for (int i = 0; i < 8; ++i) acc += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> float run(double *d, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, 512>>>(d, 10, 1.0000001);
    cudaEventRecord(a); k<MODE><<<148, 512>>>(d, iters, 1.0000001); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double *d; cudaMalloc(&d, 148 * 512 * 8);
    const int iters = 20000;
    const double ops = 512.0 * 8 * iters;     // thread-level loop bodies per SM (one 512-thread CTA per SM, like the knapsack kernel)
    const char *names[5] = {"DADD", "DSETP+select", "DADD+DSETP+select", "LDS.64+DADD", "int64 compare+select"};
    float t[5] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters)};
    for (int m = 0; m < 5; ++m) printf("%-22s: %8.3f ms  -> %.1f thread-ops / clk / SM @1.965 GHz\n", names[m], t[m], ops / (t[m] * 1e-3 * 1.965e9));
    return 0;
}
