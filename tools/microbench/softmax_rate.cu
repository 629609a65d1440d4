// Microbenchmark: how fast can the softmax warps of the attention kernel chew through S when nothing else is in the
// way (no tensor memory, no barriers)?  Each warp-iteration processes 32 rows x COLS columns exactly like the kernel:
// FMNMX3 running max, FFMA2 scale/shift, exp2 (MUFU, every POLY-th pair as a polynomial), FADD2 row sum, bf16x2 pack.
// WARPS warps per SM (one CTA per SM): 8 x 128 columns models one thread per row on two 128 x 128 tiles (2 warps per
// SM sub-partition), 16 x 64 columns models two threads per row (4 warps per sub-partition).  Reported: SM clocks per
// 128 x 128 tile.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_rate softmax_rate.cu && ./softmax_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)), "l"(*reinterpret_cast<const uint64_t *>(&c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return d;
}
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.0f); x.y = fmaxf(x.y, -126.0f);
    const float2 t = fadd2(x, make_float2(12582912.0f, 12582912.0f));
    const float2 r = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = ffma2(r, make_float2(-1.0f, -1.0f), x);
    float2 p = ffma2(make_float2(0.05517147481441498f, 0.05517147481441498f), f, make_float2(0.242610901594162f, 0.242610901594162f));
    p = ffma2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
    p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return p;
}
__device__ __forceinline__ void fresh32(uint32_t (&r)[32]) {   // "these 32 registers now hold new unknown values" (stands for tcgen05.ld)
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ void sink16(const uint32_t (&r)[16]) {   // stands for tcgen05.st of 16 packed columns
    asm volatile("" :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

// MODE 0: one pass (max tracked beside the exponentials, as vsum_attn2_tc05.cu); MODE 1: two passes (max over all COLS first)
template <int COLS, int POLY, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float *out, long long *clk, int iters, float c, float m0) {
    uint32_t s[COLS];
#pragma unroll
    for (int i = 0; i < COLS; ++i) s[i] = __float_as_uint(-0.01f * (float)(threadIdx.x + i));
    float l = 0.f, m_ref = m0, mall = -INFINITY;
    const float2 c2 = make_float2(c, c);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        float mx = -INFINITY;
        if (MODE == 1) {
#pragma unroll
            for (int ch = 0; ch < COLS / 32; ++ch) fresh32(*reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
#pragma unroll
            for (int i = 0; i < COLS; i += 2) mx = fmax3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
            if (mx * c > m_ref + 8.0f) m_ref = mx * c;
        }
        if (MODE == 0) m_ref = m0 + (float)it * 1e-9f;      // keeps the loop body from being hoisted (the empty asm is no barrier to that)
        const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
        for (int ch = 0; ch < COLS / 32; ++ch) {
            uint32_t(&sc)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]);
            uint32_t wv[16];
            if (MODE == 0) {
                fresh32(sc);
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(sc[i]), __uint_as_float(sc[i + 1]));
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 x = ffma2(make_float2(__uint_as_float(sc[2 * e]), __uint_as_float(sc[2 * e + 1])), c2, nm2);
                const bool poly = POLY > 0 && (e % (POLY > 0 ? POLY : 1)) == POLY - 1;
                const float2 p = poly ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                ps[e & 3] = fadd2(ps[e & 3], p);
                wv[e] = pack_bf16x2(p.x, p.y);
            }
            sink16(wv);
        }
        const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
        l += pq.x + pq.y;
        mall = fmaxf(mall, mx);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = l + mall;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int COLS, int POLY, int MODE, int WARPS>
void run(const char *name, float *d, long long *dc) {
    const int warps = WARPS;
    const int iters = 2000;
    k<COLS, POLY, MODE, WARPS><<<148, warps * 32>>>(d, dc, 10, 0.09f, 0.f);
    k<COLS, POLY, MODE, WARPS><<<148, warps * 32>>>(d, dc, iters, 0.09f, 0.f);
    long long h[148];
    cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    // one iteration of all warps covers warps * 32 rows x COLS columns = warps * COLS / 512 tiles of 128 x 128
    const double tiles_per_iter = warps * COLS / 512.0;
    printf("%-58s %2d warps/SM: %7.1f clk per warp-iteration, %7.1f clk per 128x128 tile per SM\n", name, warps, avg / iters, avg / iters / tiles_per_iter);
}

int main() {
    float *d; long long *dc;
    cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 148 * 8);
    run<128, 4, 0, 8>("one pass, 128 cols/thread, poly every 4th", d, dc);
    run<128, 4, 0, 4>("one pass, 128 cols/thread, poly every 4th", d, dc);
    run<128, 4, 0, 16>("one pass, 128 cols/thread, poly every 4th", d, dc);
    run<128, 0, 0, 8>("one pass, 128 cols/thread, no poly", d, dc);
    run<128, 2, 0, 8>("one pass, 128 cols/thread, poly every 2nd", d, dc);
    run<128, 3, 0, 8>("one pass, 128 cols/thread, poly every 3rd", d, dc);
    run<64, 4, 1, 16>("two pass, 64 cols/thread, poly every 4th", d, dc);
    run<64, 4, 1, 8>("two pass, 64 cols/thread, poly every 4th", d, dc);
    run<64, 0, 1, 16>("two pass, 64 cols/thread, no poly", d, dc);
    run<64, 3, 1, 16>("two pass, 64 cols/thread, poly every 3rd", d, dc);
    run<64, 2, 1, 16>("two pass, 64 cols/thread, poly every 2nd", d, dc);
    run<64, 4, 0, 16>("one pass, 64 cols/thread, poly every 4th", d, dc);
    run<64, 3, 0, 16>("one pass, 64 cols/thread, poly every 3rd", d, dc);
    run<32, 4, 0, 32>("one pass, 32 cols/thread, poly every 4th", d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
