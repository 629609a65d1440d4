// Microbenchmark: does the softmax arithmetic of the attention kernel slow down when (a) the tensor pipe is busy with
// the kernel's MMA shapes and (b) its S / P traffic really goes through tensor memory?  One CTA per SM: warp 0 issues
// tcgen05.mma groups (4 x SS M128 N128 = QK^T, 8 x TS M128 N64 = PV) back to back until the softmax warps are done;
// SW softmax warps run the register arithmetic of vsum_attn2_tc05.cu on COLS columns per thread per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../video-summarization_b200/csrc -o interference interference.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "vsum_tc05.cuh"

using namespace vsum;

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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr) : "memory");
}
__device__ __forceinline__ void keep16(const uint32_t (&r)[16]) {
    asm volatile("" :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

// SW softmax warps (warps 4 .. 4+SW), COLS columns per thread per iteration (in chunks of 32), TMEM: real tcgen05.ld / st
template <int SW, int COLS, bool TMEM, bool MMA, bool PIPE = false>
__global__ void __launch_bounds__((4 + SW) * 32, 1) k(float *out, long long *clk, long long *mma_count, int iters, float c) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    volatile int *done = reinterpret_cast<volatile int *>(slot + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ (uint32_t)(i * 2654435761u >> 20);
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); *done = 0; }
    if (warp == 1) { tc::tmem_alloc(slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (MMA) {
            constexpr uint32_t IDESC_QK = tc::make_idesc(1, 128, 128, 0, 0), IDESC_PV = tc::make_idesc(1, 128, 64, 0, 1);
            const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
            const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem + 32 * 1024), 16, 1024);
            long long groups = 0;
            uint32_t ph = 0;
            while (*done < SW) {
                if (tc::elect_one()) {
                    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) tc::mma_f16_ss(tmem + (rep & 1) * 128, a_desc + (uint64_t)(k2 * 2), b_desc + (uint64_t)(k2 * 2), IDESC_QK, k2 != 0);
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) tc::mma_f16_ts(tmem + 256 + (rep & 1) * 64, tmem + 384 + (rep & 1) * 64 + k2 * 8, b_desc + (uint64_t)(k2 * 128), IDESC_PV, 1);
                    }
                    tc::mma_commit(bar);
                }
                __syncwarp();
                tc::mbar_wait(bar, ph);
                ph ^= 1;
                groups += 4;
            }
            if (lane == 0) mma_count[blockIdx.x] = groups;
        }
    } else if (warp >= 4) {
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int grp = (warp - 4) >> 2;                         // which column range of S / P this warp owns
        const uint32_t tS = tmem + lane_off + (uint32_t)((grp * COLS) & 255), tP = tmem + 384 + lane_off + (uint32_t)((grp * COLS / 2) & 127);
        uint32_t s[COLS];
#pragma unroll
        for (int i = 0; i < COLS; ++i) s[i] = __float_as_uint(-0.01f * (float)(threadIdx.x + i));
        float l = 0.f, mall = -INFINITY;
        const float2 c2 = make_float2(c, c);
        if (TMEM && !MMA) {      // benign S values in tensor memory (with MMA on, the products of the fill pattern are small finite numbers too)
#pragma unroll
            for (int ch = 0; ch < COLS / 32; ++ch) tc::tmem_st32(tS + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
            tc::tmem_wait_st();
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            float mx = -INFINITY;
            if (PIPE) {          // the loop of the one-thread-per-row kernel: chunk c+1 loads while chunk c is exponentiated
                const float m_ref = (float)it * 1e-9f;
                const float2 nm2 = make_float2(-m_ref, -m_ref);
                uint32_t(&sa)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                uint32_t(&sb)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                auto chunk = [&](uint32_t (&sc)[32], int ch) {
                    uint32_t wv[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(sc[i]), __uint_as_float(sc[i + 1]));
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 x = ffma2(make_float2(__uint_as_float(sc[2 * e]), __uint_as_float(sc[2 * e + 1])), c2, nm2);
                        const float2 p = (e % 4 == 3) ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                        ps[e & 3] = fadd2(ps[e & 3], p);
                        wv[e] = pack_bf16x2(p.x, p.y);
                    }
                    tmem_st16(tP + ch * 16, wv);
                };
                auto wait_on = [](uint32_t (&r)[32]) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
                };
                tc::tmem_ld32(tS, sa); wait_on(sa);
                tc::tmem_ld32(tS + 32, sb); chunk(sa, 0); wait_on(sb);
                tc::tmem_ld32(tS + 64, sa); chunk(sb, 1); wait_on(sa);
                tc::tmem_ld32(tS + 96, sb); chunk(sa, 2); wait_on(sb);
                chunk(sb, 3);
                tc::tmem_wait_st();
                const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
                l += pq.x + pq.y;
                mall = fmaxf(mall, mx);
                continue;
            }
            if (TMEM) {
#pragma unroll
                for (int ch = 0; ch < COLS / 32; ++ch) tc::tmem_ld32(tS + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
                tc::tmem_wait_ld();
            }
            const float m_ref = (float)it * 1e-9f;
            const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
            for (int i = 0; i < COLS; i += 2) mx = fmax3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
#pragma unroll
            for (int ch = 0; ch < COLS / 32; ++ch) {
                uint32_t wv[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 x = ffma2(make_float2(__uint_as_float(s[ch * 32 + 2 * e]), __uint_as_float(s[ch * 32 + 2 * e + 1])), c2, nm2);
                    const float2 p = (e % 4 == 3) ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                    ps[e & 3] = fadd2(ps[e & 3], p);
                    wv[e] = pack_bf16x2(p.x, p.y);
                }
                if (TMEM) tmem_st16(tP + ch * 16, wv); else keep16(wv);
            }
            if (TMEM) tc::tmem_wait_st();
            const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
            l += pq.x + pq.y;
            mall = fmaxf(mall, mx);
        }
        const long long t1 = clock64();
        out[blockIdx.x * blockDim.x + threadIdx.x] = l + mall;
        if (lane == 0) { if (warp == 4) clk[blockIdx.x] = t1 - t0; atomicAdd((int *)done, 1); }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int SW, int COLS, bool TMEM, bool MMA, bool PIPE = false>
void run(const char *name, float *d, long long *dc, long long *dm) {
    const int iters = 1000;
    auto kern = k<SW, COLS, TMEM, MMA, PIPE>;
    const int smem = 96 * 1024 + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(dm, 0, 148 * 8);
    kern<<<148, (4 + SW) * 32, smem>>>(d, dc, dm, iters, 0.09f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148], m[148];
    cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(m, dm, sizeof(m), cudaMemcpyDeviceToHost);
    double avg = 0, mm = 0; for (int i = 0; i < 148; ++i) { avg += (double)h[i]; mm += (double)m[i]; } avg /= 148; mm /= 148;
    const double tiles_per_iter = SW * COLS / 512.0;
    printf("%-34s tmem %d mma %d: %7.1f clk per 128x128 tile of softmax per SM; tensor pipe busy %4.0f %% (%.0f clk of MMA per softmax tile) %s\n", name, (int)TMEM, (int)MMA,
           avg / iters / tiles_per_iter, 100.0 * mm * 512.0 / avg, mm * 512.0 / (iters * tiles_per_iter), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float *d; long long *dc, *dm;
    cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 148 * 8); cudaMalloc(&dm, 148 * 8);
    run<8, 128, false, false>("8 warps x 128 cols", d, dc, dm);
    run<8, 128, true, false>("8 warps x 128 cols", d, dc, dm);
    run<8, 128, false, true>("8 warps x 128 cols", d, dc, dm);
    run<8, 128, true, true>("8 warps x 128 cols", d, dc, dm);
    run<8, 128, true, false, true>("8 warps x 128 cols, pipelined ld", d, dc, dm);
    run<8, 128, true, true, true>("8 warps x 128 cols, pipelined ld", d, dc, dm);
    run<16, 64, false, false>("16 warps x 64 cols", d, dc, dm);
    run<16, 64, true, false>("16 warps x 64 cols", d, dc, dm);
    run<16, 64, false, true>("16 warps x 64 cols", d, dc, dm);
    run<16, 64, true, true>("16 warps x 64 cols", d, dc, dm);
    return 0;
}
