// Variable-length (packed, padding-free) multi-head self-attention for the scorer on sm_100a.
// Replaces src/model/simnet.py:155-161 (QK^T * d_model^-0.5 -> softmax -> PV) for d_model 256,
// 4 heads of 64.  No [N,N] score tensor is ever materialised and nothing is copied to the host
// (the reference does both, simnet.py:155,164).
//
// One CTA = one (video, head, 128-query tile); two CTAs per SM so one CTA's softmax overlaps the
// other's MMAs.  Roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator,
// warps 4..7 softmax (one query row per thread).
//   S = Q K^T      : A = Q tile (K-major, SW128), B = K tile (K-major, SW128)   -> TMEM [128 x 128] fp32
//   softmax        : tcgen05.ld S -> registers, online max/sum in the exp2 domain, P -> bf16 ->
//                    128B-swizzled shared memory (K-major A operand of the second MMA)
//   O_h += P_h V_h : B = V tile as loaded by TMA ([key][head_dim], i.e. MN-major, SW128); one fp32
//                    accumulator [128 x 64] in TMEM per 64-key half h of the tiles, merged in the epilogue
// Packed layout: a tile may read rows of the next video (or TMA zero fill past T); those key
// columns are masked to -inf and those query rows are never stored.
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

#include <cstdlib>

namespace vsum {
namespace {

constexpr int HD = 64;                  // head dim
constexpr int DM = 256;                 // d_model
constexpr int NH = 4;
constexpr int BQ = 128, BKV = 128;
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16 = 16 KB
constexpr int ATT_THREADS = 384;        // 4 control warps + 8 softmax warps
constexpr int ATT_TMEM_COLS = 256;      // S: [0,128)  O of key half 0: [128,192)  O of key half 1: [192,256)
constexpr size_t ATT_SMEM = 7 * (size_t)TILE_BYTES + 128;   // Q, K, V, P (2 buffers x 2 halves) + barriers: 2 CTAs / SM

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Packed fp32x2 arithmetic (sm_100): halves the FFMA / FADD issue slots of the softmax.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)),
          "l"(*reinterpret_cast<const uint64_t *>(&c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return d;
}
// exp2 on the FMA pipe for a share of the elements (the MUFU pipe is the limiter at head_dim 64:
// 16 ex2 / clk / SM, profiles/r01_microbench_mufu_ex2.txt).  Round-to-nearest split x = n + f with the
// 1.5 * 2^23 trick, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5,
// 50x below the bf16 rounding of P), then n is added into the exponent field with one IMAD.
#ifndef VSUM_ATTN_POLY_EVERY
#define VSUM_ATTN_POLY_EVERY 5      // every k-th pair of exponentials goes to the FMA pipe (0 = none); swept 2..9 on B200, 5 is best
#endif
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
    float2 t, r, f, p;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t *>(&t))
        : "l"(*reinterpret_cast<const uint64_t *>(&x)), "l"(*reinterpret_cast<const uint64_t *>(&magic)));
    const float2 nmagic = make_float2(-12582912.0f, -12582912.0f);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t *>(&r))
        : "l"(*reinterpret_cast<const uint64_t *>(&t)), "l"(*reinterpret_cast<const uint64_t *>(&nmagic)));
    const float2 mone = make_float2(-1.0f, -1.0f);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<uint64_t *>(&f))        // f = x - r
        : "l"(*reinterpret_cast<const uint64_t *>(&r)), "l"(*reinterpret_cast<const uint64_t *>(&mone)),
          "l"(*reinterpret_cast<const uint64_t *>(&x)));
    const float2 c3 = make_float2(0.05517147481441498f, 0.05517147481441498f);
    const float2 c2 = make_float2(0.242610901594162f, 0.242610901594162f);
    const float2 c1 = make_float2(0.6932609677314758f, 0.6932609677314758f);
    const float2 c0 = make_float2(0.9999281167984009f, 0.9999281167984009f);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<uint64_t *>(&p))
        : "l"(*reinterpret_cast<const uint64_t *>(&c3)), "l"(*reinterpret_cast<const uint64_t *>(&f)),
          "l"(*reinterpret_cast<const uint64_t *>(&c2)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "+l"(*reinterpret_cast<uint64_t *>(&p))
        : "l"(*reinterpret_cast<const uint64_t *>(&p)), "l"(*reinterpret_cast<const uint64_t *>(&f)),
          "l"(*reinterpret_cast<const uint64_t *>(&c1)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "+l"(*reinterpret_cast<uint64_t *>(&p))
        : "l"(*reinterpret_cast<const uint64_t *>(&p)), "l"(*reinterpret_cast<const uint64_t *>(&f)),
          "l"(*reinterpret_cast<const uint64_t *>(&c0)));
    // 2^n: (bits(t) << 23) is n << 23 modulo 2^32 (the magic constant's own bits shift out)
    p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return p;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// One CTA = (video, head, 128 queries); 2 CTAs / SM.  Each query row is shared by TWO softmax
// threads (64 key columns each, warps q+4 and q+8 of the same TMEM lane quarter), which puts four
// softmax warps on every SM sub-partition.  The two threads never talk inside the KV loop: each
// 64-key half of a row has its own exponent reference, row sum and its own O accumulator in TMEM
// (O0: keys 0..63 of every tile, O1: keys 64..127; the PV MMA of a tile is split accordingly), and the
// halves are merged once in the epilogue: O = (O0 2^(m0-m) + O1 2^(m1-m)) / (l0 2^(m0-m) + l1 2^(m1-m)).
// (A per-tile row-max exchange through TMEM + a named barrier cost ~450 of 2800 clk per tile.)
// The exponent reference only moves when the half-row max grew by more than 2^8 (lazy rescale), so
// the read-modify-write of O is rare.
// TRAIN: additionally writes the log-sum-exp of every (row, head) in the exp2 domain for the backward
// kernel (vsum_attn_bwd_tc05.cu) and applies dropout to P (simnet.py:159); the row sum stays un-dropped.
template <bool TRAIN>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t *__restrict__ cu,
                 const int32_t *__restrict__ tile_video, const int32_t *__restrict__ tile_q0,
                 const int32_t *__restrict__ n_tiles_ptr, __nv_bfloat16 *__restrict__ out,
                 float scale_log2e,
                 float *__restrict__ lse2, float keep_scale, uint32_t drop_thresh16, unsigned long long seed) {
    if ((int)blockIdx.x >= __ldg(n_tiles_ptr)) return;
    extern __shared__ __align__(1024) uint8_t smem[];    // no alignment slack: it would cost the 2nd CTA / SM
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *sQ = smem;
    // K and V are single-buffered (each is free again long before its successor is needed); the 32 KB
    // this saves double-buffers P, so the exponentials of tile j+1 never wait for the MMAs of tile j.
    uint8_t *sK = smem + TILE_BYTES, *sV = smem + 2 * (size_t)TILE_BYTES;
    uint8_t *sP = smem + 3 * (size_t)TILE_BYTES;      // 2 buffers x two 64-key halves of 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 7 * (size_t)TILE_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4,
             *s_full = bars + 5, *s_empty = bars + 6, *p_full = bars + 7, *p_empty = bars + 9;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vid = __ldg(tile_video + blockIdx.x), q0 = __ldg(tile_q0 + blockIdx.x);
    const int h_idx = blockIdx.y;
    const int base = __ldg(cu + vid), n = __ldg(cu + vid + 1) - base;
    const int nkv = (n + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&tmQKV);
    if (warp == 1 && lane == 0) {
        tc::mbar_init(q_full, 1);
        tc::mbar_init(k_full, 1); tc::mbar_init(k_empty, 1); tc::mbar_init(v_full, 1); tc::mbar_init(v_empty, 1);
        tc::mbar_init(s_full, 1); tc::mbar_init(s_empty, 256);
        for (int b = 0; b < 2; ++b) { tc::mbar_init(p_full + b, 256); tc::mbar_init(p_empty + b, 1); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;          // tO + 64 * half: one accumulator per 64-key half

    if (warp < 4) {
        tc::setmaxnreg_dec<32>();
        if (warp == 0 && lane == 0) {  // ===== TMA producer: Q, then the K tiles =====
            tc::mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tc::tma_load_2d(sQ, &tmQKV, q_full, h_idx * HD, base + q0);
            for (int j = 0; j < nkv; ++j) {
                tc::mbar_wait(k_empty, (j & 1) ^ 1);                 // QK(j-1) has consumed K
                tc::mbar_arrive_expect_tx(k_full, TILE_BYTES);
                tc::tma_load_2d(sK, &tmQKV, k_full, DM + h_idx * HD, base + j * BKV);
            }
        } else if (warp == 3 && lane == 0) {  // ===== TMA producer: the V tiles =====
            for (int j = 0; j < nkv; ++j) {
                tc::mbar_wait(v_empty, (j & 1) ^ 1);                 // PV(j-1) has consumed V
                tc::mbar_arrive_expect_tx(v_full, TILE_BYTES);
                tc::tma_load_2d(sV, &tmQKV, v_full, 2 * DM + h_idx * HD, base + j * BKV);
            }
        } else if (warp == 1) {  // ===== MMA issuer =====
            // The WHOLE warp runs this role with warp-uniform control flow and one elected lane issuing
            // the tcgen05 instructions, so descriptors live in uniform registers.  (Under `lane == 0`
            // every MMA cost ~160 cycles of R2UR / spill traffic: measured with clock64 stamps, the
            // issuing thread -- not the tensor pipe or the MUFU -- was the limiter.)
            constexpr uint32_t IDESC_QK = tc::make_idesc(1, BQ, BKV, 0, 0);   // S[128x128], both K-major
            constexpr uint32_t IDESC_PV = tc::make_idesc(1, BQ, HD, 0, 1);    // O[128x64], B (=V) MN-major
            const uint64_t q_desc = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);   // + 2 per 32-byte K step
            const uint64_t k_desc = tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024);
            // V operand (MN-major, 128B swizzle): 8-key groups are 1024 bytes apart (SBO); the 64-wide head dim is a single
            // swizzle atom so LBO is unused; one MMA K step = 16 keys = 2048 bytes
            const uint64_t v_desc = tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024);
            const uint64_t p_desc0 = tc::make_smem_desc_sw128(tc::smem_u32(sP), 16, 1024);
            constexpr uint32_t v_step = 2048 >> 4;
            // Per-MMA descriptors are re-derived at the point of use (one 32-bit add on the address field; the add is
            // opaque to the compiler).  Hoisted out of the KV loop, the 24 descriptors of a tile do not fit the
            // 32 registers this warp keeps after setmaxnreg.dec and came back from local memory before every MMA group.
            const uint32_t q_lo = (uint32_t)q_desc, k_lo = (uint32_t)k_desc, v_lo = (uint32_t)v_desc, p_lo0 = (uint32_t)p_desc0;
            const uint32_t hi_qkp = (uint32_t)(q_desc >> 32), hi_v = (uint32_t)(v_desc >> 32);
            auto desc_at = [](uint32_t lo, uint32_t off, uint32_t hi) -> uint64_t {
                uint32_t l;
                asm volatile("add.u32 %0, %1, %2;" : "=r"(l) : "r"(lo), "r"(off));
                return ((uint64_t)hi << 32) | l;
            };
            auto issue_qk = [&]() {
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        tc::mma_f16_ss(tS, desc_at(q_lo, k * 2, hi_qkp), desc_at(k_lo, k * 2, hi_qkp), IDESC_QK, k != 0);
                    tc::mma_commit(s_full);
                    tc::mma_commit(k_empty);
                }
                __syncwarp();
            };
            tc::mbar_wait(q_full, 0);
            tc::mbar_wait(k_full, 0);
            tc::tc_fence_after();
            issue_qk();
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {   // S(j+1) as soon as the softmax warps have S(j) in registers
                    tc::mbar_wait(k_full, (j + 1) & 1);
                    tc::mbar_wait(s_empty, j & 1);
                    tc::tc_fence_after();
                    issue_qk();
                }
                tc::mbar_wait(p_full + (j & 1), (j >> 1) & 1);   // P(j) written and O rescaled where needed
                tc::mbar_wait(v_full, j & 1);
                tc::tc_fence_after();
                // A: P buffer (j&1): half k/4 (16 KB apart), 32-byte step inside the 128-byte swizzled row
                const uint32_t p_lo = p_lo0 + (uint32_t)((j & 1) * (2 * TILE_BYTES >> 4));
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)               // keys 0..63 -> O half 0, keys 64..127 -> O half 1
                        tc::mma_f16_ss(tO + (uint32_t)((k >> 2) * HD), desc_at(p_lo, (k >> 2) * (TILE_BYTES >> 4) + (k & 3) * 2, hi_qkp),
                                       desc_at(v_lo, k * v_step, hi_v), IDESC_PV, (j | (k & 3)) != 0);   // accumulates over all KV tiles
                    tc::mma_commit(v_empty);
                    tc::mma_commit(p_empty + (j & 1));           // also means "PV(j) done"
                }
                __syncwarp();
            }
        }
    } else {  // ===== softmax: two threads per query row, 64 key columns each =====
        tc::setmaxnreg_inc<104>();
        const int qd = warp & 3, hf = (warp - 4) >> 2;       // TMEM lane quarter, column half
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t tS_h = tS + lane_off + hf * 64, tO_mine = tO + lane_off + hf * HD;   // my key half's own accumulator
        const int pair_bar = 2 + qd;                          // named barrier of the two warps sharing these rows (epilogue only)
        float m_run = -INFINITY, l_part = 0.f;                // exponent reference (exp2 domain) and row sum of MY key half
        const uint32_t p_row_u32 = tc::smem_u32(sP) + (uint32_t)hf * TILE_BYTES + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        uint32_t p_off[8];                                    // swizzled 16-byte chunk offsets of this row
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) p_off[ch] = (uint32_t)((ch ^ (r & 7)) << 4);
        const float2 c2 = make_float2(scale_log2e, scale_log2e);

#ifdef VSUM_ATTN_TIMING
        long long tph[6] = {0, 0, 0, 0, 0, 0}, tmark = clock64();
#define TMARK(i) do { const long long _n = clock64(); tph[i] += _n - tmark; tmark = _n; } while (0)
#else
#define TMARK(i) do { } while (0)
#endif
        for (int j = 0; j < nkv; ++j) {
            uint32_t s[64];
            tc::mbar_wait(s_full, j & 1);
            TMARK(0);
            tc::tc_fence_after();
            {
                uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                tc::tmem_ld32(tS_h, s0);
                tc::tmem_ld32(tS_h + 32, s1);
            }
            tc::tmem_wait_ld();
            tc::tc_fence_before();
            tc::mbar_arrive(s_empty);
            TMARK(1);

            const int valid = n - j * BKV - hf * 64;      // keys of this half-tile inside the video
            if (valid < 64) {
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c >= valid) s[c] = 0xff800000u;   // -inf
            }
            float mx8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) mx8[e] = __uint_as_float(s[e]);
#pragma unroll
            for (int c = 8; c < 64; c += 8)
#pragma unroll
                for (int e = 0; e < 8; ++e) mx8[e] = fmaxf(mx8[e], __uint_as_float(s[c + e]));
            const float mxl = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                    fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))) * scale_log2e;
            // No exchange with the partner thread: each 64-key half of a row keeps its own exponent reference, row sum
            // and O accumulator in TMEM; the two halves are merged once, in the epilogue.
            const float mx = mxl;
            float alpha = 1.0f;
            const bool bump = mx > m_run + 8.0f;
            if (__any_sync(0xffffffffu, bump)) {
                if (bump) { alpha = ex2(m_run - mx); m_run = mx; }      // alpha = 0 on the first tile
            }
            TMARK(2);
            tc::mbar_wait(p_empty + (j & 1), ((j >> 1) & 1) ^ 1);    // PV(j-2) done: this P buffer is free
            TMARK(3);
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rare: rescale my accumulator (64 columns of my row)
                tc::mbar_wait(p_empty + ((j - 1) & 1), ((j - 1) >> 1) & 1);   // PV(j-1) done: O is stable
                tc::tc_fence_after();
#pragma unroll 1
                for (int hc = 0; hc < 2; ++hc) {
                    uint32_t t[32];
                    tc::tmem_ld32(tO_mine + hc * 32, t);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
                    tc::tmem_st32(tO_mine + hc * 32, t);
                    tc::tmem_wait_st();
                }
            }
            const float nm = m_run == -INFINITY ? 0.f : -m_run;      // no valid key in my half so far: exp2(-inf) = 0, not NaN
            const float2 nm2 = make_float2(nm, nm);
            const uint32_t p_buf = p_row_u32 + (uint32_t)(j & 1) * 2 * TILE_BYTES;
            float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
                float2 pv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = ffma2(make_float2(__uint_as_float(s[c + 2 * e]), __uint_as_float(s[c + 2 * e + 1])), c2, nm2);
                    if (VSUM_ATTN_POLY_EVERY > 0 && (((c >> 1) + e) % (VSUM_ATTN_POLY_EVERY > 0 ? VSUM_ATTN_POLY_EVERY : 1)) == (VSUM_ATTN_POLY_EVERY - 1))
                        pv[e] = exp2_poly2(x);
                    else
                        pv[e] = make_float2(ex2(x.x), ex2(x.y));
                    ps[e] = fadd2(ps[e], pv[e]);
                }
                uint32_t w[4] = {pack2(pv[0].x, pv[0].y), pack2(pv[1].x, pv[1].y), pack2(pv[2].x, pv[2].y), pack2(pv[3].x, pv[3].y)};
                if (TRAIN && drop_thresh16 != 0) {            // two 4-key groups per 8 columns; the row sum above stays un-dropped
                    const uint32_t th2 = drop_thresh16 * 0x00010001u;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {             // 16-bit lane k of the draw <-> key 4g+k <-> half-word of the packed pair
                        const int key = j * BKV + hf * 64 + c + 4 * g;
                        const unsigned long long z = dropout_bits64(seed, attn_drop_group_index(base + q0 + r, h_idx, NH, key >> 2));
                        w[2 * g] &= __vcmpgeu2((uint32_t)z, th2);
                        w[2 * g + 1] &= __vcmpgeu2((uint32_t)(z >> 32), th2);
                    }
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_buf + p_off[c >> 3]), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                             : "memory");
            }
            TMARK(4);
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            tc::mbar_arrive(p_full + (j & 1));
            TMARK(5);
            const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
            l_part = fmaf(l_part, alpha, pq.x + pq.y);
        }
#ifdef VSUM_ATTN_TIMING
        if (blockIdx.x == 700 && blockIdx.y == 1 && lane == 0)
            printf("warp %2d nkv %d | wait S %lld | tmem ld %lld | max+exchange %lld | wait P buf %lld | exps+stores %lld | fence+arrive %lld (clk per tile)\n",
                   warp, nkv, tph[0] / nkv, tph[1] / nkv, tph[2] / nkv, tph[3] / nkv, tph[4] / nkv, tph[5] / nkv);
#endif
        // epilogue: O / l for my 32 head-dim columns of this row
        tc::mbar_wait(p_empty + ((nkv - 1) & 1), ((nkv - 1) >> 1) & 1);   // last PV done
        tc::tc_fence_after();
        // merge the two key halves of the row: (m, l) through the Q tile's shared memory (dead after the last QK^T)
        float2 *xch = reinterpret_cast<float2 *>(sQ);
        xch[hf * 128 + r] = make_float2(m_run, l_part);
        tc::bar_sync(pair_bar, 64);
        const float2 oth = xch[(hf ^ 1) * 128 + r];
        const float m_all = fmaxf(m_run, oth.x);
        const float a_mine = m_run == -INFINITY ? 0.f : ex2(m_run - m_all), a_oth = oth.x == -INFINITY ? 0.f : ex2(oth.x - m_all);
        const float l_tot = l_part * a_mine + oth.y * a_oth;
        uint32_t t[32];
        {
            uint32_t to[32];
            tc::tmem_ld32(tO + lane_off + hf * HD + hf * 32, t);             // my 32 output columns of my half's accumulator
            tc::tmem_ld32(tO + lane_off + (hf ^ 1) * HD + hf * 32, to);      // ... and of the partner's
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a_mine + __uint_as_float(to[i]) * a_oth);
        }
        const float inv = (TRAIN ? keep_scale : 1.0f) / l_tot;
        if (TRAIN && hf == 0 && q0 + r < n) lse2[(int64_t)(base + q0 + r) * NH + h_idx] = m_all + log2f(l_tot);
        if (TRAIN) {   // fp32 output: the backward's delta = rowsum(dO o O) must not see a rounded O
            if (q0 + r < n) {
                float *dst = reinterpret_cast<float *>(out) + (int64_t)(base + q0 + r) * DM + h_idx * HD + hf * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4 *>(dst + i) = make_float4(__uint_as_float(t[i]) * inv, __uint_as_float(t[i + 1]) * inv,
                                                                       __uint_as_float(t[i + 2]) * inv, __uint_as_float(t[i + 3]) * inv);
            }
        } else if (q0 + r < n) {
            __nv_bfloat16 *dst = out + (int64_t)(base + q0 + r) * DM + h_idx * HD + hf * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 pk;
                pk.x = pack2(__uint_as_float(t[i + 0]) * inv, __uint_as_float(t[i + 1]) * inv);
                pk.y = pack2(__uint_as_float(t[i + 2]) * inv, __uint_as_float(t[i + 3]) * inv);
                pk.z = pack2(__uint_as_float(t[i + 4]) * inv, __uint_as_float(t[i + 5]) * inv);
                pk.w = pack2(__uint_as_float(t[i + 6]) * inv, __uint_as_float(t[i + 7]) * inv);
                *reinterpret_cast<uint4 *>(dst + i) = pk;
            }
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

// Tile list (video, first query row) for all videos, in input order.  One block.
__global__ void __launch_bounds__(1024)
attn_schedule_kernel(const int32_t *__restrict__ cu, int B, int32_t *__restrict__ tile_video,
                     int32_t *__restrict__ tile_q0, int32_t *__restrict__ n_tiles_out, int max_tiles) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int v0 = 0; v0 < B; v0 += 1024) {
        const int v = v0 + threadIdx.x;
        const int cnt = v < B ? (__ldg(cu + v + 1) - __ldg(cu + v) + BQ - 1) / BQ : 0;
        int x = cnt;                                    // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += warp_tot[w];
        const int start = carry + wbase + x - cnt;
        for (int t = 0; t < cnt; ++t)
            if (start + t < max_tiles) { tile_video[start + t] = v; tile_q0[start + t] = t * BQ; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + cnt;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_tiles_out = min(carry, max_tiles);
}

}  // namespace

int launch_attn_schedule(const int32_t *cu_seqlens, int B, int32_t *tile_video, int32_t *tile_q0,
                         int32_t *n_tiles_out, int max_tiles, cudaStream_t s) {
    attn_schedule_kernel<<<1, 1024, 0, s>>>(cu_seqlens, B, tile_video, tile_q0, n_tiles_out, max_tiles);
    VSUM_LAUNCH_OK("attn_schedule_kernel");
    return VSUM_OK;
}

// lse2 != NULL selects the training variant: lse2 [T, 4] receives log2-domain log-sum-exps, P is dropped
// with probability drop_p (16-bit resolution), the output is scaled by 1 / (1 - drop_p) and `out` is an
// FP32 [T,256] buffer.
int launch_attention_tc05(const __nv_bfloat16 *qkv, const int32_t *cu_seqlens, const int32_t *tile_video,
                          const int32_t *tile_q0, const int32_t *n_tiles_ptr, int max_tiles, int64_t T,
                          float scale, void *out, cudaStream_t s, float *lse2, float drop_p,
                          unsigned long long seed) {
    if (T == 0 || max_tiles == 0) return VSUM_OK;
    CUtensorMap tm;
    int rc = make_tensor_map_2d(&tm, qkv, 2, 3 * DM, (uint64_t)T, (uint64_t)3 * DM * 2, 64, 128);
    if (rc) return rc;
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(attn_tc05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM)); VSUM_CUDA_OK(cudaFuncSetAttribute(attn_tc05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM)));
    dim3 grid((unsigned)max_tiles, NH);
    ProfScope prof(PROF_ATTN, s);
    const float sl2 = scale * 1.4426950408889634f;
    if (lse2) {
        const uint32_t thresh = attn_drop_thresh16(drop_p);
        attn_tc05_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, s>>>(tm, cu_seqlens, tile_video, tile_q0, n_tiles_ptr, (__nv_bfloat16 *)out, sl2,
                                                                   lse2, 65536.0f / (float)(65536u - thresh), thresh, seed);
    } else {
        attn_tc05_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, s>>>(tm, cu_seqlens, tile_video, tile_q0, n_tiles_ptr, (__nv_bfloat16 *)out, sl2,
                                                                    nullptr, 1.0f, 0u, 0ull);
    }
    VSUM_LAUNCH_OK("attn_tc05_kernel");
    return VSUM_OK;
}

}  // namespace vsum
