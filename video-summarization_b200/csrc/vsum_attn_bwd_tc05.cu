// Backward of the variable-length multi-head attention (vsum_attn_tc05.cu) on the sm_100a tensor cores:
// the autograd of src/model/simnet.py:155-161 (softmax(QK^T * d_model^-0.5) -> dropout -> PV) that
// `loss.backward()` runs in src/train.py:125 / src/pretrain.py:63.  No [N,N] tensor is materialised.
//
// One CTA = one (video, head, 128-key block j); it loops over the video's 128-query blocks i:
//   S  = Q_i K_j^T            A = Q_i  (K-major)        B = K_j  (K-major)      -> TMEM [128 x 128]
//   dP = dO_i V_j^T           A = dO_i (K-major)        B = V_j  (K-major)      -> TMEM [128 x 128]
//   P  = exp2(S c - lse2), Pd = dropout(P), dS = P o (dropout(dP) - delta)      (8 warps, thread <-> query row half)
//   dV_j += Pd^T dO_i         A = Pd   (M-major: keys)  B = dO_i (N-major)      -> TMEM [128 x 64], over all i
//   dK_j += dS^T Q_i          A = dS   (M-major: keys)  B = Q_i  (N-major)      -> TMEM [128 x 64], over all i
//   dQ_i  = dS K_j            A = dS   (K-major)        B = K_j  (N-major)      -> TMEM [128 x 64] -> red.add to HBM
// Every operand is consumed from the one 128B-swizzled tile TMA (or the softmax warps) wrote; the
// transposed uses are MN-major descriptors of the same bytes.  S/dP of block i+1 are issued as soon as
// the softmax warps hold block i in registers, so the tensor pipe works under the exponentials.
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {
namespace {

constexpr int HD = 64, DM = 256, NH = 4;
constexpr int BLK = 128;
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int BWD_THREADS = 384;        // 4 control warps + 8 compute warps
constexpr int BWD_TMEM_COLS = 512;      // S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448)
constexpr size_t BWD_SMEM = 10 * (size_t)TILE_BYTES + 256;   // K, V, Q x2, dO x2, Pd (2 halves), dS (2 halves), barriers

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// packed fp32x2 arithmetic (sm_100): half the FFMA / FMUL issue slots of the dS computation
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)), "l"(*reinterpret_cast<const uint64_t *>(&c)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t *>(&d))
        : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return d;
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const int32_t *__restrict__ cu, const int32_t *__restrict__ tile_video,
                     const int32_t *__restrict__ tile_k0, const int32_t *__restrict__ n_tiles_ptr,
                     const float *__restrict__ lse2, const float *__restrict__ delta, float *__restrict__ dqkv,
                     float scale_log2e, float scale, float keep_scale, uint32_t drop_thresh16, unsigned long long seed) {
    if ((int)blockIdx.x >= __ldg(n_tiles_ptr)) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t *sK = smem, *sV = smem + TILE_BYTES;
    uint8_t *sQ = smem + 2 * (size_t)TILE_BYTES;        // 2 stages
    uint8_t *sG = smem + 4 * (size_t)TILE_BYTES;        // dO, 2 stages
    uint8_t *sP = smem + 6 * (size_t)TILE_BYTES;        // Pd: two 64-key halves
    uint8_t *sD = smem + 8 * (size_t)TILE_BYTES;        // dS: two 64-key halves
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 10 * (size_t)TILE_BYTES);
    uint64_t *kv_full = bars, *q_full = bars + 1, *q_empty = bars + 3, *sdp_full = bars + 5, *sdp_empty = bars + 6,
             *pds_full = bars + 7, *done345 = bars + 8, *dq_empty = bars + 9;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vid = __ldg(tile_video + blockIdx.x), k0 = __ldg(tile_k0 + blockIdx.x);
    const int h_idx = blockIdx.y;
    const int base = __ldg(cu + vid), n = __ldg(cu + vid + 1) - base;
    const int nq = (n + BLK - 1) / BLK;

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmQKV); tc::tma_prefetch_desc(&tmDO); }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) { tc::mbar_init(q_full + s, 1); tc::mbar_init(q_empty + s, 1); }
        tc::mbar_init(sdp_full, 1); tc::mbar_init(sdp_empty, 256);
        tc::mbar_init(pds_full, 256); tc::mbar_init(done345, 1); tc::mbar_init(dq_empty, 256);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, BWD_TMEM_COLS); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

    if (warp < 4) {
        tc::setmaxnreg_dec<104>();
        if (warp == 0 && lane == 0) {   // ===== TMA producer =====
            tc::mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
            tc::tma_load_2d(sK, &tmQKV, kv_full, DM + h_idx * HD, base + k0);
            tc::tma_load_2d(sV, &tmQKV, kv_full, 2 * DM + h_idx * HD, base + k0);
            for (int i = 0; i < nq; ++i) {
                const int s = i & 1;
                tc::mbar_wait(q_empty + s, ((i >> 1) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(q_full + s, 2 * TILE_BYTES);
                tc::tma_load_2d(sQ + (size_t)s * TILE_BYTES, &tmQKV, q_full + s, h_idx * HD, base + i * BLK);
                tc::tma_load_2d(sG + (size_t)s * TILE_BYTES, &tmDO, q_full + s, h_idx * HD, base + i * BLK);
            }
        } else if (warp == 1) {   // ===== MMA issuer: warp-uniform control flow, one elected lane issues =====
            constexpr uint32_t IDESC_S = tc::make_idesc(1, BLK, BLK, 0, 0);    // [128 x 128], A and B K-major
            constexpr uint32_t IDESC_T = tc::make_idesc(1, BLK, HD, 1, 1);     // dV / dK: A M-major, B N-major
            constexpr uint32_t IDESC_Q = tc::make_idesc(1, BLK, HD, 0, 1);     // dQ: A K-major, B N-major
            constexpr uint32_t TILE16 = TILE_BYTES >> 4;
            const uint64_t k_km = tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024);      // K-major: +2 per 16 head-dim
            const uint64_t v_km = tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024);
            const uint64_t q_km = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);      // + stage * TILE16
            const uint64_t g_km = tc::make_smem_desc_sw128(tc::smem_u32(sG), 16, 1024);
            // N-major B tiles ([row][64 head-dim], one swizzle atom wide): +128 (2048 B) per 16 rows of K
            const uint64_t k_nm = k_km, q_nm = q_km, g_nm = g_km;
            // M-major A (keys contiguous): the two 64-key halves are one tile apart (LBO), +128 per 16 query rows
            const uint64_t p_mm = tc::make_smem_desc_sw128(tc::smem_u32(sP), TILE_BYTES, 1024);
            const uint64_t d_mm = tc::make_smem_desc_sw128(tc::smem_u32(sD), TILE_BYTES, 1024);
            const uint64_t d_km = tc::make_smem_desc_sw128(tc::smem_u32(sD), 16, 1024);      // K-major: half k/4, +2 per 16 keys
            auto issue_s_dp = [&](int i) {
                const uint64_t so = (uint64_t)((i & 1) * TILE16);
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        tc::mma_f16_ss(tS, q_km + so + (uint64_t)(k * 2), k_km + (uint64_t)(k * 2), IDESC_S, k != 0);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        tc::mma_f16_ss(tDP, g_km + so + (uint64_t)(k * 2), v_km + (uint64_t)(k * 2), IDESC_S, k != 0);
                    tc::mma_commit(sdp_full);
                }
                __syncwarp();
            };
            tc::mbar_wait(kv_full, 0);
            tc::mbar_wait(q_full, 0);
            tc::tc_fence_after();
            issue_s_dp(0);
            for (int i = 0; i < nq; ++i) {
                if (i + 1 < nq) {
                    tc::mbar_wait(q_full + ((i + 1) & 1), ((i + 1) >> 1) & 1);
                    tc::mbar_wait(sdp_empty, i & 1);
                    tc::tc_fence_after();
                    issue_s_dp(i + 1);
                }
                tc::mbar_wait(pds_full, i & 1);
                if (i > 0) tc::mbar_wait(dq_empty, (i - 1) & 1);
                tc::tc_fence_after();
                const uint64_t so = (uint64_t)((i & 1) * TILE16);
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLK / 16; ++k)     // contraction over the 128 queries, 16 per MMA
                        tc::mma_f16_ss(tDV, p_mm + (uint64_t)(k * 128), g_nm + so + (uint64_t)(k * 128), IDESC_T, (i | k) != 0);
#pragma unroll
                    for (int k = 0; k < BLK / 16; ++k)
                        tc::mma_f16_ss(tDK, d_mm + (uint64_t)(k * 128), q_nm + so + (uint64_t)(k * 128), IDESC_T, (i | k) != 0);
#pragma unroll
                    for (int k = 0; k < BLK / 16; ++k)     // contraction over the 128 keys
                        tc::mma_f16_ss(tDQ, d_km + (uint64_t)((k >> 2) * TILE16 + (k & 3) * 2), k_nm + (uint64_t)(k * 128), IDESC_Q, k != 0);
                    tc::mma_commit(q_empty + (i & 1));
                    tc::mma_commit(done345);
                }
                __syncwarp();
            }
        }
    } else {   // ===== compute warps: two threads per query row, 64 keys each =====
        tc::setmaxnreg_inc<200>();
        const int qd = warp & 3, hf = (warp - 4) >> 2;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t row_off = (uint32_t)hf * TILE_BYTES + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t p_row = tc::smem_u32(sP) + row_off, d_row = tc::smem_u32(sD) + row_off;
        const int valid_k = n - k0 - hf * 64;       // keys of my half inside the video
        float dq[32];

        auto flush_dq = [&](int i_prev) {           // dQ rows of query block i_prev += my 32 columns
            const int qr = i_prev * BLK + r;
            if (qr < n) {
                float *dst = dqkv + (int64_t)(base + qr) * (3 * DM) + h_idx * HD + hf * 32;
#pragma unroll
                for (int c = 0; c < 32; c += 4) red_add_v4(dst + c, dq[c] * scale, dq[c + 1] * scale, dq[c + 2] * scale, dq[c + 3] * scale);
            }
        };

#ifdef VSUM_BWD_TIMING
        long long tph[7] = {0, 0, 0, 0, 0, 0, 0}, tmark = clock64();
#define TMB(i) do { const long long _n = clock64(); tph[i] += _n - tmark; tmark = _n; } while (0)
#else
#define TMB(i) do { } while (0)
#endif
        for (int i = 0; i < nq; ++i) {
            const int qr = i * BLK + r;
            const bool vq = qr < n;
            const float lse_r = vq ? __ldg(lse2 + (int64_t)(base + qr) * NH + h_idx) : INFINITY;
            const float dl_r = vq ? __ldg(delta + (int64_t)(base + qr) * NH + h_idx) : 0.f;
            uint32_t s[64], g[64];
            tc::mbar_wait(sdp_full, i & 1);
            TMB(0);
            tc::tc_fence_after();
            {
                uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                uint32_t(&g0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&g[0]);
                uint32_t(&g1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&g[32]);
                tc::tmem_ld32(tS + lane_off + hf * 64, s0);
                tc::tmem_ld32(tS + lane_off + hf * 64 + 32, s1);
                tc::tmem_ld32(tDP + lane_off + hf * 64, g0);
                tc::tmem_ld32(tDP + lane_off + hf * 64 + 32, g1);
            }
            tc::tmem_wait_ld();
            tc::tc_fence_before();
            tc::mbar_arrive(sdp_empty);
            TMB(1);

            // Pd -> s[], dS -> g[]  (fp32 in place); the tail-masked and the dropout variants are separate
            // warp-uniform paths so the common full, p = 0 tile pays for neither
            if (valid_k < 64) {
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c >= valid_k) s[c] = 0xff800000u;         // exp2(-inf) = 0: key past the end of the video
            }
            const float2 c2 = make_float2(scale_log2e, scale_log2e), nl2 = make_float2(-lse_r, -lse_r), nd2 = make_float2(-dl_r, -dl_r);
            if (drop_thresh16 != 0) {
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const unsigned long long z = dropout_bits64(seed, attn_drop_group_index(base + qr, h_idx, NH, (k0 + hf * 64 + c) >> 2));
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {
                        const float2 x = ffma2(make_float2(__uint_as_float(s[c + e]), __uint_as_float(s[c + e + 1])), c2, nl2);
                        const float2 p = make_float2(ex2(x.x), ex2(x.y));
                        const float2 ks = make_float2((uint32_t)((z >> (16 * e)) & 0xffffu) >= drop_thresh16 ? keep_scale : 0.f,
                                                      (uint32_t)((z >> (16 * e + 16)) & 0xffffu) >= drop_thresh16 ? keep_scale : 0.f);
                        const float2 pd = fmul2(p, ks);
                        // dS = P (dropout(dP) - delta)
                        const float2 t = ffma2(make_float2(__uint_as_float(g[c + e]), __uint_as_float(g[c + e + 1])), ks, nd2);
                        const float2 ds = fmul2(p, t);
                        s[c + e] = __float_as_uint(pd.x); s[c + e + 1] = __float_as_uint(pd.y);
                        g[c + e] = __float_as_uint(ds.x); g[c + e + 1] = __float_as_uint(ds.y);
                    }
                }
            } else {
                const float2 one2 = make_float2(1.f, 1.f);
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    const float2 x = ffma2(make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), c2, nl2);
                    const float2 p = make_float2(ex2(x.x), ex2(x.y));
                    const float2 t = ffma2(make_float2(__uint_as_float(g[c]), __uint_as_float(g[c + 1])), one2, nd2);
                    const float2 ds = fmul2(p, t);
                    s[c] = __float_as_uint(p.x); s[c + 1] = __float_as_uint(p.y);
                    g[c] = __float_as_uint(ds.x); g[c + 1] = __float_as_uint(ds.y);
                }
            }
            TMB(2);
            if (i > 0) {   // MMAs of block i-1 done: Pd / dS buffers are free and dQ(i-1) is complete
                tc::mbar_wait(done345, (i - 1) & 1);
                TMB(3);
                tc::tc_fence_after();
                uint32_t(&t)[32] = *reinterpret_cast<uint32_t(*)[32]>(&dq[0]);
                tc::tmem_ld32(tDQ + lane_off + hf * 32, t);
                tc::tmem_wait_ld();
                tc::tc_fence_before();
                tc::mbar_arrive(dq_empty);
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const uint32_t off = (uint32_t)((ch ^ (r & 7)) << 4);
                const int c = ch * 8;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + off),
                             "r"(pack2(__uint_as_float(s[c]), __uint_as_float(s[c + 1]))), "r"(pack2(__uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]))),
                             "r"(pack2(__uint_as_float(s[c + 4]), __uint_as_float(s[c + 5]))), "r"(pack2(__uint_as_float(s[c + 6]), __uint_as_float(s[c + 7])))
                             : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d_row + off),
                             "r"(pack2(__uint_as_float(g[c]), __uint_as_float(g[c + 1]))), "r"(pack2(__uint_as_float(g[c + 2]), __uint_as_float(g[c + 3]))),
                             "r"(pack2(__uint_as_float(g[c + 4]), __uint_as_float(g[c + 5]))), "r"(pack2(__uint_as_float(g[c + 6]), __uint_as_float(g[c + 7])))
                             : "memory");
            }
            TMB(4);
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            tc::mbar_arrive(pds_full);
            TMB(5);
            if (i > 0) flush_dq(i - 1);
            TMB(6);
        }
#ifdef VSUM_BWD_TIMING
        if (blockIdx.x == 70 && blockIdx.y == 1 && lane == 0 && (warp == 4 || warp == 9))
            printf("bwd warp %2d nq %d | wait S,dP %lld | tmem ld %lld | exps, dS %lld | wait MMA345(i-1) %lld | dQ ld + P,dS stores %lld | fence+arrive %lld | dQ atomics %lld (clk per query block)\n",
                   warp, nq, tph[0] / nq, tph[1] / nq, tph[2] / nq, tph[3] / nq, tph[4] / nq, tph[5] / nq, tph[6] / nq);
#endif
        // last dQ block, then dK / dV of my key row
        tc::mbar_wait(done345, (nq - 1) & 1);
        tc::tc_fence_after();
        {
            uint32_t(&t)[32] = *reinterpret_cast<uint32_t(*)[32]>(&dq[0]);
            tc::tmem_ld32(tDQ + lane_off + hf * 32, t);
            tc::tmem_wait_ld();
        }
        flush_dq(nq - 1);
        uint32_t kv[32];
        const bool vk = k0 + r < n;
        float *row = dqkv + (int64_t)(base + k0 + r) * (3 * DM) + h_idx * HD + hf * 32;
        tc::tmem_ld32(tDK + lane_off + hf * 32, kv);
        tc::tmem_wait_ld();
        if (vk) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
                *reinterpret_cast<float4 *>(row + DM + c) = make_float4(__uint_as_float(kv[c]) * scale, __uint_as_float(kv[c + 1]) * scale,
                                                                        __uint_as_float(kv[c + 2]) * scale, __uint_as_float(kv[c + 3]) * scale);
        }
        tc::tmem_ld32(tDV + lane_off + hf * 32, kv);
        tc::tmem_wait_ld();
        if (vk) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
                *reinterpret_cast<float4 *>(row + 2 * DM + c) = make_float4(__uint_as_float(kv[c]), __uint_as_float(kv[c + 1]),
                                                                            __uint_as_float(kv[c + 2]), __uint_as_float(kv[c + 3]));
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, BWD_TMEM_COLS); }
}

// One warp per (frame, head): lane covers 2 of the 64 head-dim columns.
__global__ void __launch_bounds__(256)
attn_delta_bf16_kernel(const float *__restrict__ o, const __nv_bfloat16 *__restrict__ d_o, float *__restrict__ delta, int64_t T) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= T * NH) return;
    const int64_t off = w * HD + 2 * lane;           // [T,256] row-major == [(t,h), 64]
    const float2 a = *reinterpret_cast<const float2 *>(o + off);
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(d_o + off));
    float acc = fmaf(a.x, b.x, a.y * b.y);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) delta[w] = acc;
}

}  // namespace

int launch_attn_delta_bf16(const float *o, const __nv_bfloat16 *dO16, float *delta, int64_t T, cudaStream_t s) {
    if (T == 0) return VSUM_OK;
    attn_delta_bf16_kernel<<<(unsigned)ceil_div(T * NH * 32, 256), 256, 0, s>>>(o, dO16, delta, T);
    VSUM_LAUNCH_OK("attn_delta_bf16_kernel");
    return VSUM_OK;
}

int launch_attention_bwd_tc05(const __nv_bfloat16 *qkv16, const __nv_bfloat16 *dO16, const float *lse2, const float *delta,
                              const int32_t *cu_seqlens, const int32_t *tile_video, const int32_t *tile_k0,
                              const int32_t *n_tiles_ptr, int max_tiles, int64_t T, float scale, float drop_p,
                              unsigned long long seed, float *dqkv, cudaStream_t s) {
    if (T == 0 || max_tiles == 0) return VSUM_OK;
    CUtensorMap tmQKV, tmDO;
    int rc = make_tensor_map_2d(&tmQKV, qkv16, 2, 3 * DM, (uint64_t)T, (uint64_t)3 * DM * 2, 64, 128);
    if (rc) return rc;
    if ((rc = make_tensor_map_2d(&tmDO, dO16, 2, DM, (uint64_t)T, (uint64_t)DM * 2, 64, 128))) return rc;
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM)));
    // dQ is accumulated with red.add from every key block; dK / dV rows are each written by one CTA
    VSUM_CUDA_OK(cudaMemsetAsync(dqkv, 0, (size_t)T * 3 * DM * sizeof(float), s));
    const uint32_t thresh = attn_drop_thresh16(drop_p);
    dim3 grid((unsigned)max_tiles, NH);
    ProfScope prof(PROF_OTHER, s);
    attn_bwd_tc05_kernel<<<grid, BWD_THREADS, BWD_SMEM, s>>>(tmQKV, tmDO, cu_seqlens, tile_video, tile_k0, n_tiles_ptr, lse2, delta,
                                                            dqkv, scale * 1.4426950408889634f, scale,
                                                            65536.0f / (float)(65536u - thresh), thresh, seed);
    VSUM_LAUNCH_OK("attn_bwd_tc05_kernel");
    return VSUM_OK;
}

}  // namespace vsum
