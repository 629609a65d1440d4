// Fused feed-forward block of the scorer on the 5th-gen tensor cores (sm_100a):
//     out = LayerNorm(relu(X W1^T + b1) W2^T + b2 + X) * gamma + beta      (+ regression head and sigmoid on the last layer)
// Replaces src/model/simnet.py:180-183 (MLP: fc1 -> ReLU -> fc2) and 109-110 (residual + norm2) -- and, for the last
// encoder block, final_layer (simnet.py:42) with the callers' sigmoid (train.py:144) -- in ONE kernel: the [T,1024]
// hidden activation never leaves the SM (the two-kernel form writes and re-reads 2 KB per frame and layer through HBM).
//
// Persistent, one CTA per SM, one 128-row tile of X at a time.  The hidden dimension is walked in eight chunks of 128:
//     G1(j): H_j[128x128] = X[128x256] W1_j^T            (tcgen05.mma M128 N128, 16 K-steps, fp32 in tensor memory)
//     E(j) : eight epilogue warps: tcgen05.ld H_j, + b1, ReLU, -> bf16, written to shared memory as a K-major SW128 operand
//     G2(j): Y[128x256] += H_j[128x128] W2_j^T            (M128 N256, 8 K-steps, fp32 accumulator for the whole tile)
// issued as G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ... so that the conversion of chunk j runs under G1(j+1): H is double
// buffered in tensor memory (2 x 128 columns next to the 256 columns of Y = all 512) and single buffered in shared memory.
// After G2(7) the same warps run the residual + LayerNorm epilogue of vsum_gemm_tc05.cu (two threads per row, pre-norm row
// parked in tensor memory, residual in / output out through TMA and a staging tile per column half).
// Shared memory (227 KB): X tile 64 KB (4 k-blocks), weight ring 3 x 32 KB (half a W1 chunk or half a W2 chunk per stage,
// in MMA issue order), H 32 KB, staging 2 x 16 KB.  Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 epilogue.
// Weights stream from L2 (1 MB per tile, ~40 B/clk/SM at the sustained tensor rate): ncu shows the L2 at 33 % of its peak and
// the tensor pipe at 51 % (= 80 % of the sustained bf16 peak), so CTA pairs sharing weight stages by multicast were not built.
// Tiles come from a dynamic scheduler (global counter -> shared-memory ring), see vsum_gemm_tc05.cu.
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {
namespace {

constexpr int BM = 128, DM = 256, DFF = 1024, CH = 128;       // rows per tile, d_model, hidden width, hidden chunk
constexpr int NCH = DFF / CH;                                 // 8 chunks
constexpr int KB16 = BM * 128;                                // 16 KB: [128 rows x 64 bf16] SW128 block
constexpr int STAGE = 2 * KB16;                               // 32 KB ring stage
constexpr int NST = 3;
constexpr int FFN_THREADS = 384;
constexpr size_t OFF_X = 0, OFF_RING = 4 * (size_t)KB16, OFF_H = OFF_RING + (size_t)NST * STAGE, OFF_STG = OFF_H + 2 * (size_t)KB16,
                 OFF_BARS = OFF_STG + 2 * (size_t)KB16, FFN_SMEM = OFF_BARS + 384 + 2560;
static_assert(FFN_SMEM <= 232448, "fused FFN: shared memory budget");
constexpr int TM_Y = 0, TM_H = 256;                            // tensor-memory columns
constexpr int EPI_BAR = 1;

struct FfnParams {
    int64_t M, m_tiles;
    const float *b1, *b2, *gamma, *beta, *head_w, *head_b;
    float *scores_out, *feats_out;
    int apply_sigmoid, store_out, head;
    int *sched;               // dynamic tile scheduler: sched[0] = next tile, sched[8] = CTAs that have left
};
constexpr int SCHED_RING = 8;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(FFN_THREADS, 1)
ffn_tc05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const FfnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t *sX = smem + OFF_X, *ring = smem + OFF_RING, *sH = smem + OFF_H, *stg = smem + OFF_STG;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint64_t *full = bars, *empty = bars + NST, *x_full = bars + 2 * NST, *x_free = x_full + 1;
    uint64_t *h_tfull = x_free + 1, *h_tfree = h_tfull + 2, *h_sfull = h_tfree + 2, *h_sfree = h_sfull + 1;
    uint64_t *y_full = h_sfree + 1, *y_free = y_full + 1, *rfull = y_free + 1;                 // rfull[2]: residual staging per half
    uint64_t *sched_full = rfull + 2, *sched_empty = sched_full + SCHED_RING;
    int32_t *sched_tile = reinterpret_cast<int32_t *>(sched_empty + SCHED_RING);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sched_tile + SCHED_RING);
    float *xch = reinterpret_cast<float *>(bars) + 96;                                            // [2][128][2] sums + [128] head dots

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmX); tc::tma_prefetch_desc(&tmW1); tc::tma_prefetch_desc(&tmW2); tc::tma_prefetch_desc(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NST; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        tc::mbar_init(x_full, 1); tc::mbar_init(x_free, 1);
        for (int b = 0; b < 2; ++b) { tc::mbar_init(h_tfull + b, 1); tc::mbar_init(h_tfree + b, 256); tc::mbar_init(rfull + b, 1); }
        tc::mbar_init(h_sfull, 256); tc::mbar_init(h_sfree, 1);
        tc::mbar_init(y_full, 1); tc::mbar_init(y_free, 256);
        for (int a = 0; a < SCHED_RING; ++a) { tc::mbar_init(sched_full + a, 1); tc::mbar_init(sched_empty + a, 9); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Dynamic tile scheduler (as in vsum_gemm_tc05.cu): the producer takes tiles from a global counter and hands them to
    // the MMA and epilogue warps through a shared-memory ring; -1 ends the walk.
    auto next_tile = [&](uint32_t n) -> int64_t {
        const int slot = n % SCHED_RING;
        tc::mbar_wait(sched_full + slot, (n / SCHED_RING) & 1);
        const int64_t t = sched_tile[slot];
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(sched_empty + slot);
        return t;
    };
    auto fetch_tile = [&](uint32_t n) -> int64_t {
        const int slot = n % SCHED_RING;
        tc::mbar_wait(sched_empty + slot, ((n / SCHED_RING) & 1) ^ 1);
        const int64_t t = n == 0 ? (int64_t)blockIdx.x : (int64_t)gridDim.x + atomicAdd(p.sched, 1);   // first tile: the static one, no atomic round trip
        sched_tile[slot] = t < p.m_tiles ? (int32_t)t : -1;
        tc::mbar_arrive(sched_full + slot);
        return t < p.m_tiles ? t : -1;
    };

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer: X tile + weight ring in MMA issue order =====
            uint32_t it = 0, tl = 0, nf = 0;
            bool x_issued = false;                        // X of the tile about to start is already on its way
            int64_t t_next = fetch_tile(nf++);
            auto load_x = [&](int64_t t) {
                tc::mbar_arrive_expect_tx(x_full, 4 * KB16);
                for (int kb = 0; kb < 4; ++kb) tc::tma_load_2d(sX + (size_t)kb * KB16, &tmX, x_full, kb * 64, (int)(t * BM));
            };
            for (; t_next >= 0; ++tl) {
                const int64_t t = t_next;
                if (!x_issued) { tc::mbar_wait(x_free, (tl & 1) ^ 1); load_x(t); }
                x_issued = false;
                t_next = fetch_tile(nf++);                // one tile ahead (the early X load below needs it), after this tile's X is on its way
                const bool has_next = t_next >= 0;
                auto stage_w1 = [&](int j, int hf) {      // k-blocks 2 hf, 2 hf + 1 of W1 chunk j
                    const int s = it % NST;
                    tc::mbar_wait(empty + s, ((it / NST) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(full + s, STAGE);
                    tc::tma_load_2d(ring + (size_t)s * STAGE, &tmW1, full + s, (2 * hf) * 64, j * CH);
                    tc::tma_load_2d(ring + (size_t)s * STAGE + KB16, &tmW1, full + s, (2 * hf + 1) * 64, j * CH);
                    ++it;
                };
                auto stage_w2 = [&](int j, int kb) {      // [256 x 64] k-block kb of W2's columns [128 j, 128 j + 128)
                    const int s = it % NST;
                    tc::mbar_wait(empty + s, ((it / NST) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(full + s, STAGE);
                    tc::tma_load_2d(ring + (size_t)s * STAGE, &tmW2, full + s, j * CH + kb * 64, 0);
                    ++it;
                    // the next tile's X as soon as this tile's last G1 has released the buffer (not only after the ring drained)
                    if (has_next && !x_issued && tc::mbar_test(x_free, tl & 1)) { load_x(t_next); x_issued = true; }
                };
                stage_w1(0, 0); stage_w1(0, 1); stage_w1(1, 0); stage_w1(1, 1);
                for (int j = 0; j < NCH; ++j) {
                    stage_w2(j, 0); stage_w2(j, 1);
                    if (j + 2 < NCH) { stage_w1(j + 2, 0); stage_w1(j + 2, 1); }
                }
            }
        }
    } else if (warp == 1) {   // ===== MMA issuer (whole warp, warp-uniform control flow, one elected lane issues) =====
        constexpr uint32_t IDESC1 = tc::make_idesc(1, BM, CH, 0, 0), IDESC2 = tc::make_idesc(1, BM, DM, 0, 0);
        const uint64_t desc0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
        const uint32_t x_off = (uint32_t)OFF_X >> 4, ring_off = (uint32_t)OFF_RING >> 4, h_off = (uint32_t)OFF_H >> 4;
        uint32_t it = 0, tl = 0, c1 = 0, c2 = 0;           // ring position, tile count, G1 chunks issued, G2 chunks issued
        for (;; ++tl) {
            if (next_tile(tl) < 0) break;
            auto g1 = [&](int j) {
                const uint32_t b = c1 & 1;
                tc::mbar_wait(h_tfree + b, ((c1 >> 1) & 1) ^ 1);      // the epilogue has drained the chunk that used this TMEM buffer
                tc::tc_fence_after();
                const uint32_t d = tmem_base + TM_H + b * CH;
                for (int hf = 0; hf < 2; ++hf, ++it) {
                    const int s = it % NST;
                    tc::mbar_wait(full + s, (it / NST) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc::mma_f16_ss(d, desc0 + (uint64_t)(x_off + (uint32_t)(2 * hf + kk) * (KB16 >> 4) + k * 2),
                                               desc0 + (uint64_t)(ring_off + (uint32_t)s * (STAGE >> 4) + (uint32_t)kk * (KB16 >> 4) + k * 2),
                                               IDESC1, (hf | kk | k) != 0);
                        tc::mma_commit(empty + s);
                        if (hf == 1) {
                            tc::mma_commit(h_tfull + b);
                            if (j == NCH - 1) tc::mma_commit(x_free);     // the tile's last product that reads X
                        }
                    }
                    __syncwarp();
                }
                ++c1;
            };
            auto g2 = [&](int j) {
                tc::mbar_wait(h_sfull, c2 & 1);                        // H_j sits in shared memory
                if (j == 0) tc::mbar_wait(y_free, (tl & 1) ^ 1);       // the previous tile's LayerNorm epilogue has left Y
                tc::tc_fence_after();
                for (int kb = 0; kb < 2; ++kb, ++it) {
                    const int s = it % NST;
                    tc::mbar_wait(full + s, (it / NST) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc::mma_f16_ss(tmem_base + TM_Y, desc0 + (uint64_t)(h_off + (uint32_t)kb * (KB16 >> 4) + k * 2),
                                           desc0 + (uint64_t)(ring_off + (uint32_t)s * (STAGE >> 4) + k * 2), IDESC2, (j | kb | k) != 0);
                        tc::mma_commit(empty + s);
                        if (kb == 1) {
                            tc::mma_commit(h_sfree);
                            if (j == NCH - 1) tc::mma_commit(y_full);
                        }
                    }
                    __syncwarp();
                }
                ++c2;
            };
            tc::mbar_wait(x_full, tl & 1);
            tc::tc_fence_after();
            g1(0); g1(1);
            for (int j = 0; j < NCH; ++j) {
                g2(j);
                if (j + 2 < NCH) g1(j + 2);
            }
        }
    } else if (warp >= 4) {   // ===== epilogue warps: thread <-> (row, column half) =====
        const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const bool leader = r == 0;
        const int epi_bar = EPI_BAR + half;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t h_row = tc::smem_u32(sH) + (uint32_t)half * KB16 + row_off;
        const uint32_t stg_row = tc::smem_u32(stg) + (uint32_t)half * KB16 + row_off;
        uint8_t *const stg_half = stg + (size_t)half * KB16;
        uint32_t sw_off[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sw_off[ch] = (uint32_t)((ch ^ (r & 7)) << 4);
        uint32_t tl = 0, hc = 0, res_it = 0;
        uint32_t ra[32], rb[32];
        for (;; ++tl) {
            const int64_t t = next_tile(tl);
            if (t < 0) break;
            const int64_t row = t * BM + r;
            const bool valid = row < p.M;
            if (leader) {   // this half's first residual chunk on its way while the tile's products run
                tc::bulk_wait_read<0>();
                tc::mbar_arrive_expect_tx(rfull + half, KB16);
                tc::tma_load_2d(stg_half, &tmX, rfull + half, (2 * half) * 64, (int)(t * BM));
            }
            // ---- hidden chunks: TMEM -> + b1 -> ReLU -> bf16 -> shared memory (A operand of G2)
#pragma unroll 1
            for (int j = 0; j < NCH; ++j, ++hc) {
                const uint32_t b = hc & 1;
                tc::mbar_wait(h_tfull + b, (hc >> 1) & 1);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + lane_off + TM_H + b * CH + half * 64;
                tc::tmem_ld32(taddr, ra);
                tc::tmem_ld32(taddr + 32, rb);
                tc::tmem_wait_ld();
                tc::tc_fence_before();
                tc::mbar_arrive(h_tfree + b);
                uint32_t pk[32];
                const float *bp = p.b1 + j * CH + half * 64;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint32_t *src = ch < 4 ? &ra[ch * 8] : &rb[(ch - 4) * 8];
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bp + ch * 8)), b1v = __ldg(reinterpret_cast<const float4 *>(bp + ch * 8 + 4));
                    pk[ch * 4 + 0] = pack_bf16(fmaxf(__uint_as_float(src[0]) + b0.x, 0.f), fmaxf(__uint_as_float(src[1]) + b0.y, 0.f));
                    pk[ch * 4 + 1] = pack_bf16(fmaxf(__uint_as_float(src[2]) + b0.z, 0.f), fmaxf(__uint_as_float(src[3]) + b0.w, 0.f));
                    pk[ch * 4 + 2] = pack_bf16(fmaxf(__uint_as_float(src[4]) + b1v.x, 0.f), fmaxf(__uint_as_float(src[5]) + b1v.y, 0.f));
                    pk[ch * 4 + 3] = pack_bf16(fmaxf(__uint_as_float(src[6]) + b1v.z, 0.f), fmaxf(__uint_as_float(src[7]) + b1v.w, 0.f));
                }
                tc::mbar_wait(h_sfree, (hc & 1) ^ 1);                   // G2 of the previous chunk has read the shared-memory H tile
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) sts128(h_row + sw_off[ch], pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(h_sfull);
            }
            // ---- Y + b2 + residual -> LayerNorm (-> head): two threads per row, columns [128 half, 128 half + 128)
            tc::mbar_wait(y_full, tl & 1);
            tc::tc_fence_after();
            const uint32_t tY = tmem_base + lane_off + TM_Y;
            float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                const int cc = 2 * half + i;
                tc::tmem_ld32(tY + cc * 64, ra);
                tc::tmem_ld32(tY + cc * 64 + 32, rb);
                tc::mbar_wait(rfull + half, res_it & 1);
                ++res_it;
                uint4 rs[8];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) rs[ch] = lds128(stg_row + sw_off[ch]);
                tc::tmem_wait_ld();
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t *a = ch < 4 ? &ra[ch * 8] : &rb[(ch - 4) * 8];
                    const float *bp = p.b2 + cc * 64 + ch * 8;
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bp)), b1v = __ldg(reinterpret_cast<const float4 *>(bp + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
                    const uint32_t rw[4] = {rs[ch].x, rs[ch].y, rs[ch].z, rs[ch].w};
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162 *>(&rw[e >> 1]);
                        const float v0 = __uint_as_float(a[e]) + bb[e] + __low2float(h2);
                        const float v1 = __uint_as_float(a[e + 1]) + bb[e + 1] + __high2float(h2);
                        sum += v0 + v1;
                        sumsq = fmaf(v0, v0, fmaf(v1, v1, sumsq));
                        a[e] = __float_as_uint(v0);
                        a[e + 1] = __float_as_uint(v1);
                    }
                }
                tc::tmem_st32(tY + cc * 64, ra);          // park the pre-norm row in tensor memory
                tc::tmem_st32(tY + cc * 64 + 32, rb);
                tc::bar_sync(epi_bar, 128);               // my half has read this residual chunk
                if (leader && i == 0) {
                    tc::mbar_arrive_expect_tx(rfull + half, KB16);
                    tc::tma_load_2d(stg_half, &tmX, rfull + half, (cc + 1) * 64, (int)(t * BM));
                }
            }
            tc::tmem_wait_st();
            xch[(half * 128 + r) * 2] = sum;
            xch[(half * 128 + r) * 2 + 1] = sumsq;
            tc::bar_sync(EPI_BAR + 2, 256);
            sum += xch[((half ^ 1) * 128 + r) * 2];
            sumsq += xch[((half ^ 1) * 128 + r) * 2 + 1];
            tc::bar_sync(EPI_BAR + 2, 256);
            const float mean = sum * (1.0f / DM);
            const float var = fmaxf(sumsq * (1.0f / DM) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + 1e-5f);
            float dot = 0.f;
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                const int cc = 2 * half + i;
                tc::tmem_ld32(tY + cc * 64, ra);
                tc::tmem_ld32(tY + cc * 64 + 32, rb);
                tc::tmem_wait_ld();
                if (i == 1) { tc::tc_fence_before(); tc::mbar_arrive(y_free); }
                if (p.store_out) {
                    if (leader) tc::bulk_wait_read<0>();  // my previous store is done reading the staging tile
                    tc::bar_sync(epi_bar, 128);
                }
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint32_t *a = ch < 4 ? &ra[ch * 8] : &rb[(ch - 4) * 8];
                    const int c0 = cc * 64 + ch * 8;
                    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(p.gamma + c0)), g1 = __ldg(reinterpret_cast<const float4 *>(p.gamma + c0 + 4));
                    const float4 e0 = __ldg(reinterpret_cast<const float4 *>(p.beta + c0)), e1 = __ldg(reinterpret_cast<const float4 *>(p.beta + c0 + 4));
                    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = (__uint_as_float(a[e]) - mean) * rstd * gg[e] + be[e];
                    if (p.head) {
                        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(p.head_w + c0)), w1 = __ldg(reinterpret_cast<const float4 *>(p.head_w + c0 + 4));
                        dot = fmaf(y[0], w0.x, fmaf(y[1], w0.y, fmaf(y[2], w0.z, fmaf(y[3], w0.w, dot))));
                        dot = fmaf(y[4], w1.x, fmaf(y[5], w1.y, fmaf(y[6], w1.z, fmaf(y[7], w1.w, dot))));
                        if (p.feats_out && valid) {
                            float4 *fo = reinterpret_cast<float4 *>(p.feats_out + row * DM + c0);
                            fo[0] = make_float4(y[0], y[1], y[2], y[3]);
                            fo[1] = make_float4(y[4], y[5], y[6], y[7]);
                        }
                    }
                    if (p.store_out)
                        sts128(stg_row + sw_off[ch], pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                }
                if (p.store_out) {
                    tc::fence_proxy_async_smem();
                    tc::bar_sync(epi_bar, 128);
                    if (leader) {
                        tc::tma_store_2d(stg_half, &tmOut, cc * 64, (int)(t * BM));
                        tc::bulk_commit();
                    }
                }
            }
            if (p.head) {                                  // the two half-row dot products meet in half 0
                if (half == 1) xch[512 + r] = dot;
                tc::bar_sync(EPI_BAR + 2, 256);
                if (half == 0 && valid) {
                    float sc = dot + xch[512 + r] + __ldg(p.head_b);
                    if (p.apply_sigmoid) sc = 1.0f / (1.0f + __expf(-sc));
                    p.scores_out[row] = sc;
                }
            }
        }
        if (leader) tc::bulk_wait_all<0>();
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
    if (threadIdx.x == 0 && atomicAdd(p.sched + 8, 1) == (int)gridDim.x - 1) {   // the last CTA hands the scheduler slot back zeroed
        p.sched[0] = 0;
        p.sched[8] = 0;
        __threadfence();
    }
}

}  // namespace

int launch_ffn_tc05(const Tc05FfnArgs &a, cudaStream_t s) {
    if (a.M == 0) return VSUM_OK;
    VSUM_REQUIRE(a.x && a.w1 && a.w2 && a.b1 && a.b2 && a.gamma && a.beta, VSUM_EINVAL, "ffn_tc05: null pointer");
    VSUM_REQUIRE(a.out || (a.head_w && a.head_b && a.scores_out), VSUM_EINVAL, "ffn_tc05: needs an output (rows or scores)");
    VSUM_REQUIRE(a.M < ((int64_t)1 << 31), VSUM_EUNSUPPORTED, "ffn_tc05: M=%lld exceeds the TMA coordinate range", (long long)a.M);
    CUtensorMap tmX, tmW1, tmW2, tmOut;
    int rc = make_tensor_map_2d(&tmX, a.x, 2, DM, (uint64_t)a.M, DM * 2, 64, BM);
    if (rc) return rc;
    if ((rc = make_tensor_map_2d(&tmW1, a.w1, 2, DM, DFF, DM * 2, 64, CH))) return rc;
    if ((rc = make_tensor_map_2d(&tmW2, a.w2, 2, DFF, DM, DFF * 2, 64, DM))) return rc;
    tmOut = tmX;
    if (a.out && (rc = make_tensor_map_2d(&tmOut, a.out, 2, DM, (uint64_t)a.M, DM * 2, 64, BM))) return rc;
    FfnParams p{};
    p.M = a.M; p.m_tiles = ceil_div(a.M, BM);
    p.b1 = a.b1; p.b2 = a.b2; p.gamma = a.gamma; p.beta = a.beta; p.head_w = a.head_w; p.head_b = a.head_b;
    p.scores_out = a.scores_out; p.feats_out = a.feats_out; p.apply_sigmoid = a.apply_sigmoid;
    p.store_out = a.out != nullptr; p.head = a.head_w != nullptr && a.scores_out != nullptr;
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(ffn_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFN_SMEM)));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms = max(8, sms - scorer_sm_reserve());
    const unsigned grid = (unsigned)min(p.m_tiles, (int64_t)sms);
    p.sched = sched_slot();
    VSUM_REQUIRE(p.sched != nullptr, VSUM_ENOMEM, "ffn_tc05: no device memory for the tile scheduler");
    ProfScope prof(PROF_FFN, s);
    ffn_tc05_kernel<<<grid, FFN_THREADS, FFN_SMEM, s>>>(tmX, tmW1, tmW2, tmOut, p);
    VSUM_LAUNCH_OK("ffn_tc05_kernel");
    return VSUM_OK;
}

}  // namespace vsum
