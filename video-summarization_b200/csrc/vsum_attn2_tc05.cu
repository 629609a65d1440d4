// Variable-length (packed, padding-free) multi-head self-attention forward for the scorer on sm_100a, second
// generation.  Replaces src/model/simnet.py:155-161 (QK^T * d_model^-0.5 -> softmax -> PV; dropout on P at
// simnet.py:159 in the TRAIN variant) for d_model 256, 4 heads of 64.
//
// PERSISTENT kernel, one CTA per SM, work items = (video, head, 256 queries) handed out longest-video-first by an
// atomic counter.  A CTA works on TWO 128-query tiles (A, B) of its item at once, each tile an independent pipeline:
//   warp 0      work scheduler (atomicAdd -> shared-memory ring) and TMA producer: Q tiles of the item (double-buffered
//               across items), K / V tiles through a shared ring
//   warp 1 / 2  tcgen05.mma issuer of tile A / tile B (warp 2 also owns the TMEM allocation)
//   warp 3      idle: it only hands its registers to the softmax warps (setmaxnreg moves registers inside the CTA's own
//               allocation, 640 threads x 96 at launch)
//   warps 4-19  softmax: TWO threads per query row (64 key columns each), i.e. four warps on every SM sub-partition --
//               with one thread per row (two warps per sub-partition) the warps ran at one instruction per 4 clocks and
//               nothing hid their tensor-memory loads and barrier waits (profiles/r02_attn2_one_thread_per_row.txt)
// Tensor memory (512 columns): S_A, S_B fp32 [128 x 128] (0..255), O_A, O_B fp32 [128 x 64] (256..383),
// P_A, P_B bf16 [128 x 128] as packed pairs (384..511).  P is written with tcgen05.st and consumed as the TMEM A operand
// of the PV MMA, so the probabilities never touch shared memory, every K / V tile is fetched once per 256 queries, and
// the shared-memory port carries 64 KB per 128 x 128 tile instead of 144 KB (profiles/r01_microbench_mma_rate.txt).
//
// Exponent reference without a per-tile exchange.  The two threads of a row share ONE accumulator O and must therefore
// exponentiate against the same reference m_ref, but they never wait for each other inside a tile: every thread
// publishes the maximum of its 64 columns of tile j to shared memory BEFORE it releases S (s_empty), so when either of
// them gets the next S (s_full of tile j+1, which the MMA warp only issues after BOTH released tile j) both maxima of
// tile j are visible, and both derive the same decision from the same two numbers: m_ref (0 at the start of an item)
// moves to the row maximum when that maximum has left the window [m_ref - 24, m_ref + 24] (log2 units), O and the row
// sums being rescaled by the owner of each half.  Softmax is shift-invariant and fp32 / bf16 keep their relative
// precision, so a reference that lags one tile behind gives the same result as the exact running maximum -- as long as
// no exponential overflows or the whole first tile underflows, i.e. while a tile's scores stay within +-96 log2 units
// (+-66 nats) of the reference.  A row that sees more than that raises its item's flag, and a second launch of the same
// kernel (SAFE = true: exact maximum per tile, the two threads of a row exchange through a named barrier) recomputes
// the flagged items; with no flag raised that launch ends after reading one flag per item.
// A share of the exponentials runs as a degree-3 polynomial on the FMA pipe (the MUFU pipe, 16 ex2 / clk / SM, is the
// limiter at head_dim 64: profiles/r01_microbench_mufu_ex2.txt).
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {
namespace {

constexpr int HD = 64, DM = 256, NH = 4;
constexpr int BQ2 = 256, BKV = 128;
constexpr int TILE_BYTES = 128 * 128;            // 128 rows x 64 bf16
constexpr int A2_THREADS = 640;                  // 3 control warps + 1 register donor + 16 softmax warps; 640 x 96 registers at launch
constexpr int A2_REGS_CONTROL = 64, A2_REGS_DONOR = 24, A2_REGS_SOFTMAX = 104;   // setmaxnreg: 96 x (96 - 64) + 32 x (96 - 24) released >= 512 x (104 - 96) acquired
constexpr int A2_TMEM_COLS = 512;
#ifndef VSUM_A2_STAGES
#define VSUM_A2_STAGES 4
#endif
constexpr int KV_STAGES = VSUM_A2_STAGES;
constexpr int SCHED_RING = 4;
constexpr size_t A2_SMEM = (4 + 2 * (size_t)KV_STAGES) * TILE_BYTES + 512 + 6 * 1024;   // Q (2 items x 2 tiles), K ring, V ring, barriers, row exchange
#ifndef VSUM_A2_POLY_PERIOD
#define VSUM_A2_POLY_PERIOD 4        // every k-th pair of exponentials is a polynomial on the FMA pipe (0 = none)
#endif
#ifndef VSUM_A2_WINDOW
#define VSUM_A2_WINDOW 24.0f         // the exponent reference moves when a tile maximum leaves [m_ref - W, m_ref + W]
#endif
#define VSUM_A2_DANGER 96.0f         // scores this far from the reference could overflow / underflow: exact pass for the item

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
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
// 2^x on the FMA pipe: round-to-nearest split x = n + f (1.5 * 2^23 trick), degree-3 minimax polynomial for 2^f on
// [-0.5, 0.5] (max relative error 7.5e-5, 50x below the bf16 rounding of P), n added into the exponent field.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
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

// tcgen05.ld of 32 columns WITHOUT the wait, and a wait that names the destination registers so that no use of them can
// be scheduled above it (the plain wait is only ordered against memory).
__device__ __forceinline__ void tmem_wait_ld_on(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
        : "memory");
}

#ifdef VSUM_A2_TIMING      // clock64 phase stamps (timing experiments only): per-tile averages printed by CTA 0 after its first long items
#define A2_TDECL(n) long long tph[n] = {}, tmark = clock64(); int tcount = 0
#define A2_TMARK(i) do { const long long _n = clock64(); tph[i] += _n - tmark; tmark = _n; } while (0)
#else
#define A2_TDECL(n) do { } while (0)
#define A2_TMARK(i) do { } while (0)
#endif

struct ItemInfo {
    int base, n, q0, head, nkv;
    bool has_b;
};
__device__ __forceinline__ ItemInfo decode_item(int idx, const int32_t *__restrict__ cu, const int32_t *__restrict__ item_video,
                                                const int32_t *__restrict__ item_q0) {
    ItemInfo it;
    const int blk = idx >> 2;
    it.head = idx & 3;
    const int vid = __ldg(item_video + blk);
    it.q0 = __ldg(item_q0 + blk);
    it.base = __ldg(cu + vid);
    it.n = __ldg(cu + vid + 1) - it.base;
    it.nkv = (it.n + BKV - 1) / BKV;
    it.has_b = it.q0 + 128 < it.n;
    return it;
}

// counters[0] = number of 256-query blocks, counters[1] = work counter, counters[2] = CTAs that have finished (both zero
// between launches: the schedule kernel zeroes them, the last CTA of every launch rewinds them); flags[item] != 0: the
// item needs the exact pass (raised by the SAFE = false launch, consumed by the SAFE = true launch).
template <bool TRAIN, bool SAFE>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t *__restrict__ cu,
                  const int32_t *__restrict__ item_video, const int32_t *__restrict__ item_q0,
                  int32_t *__restrict__ counters, int32_t *__restrict__ flags, void *__restrict__ out_v, float scale_log2e,
                  float *__restrict__ lse2, float keep_scale, uint32_t drop_thresh16, unsigned long long seed) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t *sQ = smem;                                          // [item parity][tile] x 16 KB
    uint8_t *sK = smem + 4 * (size_t)TILE_BYTES;                 // ring
    uint8_t *sV = sK + (size_t)KV_STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + (size_t)KV_STAGES * TILE_BYTES);
    uint64_t *q_full = bars, *q_empty = bars + 2;                // [2]
    uint64_t *k_full = bars + 4, *k_empty = k_full + KV_STAGES, *v_full = k_empty + KV_STAGES, *v_empty = v_full + KV_STAGES;
    uint64_t *s_full = v_empty + KV_STAGES, *s_empty = s_full + 2, *p_full = s_full + 4, *p_empty = s_full + 6;   // [tile]
    uint64_t *sched_full = s_full + 8, *sched_empty = sched_full + SCHED_RING;
    int32_t *sched_idx = reinterpret_cast<int32_t *>(sched_empty + SCHED_RING);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sched_idx + SCHED_RING);
    float *hm = reinterpret_cast<float *>(smem + (4 + 2 * (size_t)KV_STAGES) * TILE_BYTES + 512);   // [tile][parity][half][128] half-row maxima
    float *lx = hm + 2 * 2 * 2 * 128;                                                               // [tile][half][128] half-row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = __ldg(counters) * NH;

    if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&tmQKV);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(q_full + i, 1); tc::mbar_init(q_empty + i, 2); }
        for (int s = 0; s < KV_STAGES; ++s) {
            tc::mbar_init(k_full + s, 1); tc::mbar_init(k_empty + s, 2);
            tc::mbar_init(v_full + s, 1); tc::mbar_init(v_empty + s, 2);
        }
        for (int t = 0; t < 2; ++t) {
            tc::mbar_init(s_full + t, 1); tc::mbar_init(s_empty + t, 256);
            tc::mbar_init(p_full + t, 256); tc::mbar_init(p_empty + t, 1);
        }
        for (int i = 0; i < SCHED_RING; ++i) { tc::mbar_init(sched_full + i, 1); tc::mbar_init(sched_empty + i, 18); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, A2_TMEM_COLS); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 256, tP = tmem_base + 384;

    // Every role walks the same item sequence through the scheduler ring.
    auto next_item = [&](int it) -> int {
        const int slot = it % SCHED_RING;
        tc::mbar_wait(sched_full + slot, (it / SCHED_RING) & 1);
        return sched_idx[slot];
    };
    auto release_item = [&](int it) { tc::mbar_arrive(sched_empty + (it % SCHED_RING)); };

    if (warp == 3) {
        tc::setmaxnreg_dec<A2_REGS_DONOR>();
    } else if (warp < 3) {
        tc::setmaxnreg_dec<A2_REGS_CONTROL>();
        if (warp == 0) {                 // ===== work scheduler + TMA producer =====
            if (lane == 0) {
                uint32_t g = 0;                                           // running K/V tile counter (ring position)
                for (int it = 0;; ++it) {
                    const int slot = it % SCHED_RING;
                    tc::mbar_wait(sched_empty + slot, ((it / SCHED_RING) & 1) ^ 1);
                    int idx;
                    do idx = atomicAdd(counters + 1, 1);
                    while (SAFE && idx < n_items && *reinterpret_cast<volatile int32_t *>(flags + idx) == 0);   // exact pass: flagged items only
                    sched_idx[slot] = idx;
                    tc::mbar_arrive(sched_full + slot);
                    if (idx >= n_items) break;
                    const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
                    const int qb = it & 1, n_q = w.has_b ? 2 : 1;
                    tc::mbar_wait(q_empty + qb, ((it >> 1) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(q_full + qb, (uint32_t)n_q * TILE_BYTES);
                    for (int t = 0; t < n_q; ++t)
                        tc::tma_load_2d(sQ + (size_t)(qb * 2 + t) * TILE_BYTES, &tmQKV, q_full + qb, w.head * HD, w.base + w.q0 + t * 128);
                    for (int j = 0; j < w.nkv; ++j, ++g) {
                        const int s = g % KV_STAGES;
                        const uint32_t ph = ((g / KV_STAGES) & 1) ^ 1;
                        tc::mbar_wait(k_empty + s, ph);
                        tc::mbar_arrive_expect_tx(k_full + s, TILE_BYTES);
                        tc::tma_load_2d(sK + (size_t)s * TILE_BYTES, &tmQKV, k_full + s, DM + w.head * HD, w.base + j * BKV);
                        tc::mbar_wait(v_empty + s, ph);
                        tc::mbar_arrive_expect_tx(v_full + s, TILE_BYTES);
                        tc::tma_load_2d(sV + (size_t)s * TILE_BYTES, &tmQKV, v_full + s, 2 * DM + w.head * HD, w.base + j * BKV);
                    }
                }
            }
        } else {                         // ===== MMA issuer of tile t (whole warp, warp-uniform control flow, one elected lane) =====
            const int t = warp - 1;
            constexpr uint32_t IDESC_QK = tc::make_idesc(1, 128, BKV, 0, 0);   // S[128 x 128], A and B K-major
            constexpr uint32_t IDESC_PV = tc::make_idesc(1, 128, HD, 0, 1);    // O[128 x 64], A = P in TMEM, B = V MN-major
            const uint32_t q_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);
            const uint32_t k_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024);
            const uint32_t v_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024);
            const uint32_t hi = (uint32_t)(tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024) >> 32);
            auto desc_at = [hi](uint32_t lo, uint32_t off) -> uint64_t {       // derived at the point of use (see vsum_attn_tc05.cu)
                uint32_t l;
                asm volatile("add.u32 %0, %1, %2;" : "=r"(l) : "r"(lo), "r"(off));
                return ((uint64_t)hi << 32) | l;
            };
            constexpr uint32_t TILE16 = TILE_BYTES >> 4;
            const uint32_t tS_t = tS + (uint32_t)(t * 128), tO_t = tO + (uint32_t)(t * HD), tP_t = tP + (uint32_t)(t * 64);
            uint32_t g0 = 0;             // ring position of the item's first K/V tile
            uint32_t cs = 0;             // S tiles issued so far for this tile slot
            uint32_t cp = 0;             // PV products issued so far
            for (int it = 0;; ++it) {
                const int idx = next_item(it);
                __syncwarp();
                if (lane == 0) release_item(it);
                if (idx >= n_items) break;
                const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
                const int qb = it & 1;
                if (t == 1 && !w.has_b) {        // no second tile: keep the shared barriers' arrival counts whole
                    for (int j = 0; j < w.nkv; ++j) {
                        const uint32_t g = g0 + j;
                        const int s = g % KV_STAGES;
                        const uint32_t ph = (g / KV_STAGES) & 1;
                        tc::mbar_wait(k_full + s, ph);
                        tc::mbar_wait(v_full + s, ph);
                        if (lane == 0) { tc::mbar_arrive(k_empty + s); tc::mbar_arrive(v_empty + s); }
                    }
                    if (lane == 0) tc::mbar_arrive(q_empty + qb);
                    g0 += w.nkv;
                    continue;
                }
                auto issue_qk = [&](int j) {     // S_t = Q_t K(j)^T once the rows hold the previous S_t in registers
                    const uint32_t g = g0 + j;
                    const int s = g % KV_STAGES;
                    tc::mbar_wait(k_full + s, (g / KV_STAGES) & 1);
                    tc::mbar_wait(s_empty + t, (cs & 1) ^ 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc::mma_f16_ss(tS_t, desc_at(q_lo, (uint32_t)(qb * 2 + t) * TILE16 + k * 2),
                                           desc_at(k_lo, (uint32_t)s * TILE16 + k * 2), IDESC_QK, k != 0);
                        tc::mma_commit(s_full + t);
                        tc::mma_commit(k_empty + s);
                        if (j == w.nkv - 1) tc::mma_commit(q_empty + qb);
                    }
                    __syncwarp();
                    ++cs;
                };
                tc::mbar_wait(q_full + qb, (it >> 1) & 1);
                issue_qk(0);
                for (int j = 0; j < w.nkv; ++j) {
                    if (j + 1 < w.nkv) issue_qk(j + 1);
                    const uint32_t g = g0 + j;
                    const int s = g % KV_STAGES;
                    tc::mbar_wait(p_full + t, cp & 1);                    // P_t(j) stored (and O_t rescaled where needed)
                    tc::mbar_wait(v_full + s, (g / KV_STAGES) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            tc::mma_f16_ts(tO_t, tP_t + (uint32_t)(k * 8), desc_at(v_lo, (uint32_t)s * TILE16 + k * 128), IDESC_PV,
                                           (j | k) != 0);
                        tc::mma_commit(p_empty + t);
                        tc::mma_commit(v_empty + s);
                    }
                    __syncwarp();
                    ++cp;
                }
                g0 += w.nkv;
            }
        }
    } else {   // ===== softmax: two threads per query row, 64 key columns each =====
        tc::setmaxnreg_inc<A2_REGS_SOFTMAX>();
        const int grp = (warp - 4) >> 2;                          // warps 4..7, 8..11, 12..15, 16..19: each group covers the four lane quarters
        const int t = grp >> 1, hf = grp & 1, qd = warp & 3;      // tile, column half, TMEM lane quarter (= warp % 4 = SM sub-partition)
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t tS_r = tS + lane_off + (uint32_t)(t * 128 + hf * 64), tO_r = tO + lane_off + (uint32_t)(t * HD + hf * 32),
                       tP_r = tP + lane_off + (uint32_t)(t * 64 + hf * 32);
        const int pair_bar = 1 + t * 4 + qd;                      // named barrier of the two warps that share these rows
        float *hm_mine = hm + (t * 4 + hf) * 128 + r, *hm_other = hm + (t * 4 + (hf ^ 1)) * 128 + r;   // + parity * 256
        const float2 c2 = make_float2(scale_log2e, scale_log2e);
        uint32_t c = 0;                                           // tiles processed so far by this tile slot (barrier phases)
        A2_TDECL(10);
        for (int it = 0;; ++it) {
            const int idx = next_item(it);
            __syncwarp();
            if (lane == 0) release_item(it);
            if (idx >= n_items) break;
            const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
            if (t == 1 && !w.has_b) continue;
            const int row = w.q0 + t * 128 + r;                   // query row inside the video
            float m_ref = 0.f, l_part = 0.f, m_prev = 0.f;        // exponent reference, row sum of MY columns, my maximum of the previous tile
            for (int j = 0; j < w.nkv; ++j, ++c) {
                A2_TMARK(0);
                tc::mbar_wait(s_full + t, c & 1);
                A2_TMARK(1);
                tc::tc_fence_after();
                uint32_t sa[32], sb[32];
                tc::tmem_ld32(tS_r, sa);
                tc::tmem_ld32(tS_r + 32, sb);
                bool p_free = false;                              // PV(j-1) has completed: P_t may be overwritten, O_t is stable
                auto move_reference = [&](float m_row, bool low_side_too) {   // same decision in both threads of the row
                    const bool move = (m_row - m_ref > VSUM_A2_WINDOW) || (low_side_too && m_row - m_ref < -VSUM_A2_WINDOW);
                    if (__any_sync(0xffffffffu, move)) {           // rare
                        if (!p_free) { tc::mbar_wait(p_empty + t, (c & 1) ^ 1); tc::tc_fence_after(); p_free = true; }
                        float alpha = 1.0f;
                        if (move) { alpha = ex2f(m_ref - m_row); m_ref = m_row; }
                        if (j > 0) {                              // my 32 columns of O_t and my half of the row sum
                            l_part *= alpha;
                            uint32_t o[32];
                            tc::tmem_ld32(tO_r, o);
                            tmem_wait_ld_on(o);
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tc::tmem_st32(tO_r, o);
                            tc::tmem_wait_st();
                        }
                    }
                };
                if (!SAFE && j > 0)      // reference for this tile from BOTH halves' maxima of the previous tile (visible: see the header)
                    move_reference(fmaxf(m_prev, hm_other[((c - 1) & 1) * 256]), j == 1);
                tmem_wait_ld_on(sa);
                tmem_wait_ld_on(sb);
                A2_TMARK(2);
                const int valid = w.n - j * BKV - hf * 64;        // keys of my half-tile inside the video (may be <= 0)
                if (valid < 64) {                                 // last tile of the video: next video's rows / TMA zero fill
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= valid) sa[i] = 0xff800000u;
                        if (32 + i >= valid) sb[i] = 0xff800000u;
                    }
                }
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(sb[i]), __uint_as_float(sb[i + 1]));
                const float m_half = mx * scale_log2e;
                hm_mine[(c & 1) * 256] = m_half;
                if (SAFE) {              // exact: this tile's own row maximum, exchanged now
                    tc::bar_sync(pair_bar, 64);
                    move_reference(fmaxf(m_half, hm_other[(c & 1) * 256]), j == 0);
                } else if (row < w.n && (m_half - m_ref > VSUM_A2_DANGER || (j == 0 && m_half < -VSUM_A2_DANGER && valid > 0))) {
                    *reinterpret_cast<volatile int32_t *>(flags + idx) = 1;   // the exact pass redoes this item
                }
                m_prev = m_half;
                tc::tc_fence_before();
                tc::mbar_arrive(s_empty + t);                     // S_t is in registers (and my maximum published): QK(j+1) may overwrite it
                A2_TMARK(3);

                const float2 nm2 = make_float2(-m_ref, -m_ref);
                float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                auto chunk = [&](const uint32_t (&s)[32], int ch) {
                    uint32_t wv[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 x = ffma2(make_float2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1])), c2, nm2);
                        const bool poly = VSUM_A2_POLY_PERIOD > 0 && (e % (VSUM_A2_POLY_PERIOD > 0 ? VSUM_A2_POLY_PERIOD : 1)) == VSUM_A2_POLY_PERIOD - 1;
                        const float2 p = poly ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                        ps[e & 3] = fadd2(ps[e & 3], p);
                        wv[e] = pack_bf16x2(p.x, p.y);
                    }
                    if (TRAIN && drop_thresh16 != 0) {            // one 64-bit draw decides 4 consecutive keys; the row sum stays un-dropped
                        const uint32_t th2 = drop_thresh16 * 0x00010001u;
#pragma unroll
                        for (int gq = 0; gq < 8; ++gq) {
                            const int key = j * BKV + hf * 64 + ch * 32 + 4 * gq;
                            const unsigned long long z = dropout_bits64(seed, attn_drop_group_index(w.base + row, w.head, NH, key >> 2));
                            wv[2 * gq] &= __vcmpgeu2((uint32_t)z, th2);
                            wv[2 * gq + 1] &= __vcmpgeu2((uint32_t)(z >> 32), th2);
                        }
                    }
                    if (!p_free) { tc::mbar_wait(p_empty + t, (c & 1) ^ 1); tc::tc_fence_after(); p_free = true; }
                    tmem_st16(tP_r + (uint32_t)(ch * 16), wv);
                };
                chunk(sa, 0);
                A2_TMARK(4);
                chunk(sb, 1);
                A2_TMARK(5);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                tc::mbar_arrive(p_full + t);
                const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
                l_part += pq.x + pq.y;
                A2_TMARK(6);
            }
#ifdef VSUM_A2_TIMING
            if (blockIdx.x == 0 && lane == 0 && w.nkv >= 16 && tcount++ == 1)
                printf("softmax warp %2d nkv %d | other %lld | wait S %lld | agree + ld %lld | mask + max + publish + arrive %lld | chunk0 %lld | chunk1 %lld | "
                       "wait st + arrive %lld (clk per tile)\n", warp, w.nkv, tph[0] / w.nkv, tph[1] / w.nkv, tph[2] / w.nkv, tph[3] / w.nkv,
                       tph[4] / w.nkv, tph[5] / w.nkv, tph[6] / w.nkv);
            for (int i = 0; i < 10; ++i) tph[i] = 0;
            tmark = clock64();
#endif
            // ---- epilogue: O_t / l -> global, my 32 of the head's 64 columns
            tc::mbar_wait(p_empty + t, (c & 1) ^ 1);              // the last PV product of the item has completed
            tc::tc_fence_after();
            lx[(t * 2 + hf) * 128 + r] = l_part;
            tc::bar_sync(pair_bar, 64);
            const float l_tot = l_part + lx[(t * 2 + (hf ^ 1)) * 128 + r];
            const float inv = (TRAIN ? keep_scale : 1.0f) / l_tot;
            if (TRAIN && hf == 0 && row < w.n) lse2[(int64_t)(w.base + row) * NH + w.head] = m_ref + log2f(l_tot);
            uint32_t o[32];
            tc::tmem_ld32(tO_r, o);
            tmem_wait_ld_on(o);
            if (row < w.n) {
                if (TRAIN) {   // fp32 output: the backward's delta = rowsum(dO o O) must not see a rounded O
                    float *dst = reinterpret_cast<float *>(out_v) + (int64_t)(w.base + row) * DM + w.head * HD + hf * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4 *>(dst + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                                                           __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                } else {
                    __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(out_v) + (int64_t)(w.base + row) * DM + w.head * HD + hf * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 pk;
                        pk.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                        pk.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                        pk.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                        pk.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                        *reinterpret_cast<uint4 *>(dst + i) = pk;
                    }
                }
            }
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, A2_TMEM_COLS); }
    // The last CTA to leave rewinds the work counter, so one schedule serves every launch over the same batch.
    if (threadIdx.x == 0 && atomicAdd(counters + 2, 1) == (int)gridDim.x - 1) {
        counters[1] = 0;
        counters[2] = 0;
        __threadfence();
    }
}

// 256-query blocks of all videos, longest video first (counting sort on the number of 128-key tiles), the work
// counter reset and the exact-pass flags cleared.  One block.
__global__ void __launch_bounds__(1024)
attn2_schedule_kernel(const int32_t *__restrict__ cu, int B, int32_t *__restrict__ item_video, int32_t *__restrict__ item_q0,
                      int32_t *__restrict__ counters, int32_t *__restrict__ flags, int max_blocks) {
    __shared__ int hist[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < max_blocks * NH; i += blockDim.x) flags[i] = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < B; v += blockDim.x) {
        const int n = __ldg(cu + v + 1) - __ldg(cu + v);
        if (n > 0) atomicAdd(&hist[min((n + BKV - 1) / BKV, 255)], (n + BQ2 - 1) / BQ2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 255; k >= 0; --k) { const int h = hist[k]; hist[k] = run; run += h; }
        counters[0] = min(run, max_blocks);
        counters[1] = 0;
        counters[2] = 0;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < B; v += blockDim.x) {
        const int n = __ldg(cu + v + 1) - __ldg(cu + v);
        if (n <= 0) continue;
        const int nb = (n + BQ2 - 1) / BQ2;
        const int start = atomicAdd(&hist[min((n + BKV - 1) / BKV, 255)], nb);
        for (int b = 0; b < nb; ++b)
            if (start + b < max_blocks) { item_video[start + b] = v; item_q0[start + b] = b * BQ2; }
    }
}

}  // namespace

// item_video[max_blocks], item_q0[max_blocks], counters[4], flags[4 * max_blocks]
size_t attention2_scratch_ints(int64_t T, int B) { return 6 * (size_t)(T / BQ2 + B) + 4; }

// Work list for launch_attention2_tc05 (valid for any number of launches over the same cu_seqlens).
int launch_attn2_schedule(const int32_t *cu_seqlens, int B, int64_t T, int32_t *scratch, cudaStream_t s) {
    if (T == 0 || B == 0) return VSUM_OK;
    const int max_blocks = (int)(T / BQ2 + B);
    attn2_schedule_kernel<<<1, 1024, 0, s>>>(cu_seqlens, B, scratch, scratch + max_blocks, scratch + 2 * (size_t)max_blocks,
                                             scratch + 2 * (size_t)max_blocks + 4, max_blocks);
    VSUM_LAUNCH_OK("attn2_schedule_kernel");
    return VSUM_OK;
}

// qkv [T,768] bf16 -> out [T,256] (bf16; fp32 when lse2 != NULL, the training variant: log2-domain log-sum-exp [T,4]
// out, dropout on P with the grouped hash).  scratch: attention2_scratch_ints(T, B) int32 filled by launch_attn2_schedule.
// Two launches: the fast pass, then the exact pass over the items the fast pass flagged (normally none).
int launch_attention2_tc05(const __nv_bfloat16 *qkv, const int32_t *cu_seqlens, int B, int64_t T, float scale, void *out,
                           int32_t *scratch, cudaStream_t s, float *lse2, float drop_p, unsigned long long seed) {
    if (T == 0 || B == 0) return VSUM_OK;
    const int max_blocks = (int)(T / BQ2 + B);
    int32_t *item_video = scratch, *item_q0 = scratch + max_blocks, *counters = scratch + 2 * (size_t)max_blocks, *flags = counters + 4;
    CUtensorMap tm;
    int rc = make_tensor_map_2d(&tm, qkv, 2, 3 * DM, (uint64_t)T, (uint64_t)3 * DM * 2, 64, 128);
    if (rc) return rc;
    static int n_sm[64];
    int dev = 0;
    VSUM_CUDA_OK(cudaGetDevice(&dev));
    VSUM_ONCE_PER_DEVICE(
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev)));
    {   // setmaxnreg moves registers inside the CTA's own allocation: what the control warps release must cover what the softmax warps acquire
        static std::atomic<int> checked{0};
        if (!checked.load(std::memory_order_relaxed)) {
            cudaFuncAttributes fa;
            VSUM_CUDA_OK(cudaFuncGetAttributes(&fa, attn2_tc05_kernel<false, false>));
            cudaFuncAttributes fb;
            VSUM_CUDA_OK(cudaFuncGetAttributes(&fb, attn2_tc05_kernel<true, false>));
            const int e = fa.numRegs < fb.numRegs ? fa.numRegs : fb.numRegs;
            VSUM_REQUIRE(96 * (e - A2_REGS_CONTROL) + 32 * (e - A2_REGS_DONOR) >= 512 * (A2_REGS_SOFTMAX - e), VSUM_EUNSUPPORTED,
                         "attn2_tc05_kernel was compiled with %d registers per thread: the softmax warps could not grow to %d", e, A2_REGS_SOFTMAX);
            checked.store(1, std::memory_order_relaxed);
        }
    }
    int grid = n_sm[dev & 63] - scorer_sm_reserve();
    if (grid < 1) grid = 1;
    const int max_items = max_blocks * NH;
    if (grid > max_items) grid = max_items;
    ProfScope prof(PROF_ATTN, s);
    const float sl2 = scale * 1.4426950408889634f;
    if (lse2) {
        const uint32_t thresh = attn_drop_thresh16(drop_p);
        const float ks = 65536.0f / (float)(65536u - thresh);
        attn2_tc05_kernel<true, false><<<grid, A2_THREADS, A2_SMEM, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2, lse2, ks, thresh, seed);
        VSUM_LAUNCH_OK("attn2_tc05_kernel");
        attn2_tc05_kernel<true, true><<<grid, A2_THREADS, A2_SMEM, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2, lse2, ks, thresh, seed);
    } else {
        attn2_tc05_kernel<false, false><<<grid, A2_THREADS, A2_SMEM, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2, nullptr, 1.0f, 0u, 0ull);
        VSUM_LAUNCH_OK("attn2_tc05_kernel");
        attn2_tc05_kernel<false, true><<<grid, A2_THREADS, A2_SMEM, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2, nullptr, 1.0f, 0u, 0ull);
    }
    VSUM_LAUNCH_OK("attn2_tc05_kernel (exact pass)");
    return VSUM_OK;
}

}  // namespace vsum

extern "C" size_t vsum_attention_scratch_ints(int64_t T, int32_t B) {
    if (T < 0 || B < 0) return 0;
    const size_t a = vsum::attention2_scratch_ints(T, B), b = 2 * (size_t)(T / 128 + B) + 1;   // two-tile kernel / one-tile kernel
    return a > b ? a : b;
}
