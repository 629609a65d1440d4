// Variable-length (packed, padding-free) multi-head self-attention forward for the scorer on sm_100a, second
// generation.  Replaces src/model/simnet.py:155-161 (QK^T * d_model^-0.5 -> softmax -> PV; dropout on P at
// simnet.py:159 in the TRAIN variant) for d_model 256, 4 heads of 64.
//
// PERSISTENT kernel, one CTA per SM, work items = (video, head, 256 queries) handed out longest-video-first by an
// atomic counter.  A CTA works on TWO 128-query tiles (A, B) of its item; a "unit" is one (tile, 128-key block) pair and
// the units of all items form one sequence A0 B0 A1 B1 ... that every role walks in the same order.
//   warps 0-3   softmax of tile A, warps 4-7 softmax of tile B: one thread per query row
//   warp 8      work scheduler (atomicAdd -> shared-memory ring) and TMA producer: Q tiles of the item (double-buffered
//               across items), K / V tiles through a shared ring
//   warp 9      tcgen05.mma issuer of all QK^T products
//   warp 10/11  tcgen05.mma issuer of the PV products of tile A / tile B (warp 10 also owns the TMEM allocation)
// (the control warps carry the HIGH warp ids: the sub-partition schedulers favour them, and an issuer that has to wait
// for the softmax warps' idle slots delays every tile)
// Tensor memory (512 columns): THREE S buffers fp32 [128 x 128] (0..383) used round-robin by the unit sequence, O_A,
// O_B fp32 [128 x 64] (384..511).  The probabilities of a unit are written IN PLACE over the first 64 columns of its S
// buffer (bf16 pairs, tcgen05.st) and consumed from there as the TMEM A operand of the PV MMA, so P never touches shared
// memory and every K / V tile is fetched once per 256 queries (profiles/r01_microbench_mma_rate.txt: the one-tile kernel
// ran at 82 % of its shared-memory bandwidth).  With three buffers for two tiles the issuer runs QK^T three units ahead
// of PV: S of unit u+3 goes into the buffer PV(u) has just read (the tensor pipe executes in order), i.e. it depends on
// the OTHER tile's previous step, not on the softmax that will consume it -- a row finds its next S ready when it has
// stored its P, where the two-buffer version of this kernel waited ~20 % of the time (profiles/r02_attn2_*.txt).
//
// One-pass streaming softmax against the FIXED reference 0.  With the softmax scale and log2(e) folded into W_q
// (PRESCALED: the scorer's bf16 copy of the weights, vsum_scorer.cu) a score IS the exponent: the fast pass computes
// P = 2^S straight from the registers tcgen05.ld filled -- no row maximum, no scale, no subtraction, no rescaling of
// O -- and the row sum.  Softmax is shift-invariant and fp32 / bf16 keep their relative precision, so this equals the
// usual running-maximum form as long as no exponential overflows and the row does not underflow as a whole, i.e. while
// the row's largest score stays within about +-90 log2 units (+-62 nats) of zero; attention logits of this model are
// O(1).  A row checks exactly that on its sums (a tile sum >= 2^90 or not finite, a row sum < 2^-80) and raises its
// item's flag, and a second launch of the same kernel (SAFE = true: one extra sweep over S for the tile maximum, classic
// online softmax with a lazily rescaled O) recomputes the flagged items; with no flag raised that launch ends after
// reading one flag per item.
// The row streams S in four 32-column chunks whose tensor-memory loads overlap the exponentials of the previous chunk.
// A share of the exponentials runs as a degree-3 polynomial on the FMA pipe (the MUFU pipe, 16 ex2 / clk / SM, is the
// limiter at head_dim 64: profiles/r01_microbench_mufu_ex2.txt).
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {
namespace {

constexpr int HD = 64, DM = 256, NH = 4;
constexpr int BQ2 = 256, BKV = 128;
constexpr int TILE_BYTES = 128 * 128;            // 128 rows x 64 bf16
constexpr int A2_THREADS = 384;                  // TPR 1: 4 control warps + 8 softmax warps; 384 x 168 registers at launch
constexpr int A2_REGS_CONTROL = 80, A2_REGS_SOFTMAX = 208;   // setmaxnreg: 128 x (168 - 80) released >= 256 x (208 - 168) acquired
// TPR 2 (two threads per query row, 64 key columns each): 16 softmax warps = four per SM sub-partition + 4 control warps;
// 640 x 96 registers at launch, 128 x (96 - 64) released >= 512 x (104 - 96) acquired (setmaxnreg only moves registers inside
// the CTA's own allocation: an increase that the releases do not cover never completes)
constexpr int A2_THREADS2 = 640;
constexpr int A2_REGS_CONTROL2 = 64, A2_REGS_SOFTMAX2 = 104;
constexpr int A2_TMEM_COLS = 512;
#ifndef VSUM_A2_STAGES
#define VSUM_A2_STAGES 4
#endif
constexpr int KV_STAGES = VSUM_A2_STAGES;
constexpr int SCHED_RING = 4;
constexpr size_t A2_SMEM = (4 + 2 * (size_t)KV_STAGES) * TILE_BYTES + 512;   // Q (2 items x 2 tiles), K ring, V ring, barriers
constexpr size_t A2_SMEM2 = A2_SMEM + 4096;                                  // + partial row sums of the two-threads-per-row variant
#ifndef VSUM_A2_POLY_PERIOD
#define VSUM_A2_POLY_PERIOD 3        // every k-th pair of exponentials is a polynomial on the FMA pipe (0 = none); swept 0 / 2..5 on B200
#endif

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
// item needs the exact pass (raised by the SAFE = false launch, consumed by the SAFE = true launch), counters[3] = how
// many items are flagged.  Flags stay up until the next schedule: a later launch over the same batch (the next layer)
// redoes those items in its exact pass too, which costs time, not correctness.
template <bool TRAIN, bool SAFE, bool PRESCALED, int TPR = 1>
__global__ void __launch_bounds__(TPR == 2 ? A2_THREADS2 : A2_THREADS, 1)
attn2_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t *__restrict__ cu,
                  const int32_t *__restrict__ item_video, const int32_t *__restrict__ item_q0,
                  int32_t *__restrict__ counters, int32_t *__restrict__ flags, void *__restrict__ out_v, float scale_log2e,
                  float *__restrict__ lse2, float keep_scale, uint32_t drop_thresh16, unsigned long long seed) {
    if (SAFE && *reinterpret_cast<volatile int32_t *>(counters + 3) == 0) return;   // no item was flagged (every CTA sees the same count)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t *sQ = smem;                                          // [item parity][tile] x 16 KB
    uint8_t *sK = smem + 4 * (size_t)TILE_BYTES;                 // ring
    uint8_t *sV = sK + (size_t)KV_STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + (size_t)KV_STAGES * TILE_BYTES);
    uint64_t *q_full = bars, *q_empty = bars + 2;                // [item parity]
    uint64_t *k_full = bars + 4, *k_empty = k_full + KV_STAGES, *v_full = k_empty + KV_STAGES, *v_empty = v_full + KV_STAGES;
    uint64_t *s_full = v_empty + KV_STAGES;                      // [tile][the tile's unit count % 3]: S of that unit is complete (a tile's rows
                                                                 // must see every phase of the barriers they wait on, so these are per tile)
    uint64_t *p_full = s_full + 6;                               // [tile][count % 3]: the P of that unit is stored (128 arrivals)
    uint64_t *pv_done = p_full + 6;                              // [tile]: a PV product of the tile has completed
    uint64_t *buf_free = pv_done + 2;                            // [S buffer]: the PV product that read P from the buffer has completed
    uint64_t *o_full = buf_free + 3;                             // [tile]: the last PV product of an item has completed (O_t is final)
    uint64_t *sched_full = o_full + 2, *sched_empty = sched_full + SCHED_RING;
    int32_t *sched_idx = reinterpret_cast<int32_t *>(sched_empty + SCHED_RING);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sched_idx + SCHED_RING);
    float *l_xch = reinterpret_cast<float *>(bars) + 128;         // TPR 2: [item parity][tile][half][128 rows] partial row sums (after 512 B of barriers)
    static_assert(TPR == 1 || (TPR == 2 && !TRAIN && !SAFE), "two threads per row: inference fast pass only");
    constexpr int NSOFT = 8 * TPR;                                // softmax warps; the control warps follow

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cw = warp - NSOFT;                                  // control warp index (0 producer, 1 QK issuer, 2 / 3 PV issuers)
    const int n_items = __ldg(counters) * NH;

    if (cw == 0 && lane == 0) tc::tma_prefetch_desc(&tmQKV);
    if (cw == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(q_full + i, 1); tc::mbar_init(q_empty + i, 1); }
        for (int s = 0; s < KV_STAGES; ++s) {
            tc::mbar_init(k_full + s, 1); tc::mbar_init(k_empty + s, 1);
            tc::mbar_init(v_full + s, 1); tc::mbar_init(v_empty + s, 2);
        }
        for (int i = 0; i < 6; ++i) { tc::mbar_init(s_full + i, 1); tc::mbar_init(p_full + i, 128 * TPR); }
        for (int i = 0; i < 3; ++i) tc::mbar_init(buf_free + i, 1);
        for (int t = 0; t < 2; ++t) tc::mbar_init(o_full + t, 1);
        for (int t = 0; t < 2; ++t) tc::mbar_init(pv_done + t, 1);
        for (int i = 0; i < SCHED_RING; ++i) { tc::mbar_init(sched_full + i, 1); tc::mbar_init(sched_empty + i, NSOFT + 3); }
        tc::fence_barrier_init();
    }
    if (cw == 2) { tc::tmem_alloc(tmem_slot, A2_TMEM_COLS); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tO = tmem_base + 384;                         // S buffer b at tmem_base + 128 b

    // Every role walks the same item sequence through the scheduler ring.
    auto next_item = [&](int it) -> int {
        const int slot = it % SCHED_RING;
        tc::mbar_wait(sched_full + slot, (it / SCHED_RING) & 1);
        return sched_idx[slot];
    };
    auto release_item = [&](int it) { tc::mbar_arrive(sched_empty + (it % SCHED_RING)); };

    if (cw >= 0) {
        if (TPR == 2) tc::setmaxnreg_dec<A2_REGS_CONTROL2>(); else tc::setmaxnreg_dec<A2_REGS_CONTROL>();
        if (cw == 0) {                   // ===== work scheduler + TMA producer =====
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
        } else if (cw == 1) {            // ===== QK^T issuer (whole warp, warp-uniform control flow, one elected lane) =====
            // One thread cannot issue both products of both tiles fast enough (12 MMAs + commits + three waits cost it
            // ~1000 clk per unit, tools/microbench/mma_hazard.cu), so QK^T and the two tiles' PV have a warp each.
            constexpr uint32_t IDESC_QK = tc::make_idesc(1, 128, BKV, 0, 0);   // S[128 x 128], A and B K-major
            const uint32_t q_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);
            const uint32_t k_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024);
            const uint32_t hi = (uint32_t)(tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024) >> 32);
            auto desc_at = [hi](uint32_t lo, uint32_t off) -> uint64_t {       // derived at the point of use (see vsum_attn_tc05.cu)
                uint32_t l;
                asm volatile("add.u32 %0, %1, %2;" : "=r"(l) : "r"(lo), "r"(off));
                return ((uint64_t)hi << 32) | l;
            };
            constexpr uint32_t TILE16 = TILE_BYTES >> 4;
            uint32_t n_qk = 0, cs0 = 0, cs1 = 0, g0 = 0;   // units issued (S buffer = unit % 3), S tiles issued per tile, K/V ring position
            for (int it = 0;; ++it) {
                const int idx = next_item(it);
                __syncwarp();
                if (lane == 0) release_item(it);
                if (idx >= n_items) break;
                const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
                const int qb = it & 1, n_q = w.has_b ? 2 : 1;
                tc::mbar_wait(q_full + qb, (it >> 1) & 1);
                for (int j = 0; j < w.nkv; ++j) {
                    const uint32_t g = g0 + (uint32_t)j;
                    const int s = g % KV_STAGES;
                    tc::mbar_wait(k_full + s, (g / KV_STAGES) & 1);
                    for (int t = 0; t < n_q; ++t, ++n_qk) {
                        const uint32_t buf = n_qk % 3;
                        if (n_qk >= 3) tc::mbar_wait(buf_free + buf, (n_qk / 3 - 1) & 1);   // the PV product of unit n_qk - 3 has read its P
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k)
                                tc::mma_f16_ss(tmem_base + buf * 128, desc_at(q_lo, (uint32_t)(qb * 2 + t) * TILE16 + k * 2),
                                               desc_at(k_lo, (uint32_t)s * TILE16 + k * 2), IDESC_QK, k != 0);
                            tc::mma_commit(s_full + t * 3 + (t ? cs1 : cs0) % 3);
                            if (t == n_q - 1) tc::mma_commit(k_empty + s);
                            if (j == w.nkv - 1 && t == n_q - 1) tc::mma_commit(q_empty + qb);
                        }
                        __syncwarp();
                        if (t) ++cs1; else ++cs0;
                    }
                }
                g0 += (uint32_t)w.nkv;
            }
        } else {                         // ===== PV issuer of tile t (warp 10: A, warp 11: B) =====
            const int t = cw - 2;
            constexpr uint32_t IDESC_PV = tc::make_idesc(1, 128, HD, 0, 1);    // O[128 x 64], A = P in TMEM, B = V MN-major
            const uint32_t v_lo = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024);
            const uint32_t hi = (uint32_t)(tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024) >> 32);
            auto desc_at = [hi](uint32_t lo, uint32_t off) -> uint64_t {
                uint32_t l;
                asm volatile("add.u32 %0, %1, %2;" : "=r"(l) : "r"(lo), "r"(off));
                return ((uint64_t)hi << 32) | l;
            };
            constexpr uint32_t TILE16 = TILE_BYTES >> 4;
            const uint32_t tO_t = tO + (uint32_t)(t * HD);
            uint32_t n_unit0 = 0, cpt = 0, g0 = 0;         // first unit of the item, P tiles of this tile consumed, K/V ring position
            for (int it = 0;; ++it) {
                const int idx = next_item(it);
                __syncwarp();
                if (lane == 0) release_item(it);
                if (idx >= n_items) break;
                const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
                const int n_q = w.has_b ? 2 : 1;
                const uint32_t unit0 = n_unit0;
                n_unit0 += (uint32_t)(w.nkv * n_q);
                for (int j = 0; j < w.nkv; ++j) {
                    const uint32_t g = g0 + (uint32_t)j;
                    const int s = g % KV_STAGES;
                    if (t == 1 && !w.has_b) {       // no second tile: keep the V ring's arrival count whole
                        tc::mbar_wait(v_full + s, (g / KV_STAGES) & 1);
                        if (lane == 0) tc::mbar_arrive(v_empty + s);
                        continue;
                    }
                    const uint32_t buf = (unit0 + (uint32_t)(j * n_q + t)) % 3;
                    tc::mbar_wait(p_full + t * 3 + cpt % 3, (cpt / 3) & 1);        // P_t(j) stored (and O_t rescaled where needed)
                    tc::mbar_wait(v_full + s, (g / KV_STAGES) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            // P of keys [16 k, 16 k + 16): 8 columns of packed pairs; with two threads per row the second thread's
                            // probabilities (keys 64..127) sit in ITS half of the S buffer, columns 64..95
                            tc::mma_f16_ts(tO_t, tmem_base + buf * 128 + (uint32_t)(TPR == 2 && k >= 4 ? 64 + (k - 4) * 8 : k * 8),
                                           desc_at(v_lo, (uint32_t)s * TILE16 + k * 128), IDESC_PV, (j | k) != 0);
                        tc::mma_commit(pv_done + t);
                        if (j == w.nkv - 1) tc::mma_commit(o_full + t);
                        tc::mma_commit(buf_free + buf);
                        tc::mma_commit(v_empty + s);
                    }
                    __syncwarp();
                    ++cpt;
                }
                g0 += (uint32_t)w.nkv;
            }
        }
    } else if constexpr (TPR == 2) {   // ===== softmax: TWO threads per query row, 64 key columns each =====
        // Four softmax warps on every SM sub-partition instead of two: while one waits for S, for a tensor-memory load or
        // on the MUFU queue, three others can issue.  The fixed exponent reference makes the two threads of a row fully
        // independent inside an item -- no row maximum to agree on, no rescaling of the shared accumulator; they meet
        // once per item to add their partial row sums.  Thread h of a row reads S columns [64 h, 64 h + 64) and writes
        // its probabilities over the first 32 columns of that SAME half (packed pairs), so it only ever overwrites
        // scores it has already loaded; the PV issuer reads keys 64..127 from columns 64..95.
        tc::setmaxnreg_inc<A2_REGS_SOFTMAX2>();
        const int hf = warp >> 3, t = (warp >> 2) & 1, qd = warp & 3;     // column half, tile, TMEM lane quarter
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t tO_r = tO + lane_off + (uint32_t)(t * HD + hf * 32);    // my 32 of the tile's 64 output columns
        const int pair_bar = 1 + t * 4 + qd;                              // named barrier of the two warps that share these rows
        uint32_t c3 = 0, cpar = 0, n_unit0 = 0, n_done = 0;
        for (int it = 0;; ++it) {
            const int idx = next_item(it);
            __syncwarp();
            if (lane == 0) release_item(it);
            if (idx >= n_items) break;
            const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
            const int n_q = w.has_b ? 2 : 1;
            const uint32_t unit0 = n_unit0;
            n_unit0 += (uint32_t)(w.nkv * n_q);
            if (t == 1 && !w.has_b) continue;
            const int row = w.q0 + t * 128 + r;
            float l_part = 0.f;
            bool danger = false;
            uint32_t ub = (unit0 + (uint32_t)t) % 3;
            uint32_t sa[32], sb[32];
            for (int j = 0; j < w.nkv; ++j) {
                const uint32_t tS_r = tmem_base + lane_off + ub * 128 + (uint32_t)(hf * 64);   // my half of my row of the unit's S / P buffer
                tc::mbar_wait(s_full + t * 3 + c3, cpar);
                tc::tc_fence_after();
                const int valid = w.n - j * BKV - hf * 64;        // keys of my half inside the video (<= 0: none)
                float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                auto chunk = [&](uint32_t (&s)[32], int ch) {
                    if (valid < 64) {                              // last tile of the video: keys past its end
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ch * 32 + i >= valid) s[i] = 0xff800000u;
                    }
                    uint32_t wv[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float2 x = make_float2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1]));
                        if (!PRESCALED) x = ffma2(x, make_float2(scale_log2e, scale_log2e), make_float2(0.f, 0.f));
                        const bool poly = VSUM_A2_POLY_PERIOD > 0 && (e % (VSUM_A2_POLY_PERIOD > 0 ? VSUM_A2_POLY_PERIOD : 1)) == VSUM_A2_POLY_PERIOD - 1;
                        const float2 pp = poly ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                        ps[e & 3] = fadd2(ps[e & 3], pp);
                        wv[e] = pack_bf16x2(pp.x, pp.y);
                    }
                    tmem_st16(tS_r + (uint32_t)(ch * 16), wv);
                };
                tc::tmem_ld32(tS_r, sa);
                tmem_wait_ld_on(sa);
                tc::tmem_ld32(tS_r + 32, sb);
                chunk(sa, 0);
                tmem_wait_ld_on(sb);
                const uint32_t p_slot = t * 3 + c3;
                ub += (uint32_t)n_q; if (ub >= 3) ub -= 3;
                if (++c3 == 3) { c3 = 0; cpar ^= 1; }
                chunk(sb, 1);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                tc::mbar_arrive(p_full + p_slot);
                const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
                const float psum = pq.x + pq.y;
                danger |= !(psum < 1.2e27f);
                l_part += psum;
            }
            // the two partial row sums meet (slots alternate with the item parity: the partner has read the previous value of a
            // slot before it can arrive at the barrier that precedes the next write to it)
            float *slot = l_xch + (((n_done & 1) * 2 + t) * 2) * 128 + r;
            slot[hf * 128] = l_part;
            tc::bar_sync(pair_bar, 64);
            const float l_run = l_part + slot[(hf ^ 1) * 128];
            if (row < w.n && (danger || !(l_run >= 8.3e-25f)) && atomicExch(flags + idx, 1) == 0) atomicAdd(counters + 3, 1);
            tc::mbar_wait(o_full + t, n_done & 1);
            ++n_done;
            tc::tc_fence_after();
            const float inv = 1.0f / l_run;
            uint32_t o[32];
            tc::tmem_ld32(tO_r, o);
            tmem_wait_ld_on(o);
            if (row < w.n) {
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
            tc::tc_fence_before();
        }
    } else {   // ===== softmax: one thread per query row =====
        tc::setmaxnreg_inc<A2_REGS_SOFTMAX>();
        const int t = warp >> 2, qd = warp & 3;                   // tile, TMEM lane quarter (= SM sub-partition)
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t tO_r = tO + lane_off + (uint32_t)(t * HD);
        const float2 c2 = make_float2(scale_log2e, scale_log2e);
        uint32_t c = 0, c3 = 0, cpar = 0;                         // units of this tile processed so far; c % 3 and (c / 3) & 1 (barrier phases)
        uint32_t n_unit0 = 0;                                     // position of the item's first unit in the CTA's unit sequence
        uint32_t n_done = 0;                                      // items of this tile finished (o_full phases)
        uint32_t pv_seen = 0;                                     // SAFE: PV completions of this tile consumed so far
        A2_TDECL(10);
        for (int it = 0;; ++it) {
            const int idx = next_item(it);
            __syncwarp();
            if (lane == 0) release_item(it);
            if (idx >= n_items) break;
            const ItemInfo w = decode_item(idx, cu, item_video, item_q0);
            const int n_q = w.has_b ? 2 : 1;
            const uint32_t unit0 = n_unit0;
            n_unit0 += (uint32_t)(w.nkv * n_q);
            if (t == 1 && !w.has_b) continue;
            const int row = w.q0 + t * 128 + r;                   // query row inside the video
            float m_ref = 0.f, l_run = 0.f;                       // exponent reference (moves in the exact pass only), row sum
            bool danger = false;
            uint32_t ub = (unit0 + (uint32_t)t) % 3;              // S / P buffer of my current unit: (unit0 + j n_q + t) % 3
            uint32_t sa[32], sb[32];
            for (int j = 0; j < w.nkv; ++j, ++c) {
                const uint32_t tS_r = tmem_base + lane_off + ub * 128;   // my row of the unit's S / P buffer
                A2_TMARK(0);
                tc::mbar_wait(s_full + t * 3 + c3, cpar);
                tc::tc_fence_after();
                A2_TMARK(1);
                const int valid = w.n - j * BKV;                  // keys of this tile inside the video (>= 1)
                auto mask_tail = [&](uint32_t (&s)[32], int ch) {  // last tile of the video: keys past its end (next video's rows / TMA zero fill)
                    if (valid < BKV) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ch * 32 + i >= valid) s[i] = 0xff800000u;
                    }
                };
                if (SAFE) {              // exact pass: one extra sweep over S for this tile's maximum, classic online softmax
                    float mx0 = -INFINITY;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        tc::tmem_ld32(tS_r + ch * 32, sa);
                        tmem_wait_ld_on(sa);
                        mask_tail(sa, ch);
#pragma unroll
                        for (int i = 0; i < 32; i += 2) mx0 = fmax3(mx0, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
                    }
                    const float m_row = mx0 * scale_log2e;
                    const bool move = j == 0 || m_row - m_ref > 8.0f;
                    if (j > 0) {         // every PV completion is consumed in order: PV(j-1) has completed, O_t is stable
                        tc::mbar_wait(pv_done + t, pv_seen & 1);
                        ++pv_seen;
                        tc::tc_fence_after();
                    }
                    if (__any_sync(0xffffffffu, move)) {
                        float alpha = 1.0f;
                        if (move) { alpha = j == 0 ? 0.f : ex2f(m_ref - m_row); m_ref = m_row; }
                        if (j > 0) {
                            l_run *= alpha;
#pragma unroll 1
                            for (int hc = 0; hc < 2; ++hc) {
                                uint32_t o[32];
                                tc::tmem_ld32(tO_r + hc * 32, o);
                                tmem_wait_ld_on(o);
#pragma unroll
                                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                tc::tmem_st32(tO_r + hc * 32, o);
                            }
                            tc::tmem_wait_st();
                        }
                    }
                }
                float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                const float2 nm2 = make_float2(-m_ref, -m_ref);
                auto chunk = [&](uint32_t (&s)[32], int ch) {
                    mask_tail(s, ch);
                    uint32_t wv[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float2 x = make_float2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1]));
                        if (SAFE || !PRESCALED) x = ffma2(x, c2, nm2);   // fast pass with pre-scaled Q: S already is the exponent, reference 0
                        const bool poly = VSUM_A2_POLY_PERIOD > 0 && (e % (VSUM_A2_POLY_PERIOD > 0 ? VSUM_A2_POLY_PERIOD : 1)) == VSUM_A2_POLY_PERIOD - 1;
                        const float2 p = poly ? exp2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
                        ps[e & 3] = fadd2(ps[e & 3], p);
                        wv[e] = pack_bf16x2(p.x, p.y);
                    }
                    if (TRAIN && drop_thresh16 != 0) {            // one 64-bit draw decides 4 consecutive keys; the row sum stays un-dropped
                        const uint32_t th2 = drop_thresh16 * 0x00010001u;
#pragma unroll
                        for (int gq = 0; gq < 8; ++gq) {
                            const int key = j * BKV + ch * 32 + 4 * gq;
                            const unsigned long long z = dropout_bits64(seed, attn_drop_group_index(w.base + row, w.head, NH, key >> 2));
                            wv[2 * gq] &= __vcmpgeu2((uint32_t)z, th2);
                            wv[2 * gq + 1] &= __vcmpgeu2((uint32_t)(z >> 32), th2);
                        }
                    }
                    tmem_st16(tS_r + (uint32_t)(ch * 16), wv);    // in place: columns [16 ch, 16 ch + 16) belong to chunks already in registers
                };
                tc::tmem_ld32(tS_r, sa);
                tmem_wait_ld_on(sa);
                A2_TMARK(2);
                tc::tmem_ld32(tS_r + 32, sb);
                chunk(sa, 0);
                A2_TMARK(3);
                tmem_wait_ld_on(sb);
                tc::tmem_ld32(tS_r + 64, sa);
                chunk(sb, 1);
                A2_TMARK(4);
                tmem_wait_ld_on(sa);
                tc::tmem_ld32(tS_r + 96, sb);
                chunk(sa, 2);
                A2_TMARK(5);
                tmem_wait_ld_on(sb);
                const uint32_t p_slot = t * 3 + c3;
                ub += (uint32_t)n_q; if (ub >= 3) ub -= 3;       // next unit of this tile
                if (++c3 == 3) { c3 = 0; cpar ^= 1; }
                chunk(sb, 3);
                A2_TMARK(6);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                tc::mbar_arrive(p_full + p_slot);
                const float2 pq = fadd2(fadd2(ps[0], ps[1]), fadd2(ps[2], ps[3]));
                const float psum = pq.x + pq.y;
                if (!SAFE) danger |= !(psum < 1.2e27f);           // 2^90: an exponential near overflow (or a NaN): the exact pass redoes the item
                l_run += psum;
                A2_TMARK(7);
            }
#ifdef VSUM_A2_TIMING
            if (blockIdx.x == 0 && lane == 0 && w.nkv >= 16 && tcount++ == 1)
                printf("softmax warp %2d nkv %d | other %lld | wait S %lld | first ld %lld | chunk0 %lld | chunk1 %lld | chunk2 %lld | chunk3 %lld | "
                       "wait st + arrive %lld (clk per tile)\n", warp, w.nkv, tph[0] / w.nkv, tph[1] / w.nkv, tph[2] / w.nkv, tph[3] / w.nkv,
                       tph[4] / w.nkv, tph[5] / w.nkv, tph[6] / w.nkv, tph[7] / w.nkv);
            for (int i = 0; i < 10; ++i) tph[i] = 0;
            tmark = clock64();
#endif
            // The fast pass exponentiates against the fixed reference 0: valid while no exponential overflows (checked per tile
            // above) and the row as a whole does not underflow (l >= 2^-80, so its largest term is >= 2^-93).
            if (!SAFE && row < w.n && (danger || !(l_run >= 8.3e-25f)) && atomicExch(flags + idx, 1) == 0)
                atomicAdd(counters + 3, 1);                       // number of flagged items: the exact pass returns at once when it is 0
            // ---- epilogue: O_t / l -> global
            tc::mbar_wait(o_full + t, n_done & 1);                // the last PV product of the item has completed
            ++n_done;
            if (SAFE) ++pv_seen;                                  // ... which is also one more completion of pv_done
            tc::tc_fence_after();
            const float inv = (TRAIN ? keep_scale : 1.0f) / l_run;
            if (TRAIN && row < w.n) lse2[(int64_t)(w.base + row) * NH + w.head] = m_ref + log2f(l_run);
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {
                uint32_t o[32];
                tc::tmem_ld32(tO_r + hc * 32, o);
                tmem_wait_ld_on(o);
                if (row < w.n) {
                    if (TRAIN) {   // fp32 output: the backward's delta = rowsum(dO o O) must not see a rounded O
                        float *dst = reinterpret_cast<float *>(out_v) + (int64_t)(w.base + row) * DM + w.head * HD + hc * 32;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4 *>(dst + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                                                               __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    } else {
                        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(out_v) + (int64_t)(w.base + row) * DM + w.head * HD + hc * 32;
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
            tc::tc_fence_before();        // the O loads above are ordered before the first P of the next item, whose PV overwrites O_t
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (cw == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, A2_TMEM_COLS); }
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
        counters[3] = 0;
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
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
        VSUM_CUDA_OK(cudaFuncSetAttribute(attn2_tc05_kernel<false, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM2));
        VSUM_CUDA_OK(cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev)));
    {   // setmaxnreg moves registers inside the CTA's own allocation: what the control warps release must cover what the softmax warps acquire
        static std::atomic<int> checked{0};
        if (!checked.load(std::memory_order_relaxed)) {
            cudaFuncAttributes fa;
            VSUM_CUDA_OK(cudaFuncGetAttributes(&fa, attn2_tc05_kernel<false, false, true>));
            cudaFuncAttributes fb;
            VSUM_CUDA_OK(cudaFuncGetAttributes(&fb, attn2_tc05_kernel<true, false, false>));
            const int e = fa.numRegs < fb.numRegs ? fa.numRegs : fb.numRegs;
            VSUM_REQUIRE(128 * (e - A2_REGS_CONTROL) >= 256 * (A2_REGS_SOFTMAX - e), VSUM_EUNSUPPORTED,
                         "attn2_tc05_kernel was compiled with %d registers per thread: the softmax warps could not grow to %d", e, A2_REGS_SOFTMAX);
            cudaFuncAttributes fc;
            VSUM_CUDA_OK(cudaFuncGetAttributes(&fc, attn2_tc05_kernel<false, false, true, 2>));
            VSUM_REQUIRE(128 * (fc.numRegs - A2_REGS_CONTROL2) >= 512 * (A2_REGS_SOFTMAX2 - fc.numRegs), VSUM_EUNSUPPORTED,
                         "attn2_tc05_kernel<TPR 2> was compiled with %d registers per thread: the softmax warps could not grow to %d", fc.numRegs, A2_REGS_SOFTMAX2);
            checked.store(1, std::memory_order_relaxed);
        }
    }
    int grid = n_sm[dev & 63] - scorer_sm_reserve();
    if (grid < 1) grid = 1;
    const int max_items = max_blocks * NH;
    if (grid > max_items) grid = max_items;
    ProfScope prof(PROF_ATTN, s);
    float sl2 = scale * 1.4426950408889634f;
    const bool prescaled = fabsf(sl2 - 1.0f) < 1e-6f;     // the caller folded scale * log2(e) into Q (vsum_scorer.cu: bf16 weight copy)
    if (prescaled) sl2 = 1.0f;
#define A2_LAUNCH(TR, SF, PS, ...) attn2_tc05_kernel<TR, SF, PS><<<grid, A2_THREADS, A2_SMEM, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2, __VA_ARGS__)
    if (lse2) {
        const uint32_t thresh = attn_drop_thresh16(drop_p);
        const float ks = 65536.0f / (float)(65536u - thresh);
        A2_LAUNCH(true, false, false, lse2, ks, thresh, seed);
        VSUM_LAUNCH_OK("attn2_tc05_kernel");
        A2_LAUNCH(true, true, false, lse2, ks, thresh, seed);
    } else if (prescaled) {
        if (attention_kernel_version() == 3)    // fast pass with two threads per row (16 softmax warps)
            attn2_tc05_kernel<false, false, true, 2><<<grid, A2_THREADS2, A2_SMEM2, s>>>(tm, cu_seqlens, item_video, item_q0, counters, flags, out, sl2,
                                                                                         nullptr, 1.0f, 0u, 0ull);
        else
        A2_LAUNCH(false, false, true, nullptr, 1.0f, 0u, 0ull);
        VSUM_LAUNCH_OK("attn2_tc05_kernel");
        A2_LAUNCH(false, true, true, nullptr, 1.0f, 0u, 0ull);
    } else {
        A2_LAUNCH(false, false, false, nullptr, 1.0f, 0u, 0ull);
        VSUM_LAUNCH_OK("attn2_tc05_kernel");
        A2_LAUNCH(false, true, false, nullptr, 1.0f, 0u, 0ull);
    }
#undef A2_LAUNCH
    VSUM_LAUNCH_OK("attn2_tc05_kernel (exact pass)");
    return VSUM_OK;
}

}  // namespace vsum

extern "C" size_t vsum_attention_scratch_ints(int64_t T, int32_t B) {
    if (T < 0 || B < 0) return 0;
    const size_t a = vsum::attention2_scratch_ints(T, B), b = 2 * (size_t)(T / 128 + B) + 1;   // two-tile kernel / one-tile kernel
    return a > b ? a : b;
}
