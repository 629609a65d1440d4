// Per-frame GEMMs of the scorer on the 5th-gen tensor cores (sm_100a):
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)
// A (activations, K contiguous) and W (nn.Linear weight [out,in], K contiguous) are both K-major,
// staged by TMA into 128B-swizzled shared memory, multiplied by tcgen05.mma (M=128, N=256,
// fp32 accumulators in TMEM, two accumulator stages), and drained by four epilogue warps with
// tcgen05.ld while the next tile's MMAs run.  Warp roles: 0 = TMA producer, 1 = MMA issuer,
// 2 = TMEM allocator, 4..7 = epilogue (one output row per thread, so LayerNorm needs no shuffles).
//
// Replaces (reference src/model/simnet.py): Embedding.feature_transform + positional add
// (211, 236-238) [tf32 MMA straight from the fp32 features], q/k/v Linears (148-153, one packed
// [768,256] weight), feature_projection + residual + norm1 (163, 107), fc1 + ReLU (181),
// fc2 + residual + norm2 (182, 110) and final_layer (+ sigmoid, train.py:144) fused into the
// last norm2 epilogue.
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {

// ------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int make_tensor_map_2d(CUtensorMap *out, const void *base, int elt_bytes, uint64_t inner, uint64_t rows,
                       uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    VSUM_REQUIRE(enc != nullptr, VSUM_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    VSUM_REQUIRE(box_inner * elt_bytes == 128, VSUM_EINVAL, "tensor map: box rows must be 128 bytes");
    VSUM_REQUIRE(((uintptr_t)base & 15) == 0 && (row_stride_bytes & 15) == 0, VSUM_EINVAL,
                 "tensor map: base and row stride must be 16-byte aligned");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VSUM_REQUIRE(r == CUDA_SUCCESS, VSUM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VSUM_OK;
}

namespace {

constexpr int BM = 128;                 // rows per tile (UMMA M)
constexpr int BN = 256;                 // columns per tile (UMMA N) -- one full d_model row
constexpr int STAGES = 4;
constexpr int A_STAGE = BM * 128;       // 16 KB: 128 rows x 128 B
constexpr int B_STAGE = BN * 128;       // 32 KB

constexpr int STG_BYTES = BM * 128;     // 16 KB: one [128 rows x 64 bf16] SW128 staging tile

constexpr int TMEM_COLS = 512;          // 2 accumulator stages x 256 fp32 columns
constexpr int WRES_MAX_KB = 4;          // weight-stationary variant: K = 256 bf16 -> 4 k-blocks = 128 KB
constexpr int EPI_BAR = 1;              // named barrier of 128 epilogue threads (EPI_BAR + half with eight epilogue warps)

// Every bf16-output epilogue runs on EIGHT epilogue warps: warps 4..7 drain columns
// [0,128) of the accumulator, warps 8..11 columns [128,256), each half with its own staging tile, named barrier and
// TMA-store leader (one warp per SM sub-partition cannot hide its own TMEM-load / store latency).  The operand ring
// keeps its four stages: taking shared memory from it for double-buffered staging per half (3 stages, or 2 A stages in
// the weight-stationary variant) made the K = 256 GEMMs 20-30 % SLOWER, so each half waits for its previous TMA
// store to finish reading the tile instead.  Measured on the same box: QKV GEMM 1.38 -> 1.09 ms per step; the
// LayerNorm epilogues (two threads per row, partial sums exchanged through shared memory) gain 4 %.
__host__ __device__ constexpr bool epi_uses_8_warps(int epi) { return true; }   // every epilogue (kept as a switch for experiments)
__host__ __device__ constexpr bool epi_is_ln(int epi) { return epi == TC_EPI_BIAS_RES_LN || epi == TC_EPI_BIAS_RES_LN_HEAD; }
__host__ __device__ constexpr int gemm_threads(int epi) { return epi_uses_8_warps(epi) ? 384 : 256; }
// DBL variant (out-projection + LayerNorm, K = 256): staging double-buffered per half (4 tiles) on a 3-stage ring that
// streams W from L2 instead of keeping it resident -- the epilogue chain (residual load -> statistics -> normalise ->
// store) is what bounds that tiny GEMM, 1.18 -> 0.96 ms per step.  With K = 1024 the 3-stage ring costs more than the
// staging returns (fc2 + LayerNorm +10 %, embedding +6 %), so those keep 4 stages and one staging tile per half.
__host__ __device__ constexpr int gemm_ring_stages(bool dbl) { return dbl ? 3 : STAGES; }
__host__ __device__ constexpr int gemm_stg_tiles(bool dbl) { return dbl ? 4 : 2; }
__host__ __device__ constexpr size_t gemm_ring_bytes(bool wres, bool dbl) {
    return wres ? (size_t)WRES_MAX_KB * B_STAGE + (size_t)gemm_ring_stages(dbl) * A_STAGE
                : (size_t)gemm_ring_stages(dbl) * (A_STAGE + B_STAGE);
}
__host__ __device__ constexpr size_t gemm_smem(int epi, bool wres, bool dbl) {   // + 2.5 KB: row-statistics / head-dot exchange of the LayerNorm epilogues
    return gemm_ring_bytes(wres, dbl) + (size_t)gemm_stg_tiles(dbl) * STG_BYTES + 384 /*barriers + scheduler ring*/ + (epi_is_ln(epi) ? 2560 : 0);
}

struct GemmParams {
    int64_t M;
    int N, num_kb, n_tiles;
    int64_t m_tiles;
    const float *bias;
    const float *gamma, *beta;
    const float *pos_table;
    const int32_t *row_pos;
    int pos_rows;
    const float *head_w, *head_b;
    float *scores_out, *feats_out;
    int apply_sigmoid;
    int store_out;
    int *sched;               // dynamic tile scheduler: sched[y] = next m-tile of column block y (WRES) / next tile, sched[8] = CTAs that have left
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

// The epilogue never touches global memory row-by-row (that costs one L1 wavefront per row and
// instruction): bf16 outputs are written to a 128B-swizzled staging tile and leave through TMA
// stores, the LayerNorm residual arrives through TMA loads into the same two staging tiles.
template <bool TF32, int EPI, bool WRES, bool DBL>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    // non-WRES: stage s = {A: s*48K, B: s*48K + 16K}.  WRES: W k-blocks at kb*32K, A ring after 128K.
    constexpr bool EPI8 = epi_uses_8_warps(EPI);
    static_assert(!DBL || (epi_uses_8_warps(EPI) && !WRES), "double-buffered staging: eight-warp, non weight-stationary only");
    constexpr int NST = gemm_ring_stages(DBL);                  // operand ring depth
    uint8_t *ring = smem;
    uint8_t *stg = smem + gemm_ring_bytes(WRES, DBL);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stg + (size_t)gemm_stg_tiles(DBL) * STG_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = tfull + 2;
    uint64_t *wfull = tempty + 2, *rfull = wfull + 1;     // rfull[4]: residual staging tiles (half, buffer)
    uint64_t *sched_full = rfull + 4, *sched_empty = sched_full + SCHED_RING;
    int32_t *sched_tile = reinterpret_cast<int32_t *>(sched_empty + SCHED_RING);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sched_tile + SCHED_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KELTS = TF32 ? 32 : 64;          // elements per 128-byte k-block
    constexpr uint32_t IDESC = tc::make_idesc(TF32 ? 2 : 1, BM, BN, 0, 0);
    constexpr bool IS_LN = EPI == TC_EPI_BIAS_RES_LN || EPI == TC_EPI_BIAS_RES_LN_HEAD;
    constexpr bool OUT_F32 = EPI >= TC_EPI_BIAS_F32;

    auto a_stage = [&](int s) -> uint8_t * {
        return WRES ? ring + (size_t)WRES_MAX_KB * B_STAGE + (size_t)s * A_STAGE : ring + (size_t)s * (A_STAGE + B_STAGE);
    };
    auto b_stage = [&](int s) -> uint8_t * { return ring + (size_t)s * (A_STAGE + B_STAGE) + A_STAGE; };

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        tc::tma_prefetch_desc(&tmOut);
        tc::tma_prefetch_desc(&tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(tfull + a, 1); tc::mbar_init(tempty + a, EPI8 ? 256 : 128); }
        for (int a = 0; a < 4; ++a) tc::mbar_init(rfull + a, 1);
        for (int a = 0; a < SCHED_RING; ++a) { tc::mbar_init(sched_full + a, 1); tc::mbar_init(sched_empty + a, 1 + (EPI8 ? 8 : 4)); }
        tc::mbar_init(wfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Dynamic tile scheduler: the producer takes tiles from a global counter (one per column block in the weight-stationary
    // variant) and hands them to the other roles through a small shared-memory ring, so a CTA that starts late -- the
    // evaluation stream holds some SMs while the next batch's scorer runs -- simply takes fewer tiles instead of leaving
    // its static share for a second wave.  -1 ends the walk.
    const int64_t total = WRES ? p.m_tiles : p.m_tiles * p.n_tiles;
    int *const counter = p.sched + (WRES ? blockIdx.y : 0);
    auto next_tile = [&](uint32_t n) -> int64_t {          // consumers: n-th tile of this CTA
        const int slot = n % SCHED_RING;
        tc::mbar_wait(sched_full + slot, (n / SCHED_RING) & 1);
        const int64_t t = sched_tile[slot];
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(sched_empty + slot);
        return t;
    };
    auto fetch_tile = [&](uint32_t n) -> int64_t {         // producer: take the next tile and publish it as the CTA's n-th
        const int slot = n % SCHED_RING;
        tc::mbar_wait(sched_empty + slot, ((n / SCHED_RING) & 1) ^ 1);
        // the CTA's first tile is its static one (no atomic round trip before the first loads), the rest come from the counter
        const int64_t t = n == 0 ? (int64_t)blockIdx.x : (int64_t)gridDim.x + atomicAdd(counter, 1);
        sched_tile[slot] = t < total ? (int32_t)t : -1;
        tc::mbar_arrive(sched_full + slot);
        return t < total ? t : -1;
    };

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            if (WRES) {
                tc::mbar_arrive_expect_tx(wfull, (uint32_t)p.num_kb * B_STAGE);
                for (int kb = 0; kb < p.num_kb; ++kb)
                    tc::tma_load_2d(ring + (size_t)kb * B_STAGE, &tmB, wfull, kb * KELTS, (int)blockIdx.y * BN);
            }
            uint32_t it = 0, nf = 0;
            int64_t t_next = fetch_tile(nf++);
            while (t_next >= 0) {
                const int64_t t = t_next;
                const int64_t m_blk = WRES ? t : t / p.n_tiles;
                const int n_blk = WRES ? (int)blockIdx.y : (int)(t % p.n_tiles);
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int s = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    tc::mbar_wait(empty + s, ph ^ 1);
                    tc::mbar_arrive_expect_tx(full + s, WRES ? A_STAGE : A_STAGE + B_STAGE);
                    tc::tma_load_2d(a_stage(s), &tmA, full + s, kb * KELTS, (int)(m_blk * BM));
                    if (!WRES) tc::tma_load_2d(b_stage(s), &tmB, full + s, kb * KELTS, n_blk * BN);
                    if (kb == 0) t_next = fetch_tile(nf++);   // one tile ahead, after this tile's first loads are on their way: the consumers find it in the ring, the L2 prefetch below uses it
                    // Weight-stationary variant (K = 256): the activation ring holds exactly one tile, so a tile's loads are
                    // issued one tile period ahead at best.  Pull the rows of this CTA's NEXT tile into L2 now (one of the
                    // column-block CTAs walking the same rows does it): QKV GEMM -4 %, fc1 -2.5 % on the same box.  Not for
                    // K = 1024: there the ring already streams 16 k-blocks per tile and the extra requests cost 15-30 %.
                    if (WRES && blockIdx.y == 0 && t_next >= 0) tc::tma_prefetch_l2_2d(&tmA, kb * KELTS, (int)(t_next * BM));
                }
            }
        }
    } else if (warp == 1) {  // ===== MMA issuer =====
        // Whole warp, warp-uniform control flow, one elected lane issues: the smem descriptors stay in
        // uniform registers (issuing under `lane == 0` costs ~160 cycles of R2UR traffic per MMA, more
        // than the 128 cycles an M128 N256 K16 MMA takes to execute).
        if (WRES) { tc::mbar_wait(wfull, 0); }
        const uint64_t ring_desc = tc::make_smem_desc_sw128(tc::smem_u32(ring), 16, 1024);
        uint32_t it = 0, tl = 0;
        for (;; ++tl) {
            if (next_tile(tl) < 0) break;
            const int acc = tl & 1;
            tc::mbar_wait(tempty + acc, ((tl >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                const int s = it % NST;
                tc::mbar_wait(full + s, (it / NST) & 1);
                tc::tc_fence_after();
                // descriptor start-address field is in 16-byte units
                const uint32_t a_off = (uint32_t)(a_stage(s) - ring) >> 4;
                const uint32_t b_off = (uint32_t)((WRES ? ring + (size_t)kb * B_STAGE : b_stage(s)) - ring) >> 4;
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 4 x (32 bytes of K) per 128-byte k-block
                        const uint64_t ad = ring_desc + (uint64_t)(a_off + k * 2);
                        const uint64_t bd = ring_desc + (uint64_t)(b_off + k * 2);
                        if (TF32) tc::mma_tf32_ss(d_tmem, ad, bd, IDESC, (kb | k) != 0);
                        else tc::mma_f16_ss(d_tmem, ad, bd, IDESC, (kb | k) != 0);
                    }
                    tc::mma_commit(empty + s);
                    if (kb == p.num_kb - 1) tc::mma_commit(tfull + acc);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {  // ===== epilogue: thread <-> one output row =====
        const int q = (warp - 4) & 3;                       // TMEM lane quarter
        const int half = EPI8 ? (warp - 4) >> 2 : 0;        // column half of the accumulator (eight-warp epilogues)
        const int r = q * 32 + lane;                        // row inside the tile
        const bool leader = r == 0;                         // one TMA-store leader per half
        const int epi_bar = EPI_BAR + half;
        // this row's eight 16-byte chunks inside a [128 x 64 bf16] SW128 staging tile (each half owns two tiles)
        constexpr int HALF_TILES = DBL ? 2 : 1;
        const uint32_t stg_row = tc::smem_u32(stg) + (uint32_t)(half * HALF_TILES) * STG_BYTES + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        uint8_t *const stg_half = stg + (size_t)(half * HALF_TILES) * STG_BYTES;
        uint32_t sw_off[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sw_off[ch] = (uint32_t)((ch ^ (r & 7)) << 4);
        uint32_t tl = 0, res_it = 0;
        for (;; ++tl) {
            const int64_t t = next_tile(tl);
            if (t < 0) break;
            const int64_t m_blk = WRES ? t : t / p.n_tiles;
            const int n_blk = WRES ? (int)blockIdx.y : (int)(t % p.n_tiles);
            const int acc = tl & 1;
            const int64_t row = m_blk * BM + r;
            const bool valid = row < p.M;
            const int n0 = n_blk * BN;
            if (IS_LN && leader) {   // my half's residual chunk(s) on their way while the MMAs still run
                tc::bulk_wait_read<0>();                     // my earlier stores no longer read the staging tiles
                for (int i = 0; i < HALF_TILES; ++i) {
                    tc::mbar_arrive_expect_tx(rfull + half * 2 + i, STG_BYTES);
                    tc::tma_load_2d(stg_half + i * STG_BYTES, &tmRes, rfull + half * 2 + i, (2 * half + i) * 64, (int)(m_blk * BM));
                }
            }
            tc::mbar_wait(tfull + acc, (tl >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            uint32_t ra[32], rb[32];

            if (OUT_F32) {   // fp32 output: 32 columns (= one 128-byte swizzled row of fp32) per pass
                const float *pos = nullptr;
                if (EPI == TC_EPI_BIAS_POS_F32 && valid)
                    pos = p.pos_table + (int64_t)min(__ldg(p.row_pos + row), p.pos_rows - 1) * BN;
#pragma unroll 1
                for (int cc = EPI8 ? 4 * half : 0; cc < (EPI8 ? 4 * half + 4 : 8); ++cc) {
                    tc::tmem_ld32(taddr + cc * 32, ra);
                    tc::tmem_wait_ld();
                    if (cc == (EPI8 ? 4 * half + 3 : 7)) { tc::tc_fence_before(); tc::mbar_arrive(tempty + acc); }
                    if (leader) { if (EPI8 && !DBL) tc::bulk_wait_read<0>(); else tc::bulk_wait_read<1>(); }
                    tc::bar_sync(epi_bar, 128);
                    const uint32_t dst = stg_row + ((EPI8 && !DBL) ? 0u : (uint32_t)(cc & 1) * STG_BYTES);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {         // 4 columns -> one 16-byte chunk
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + cc * 32 + ch * 4));
                        float v[4] = {__uint_as_float(ra[ch * 4 + 0]) + b4.x, __uint_as_float(ra[ch * 4 + 1]) + b4.y,
                                      __uint_as_float(ra[ch * 4 + 2]) + b4.z, __uint_as_float(ra[ch * 4 + 3]) + b4.w};
                        if (EPI == TC_EPI_BIAS_POS_F32 && valid) {
                            const float4 p4 = __ldg(reinterpret_cast<const float4 *>(pos + cc * 32 + ch * 4));
                            v[0] += p4.x; v[1] += p4.y; v[2] += p4.z; v[3] += p4.w;
                        }
                        if (EPI == TC_EPI_BIAS_RELU_F32) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
                        }
                        sts128(dst + sw_off[ch], __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                    }
                    tc::fence_proxy_async_smem();
                    tc::bar_sync(epi_bar, 128);
                    if (leader) {
                        tc::tma_store_2d(stg_half + ((EPI8 && !DBL) ? 0 : (cc & 1)) * STG_BYTES, &tmOut, n0 + cc * 32, (int)(m_blk * BM));
                        tc::bulk_commit();
                    }
                }
            } else if (!IS_LN) {
                const float *pos = nullptr;
                if (EPI == TC_EPI_BIAS_POS && valid)
                    pos = p.pos_table + (int64_t)min(__ldg(p.row_pos + row), p.pos_rows - 1) * BN;
#pragma unroll 1
                for (int cc = EPI8 ? 2 * half : 0; cc < (EPI8 ? 2 * half + 2 : 4); ++cc) {   // 64 output columns per pass
                    tc::tmem_ld32(taddr + cc * 64, ra);
                    tc::tmem_ld32(taddr + cc * 64 + 32, rb);
                    tc::tmem_wait_ld();
                    if (cc == (EPI8 ? 2 * half + 1 : 3)) { tc::tc_fence_before(); tc::mbar_arrive(tempty + acc); }   // my part of the accumulator is drained
                    if (leader) { if (EPI8 && !DBL) tc::bulk_wait_read<0>(); else tc::bulk_wait_read<1>(); }   // the store that last used this staging tile is done reading
                    tc::bar_sync(epi_bar, 128);
                    const uint32_t dst = stg_row + ((EPI8 && !DBL) ? 0u : (uint32_t)(cc & 1) * STG_BYTES);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {         // 8 columns -> one 16-byte chunk
                        const uint32_t *src = ch < 4 ? &ra[ch * 8] : &rb[(ch - 4) * 8];
                        const float *bp = p.bias + n0 + cc * 64 + ch * 8;
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bp)), b1 = __ldg(reinterpret_cast<const float4 *>(bp + 4));
                        float v[8] = {__uint_as_float(src[0]) + b0.x, __uint_as_float(src[1]) + b0.y, __uint_as_float(src[2]) + b0.z,
                                      __uint_as_float(src[3]) + b0.w, __uint_as_float(src[4]) + b1.x, __uint_as_float(src[5]) + b1.y,
                                      __uint_as_float(src[6]) + b1.z, __uint_as_float(src[7]) + b1.w};
                        if (EPI == TC_EPI_BIAS_POS && valid) {
                            const float4 p0 = __ldg(reinterpret_cast<const float4 *>(pos + cc * 64 + ch * 8));
                            const float4 p1 = __ldg(reinterpret_cast<const float4 *>(pos + cc * 64 + ch * 8 + 4));
                            v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
                            v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
                        }
                        if (EPI == TC_EPI_BIAS_RELU) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                        }
                        sts128(dst + sw_off[ch], pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                    tc::fence_proxy_async_smem();
                    tc::bar_sync(epi_bar, 128);
                    if (leader) {
                        tc::tma_store_2d(stg_half + ((EPI8 && !DBL) ? 0 : (cc & 1)) * STG_BYTES, &tmOut, n0 + cc * 64, (int)(m_blk * BM));
                        tc::bulk_commit();
                    }
                }
            } else {  // bias + residual + LayerNorm (+ head): N == 256, the tile is the whole row.  Two threads per row
                      // (columns [128 half, 128 half + 128) each) exchange their partial sums through shared memory.
                float *xch = reinterpret_cast<float *>(bars) + 96;          // [2 halves][128 rows][2], after the 384 B of barriers + scheduler ring
                float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    const int cc = 2 * half + i;
                    tc::tmem_ld32(taddr + cc * 64, ra);
                    tc::tmem_ld32(taddr + cc * 64 + 32, rb);
                    const int sb = DBL ? i : 0;               // staging buffer of this chunk
                    tc::mbar_wait(rfull + half * 2 + sb, DBL ? (tl & 1) : (res_it & 1));
                    ++res_it;
                    uint4 rs[8];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) rs[ch] = lds128(stg_row + sb * STG_BYTES + sw_off[ch]);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t *a = ch < 4 ? &ra[ch * 8] : &rb[(ch - 4) * 8];
                        const float *bp = p.bias + cc * 64 + ch * 8;
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bp)), b1 = __ldg(reinterpret_cast<const float4 *>(bp + 4));
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const uint32_t rw[4] = {rs[ch].x, rs[ch].y, rs[ch].z, rs[ch].w};
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&rw[e >> 1]);
                            const float v0 = __uint_as_float(a[e]) + bb[e] + __low2float(h);
                            const float v1 = __uint_as_float(a[e + 1]) + bb[e + 1] + __high2float(h);
                            sum += v0 + v1;
                            sumsq = fmaf(v0, v0, fmaf(v1, v1, sumsq));
                            a[e] = __float_as_uint(v0);
                            a[e + 1] = __float_as_uint(v1);
                        }
                    }
                    tc::tmem_st32(taddr + cc * 64, ra);       // park the pre-norm row in TMEM
                    tc::tmem_st32(taddr + cc * 64 + 32, rb);
                    tc::bar_sync(epi_bar, 128);               // my half has read this residual chunk
                    if (!DBL && leader && i == 0) {           // single staging tile: the second chunk can only be fetched now
                        tc::mbar_arrive_expect_tx(rfull + half * 2, STG_BYTES);
                        tc::tma_load_2d(stg_half, &tmRes, rfull + half * 2, (cc + 1) * 64, (int)(m_blk * BM));
                    }
                }
                tc::tmem_wait_st();
                xch[(half * 128 + r) * 2] = sum;
                xch[(half * 128 + r) * 2 + 1] = sumsq;
                tc::bar_sync(EPI_BAR + 2, 256);               // both halves of every row have published their sums
                sum += xch[((half ^ 1) * 128 + r) * 2];
                sumsq += xch[((half ^ 1) * 128 + r) * 2 + 1];
                tc::bar_sync(EPI_BAR + 2, 256);               // ... and everyone has read them (the next tile overwrites the slots)
                const float mean = sum * (1.0f / BN);
                const float var = fmaxf(sumsq * (1.0f / BN) - mean * mean, 0.f);
                const float rstd = rsqrtf(var + 1e-5f);
                float dot = 0.f;
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    const int cc = 2 * half + i;
                    tc::tmem_ld32(taddr + cc * 64, ra);
                    tc::tmem_ld32(taddr + cc * 64 + 32, rb);
                    tc::tmem_wait_ld();
                    if (i == 1) { tc::tc_fence_before(); tc::mbar_arrive(tempty + acc); }
                    const int sb = DBL ? i : 0;
                    if (p.store_out && !DBL) {                // (double-buffered: the tile was last read as residual, before a barrier)
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
                        if (EPI == TC_EPI_BIAS_RES_LN_HEAD) {
                            const float4 w0 = __ldg(reinterpret_cast<const float4 *>(p.head_w + c0)), w1 = __ldg(reinterpret_cast<const float4 *>(p.head_w + c0 + 4));
                            dot = fmaf(y[0], w0.x, fmaf(y[1], w0.y, fmaf(y[2], w0.z, fmaf(y[3], w0.w, dot))));
                            dot = fmaf(y[4], w1.x, fmaf(y[5], w1.y, fmaf(y[6], w1.z, fmaf(y[7], w1.w, dot))));
                            if (p.feats_out && valid) {
                                float4 *fo = reinterpret_cast<float4 *>(p.feats_out + row * BN + c0);
                                fo[0] = make_float4(y[0], y[1], y[2], y[3]);
                                fo[1] = make_float4(y[4], y[5], y[6], y[7]);
                            }
                        }
                        if (p.store_out)
                            sts128(stg_row + sb * STG_BYTES + sw_off[ch], pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                    }
                    if (p.store_out) {
                        tc::fence_proxy_async_smem();
                        tc::bar_sync(epi_bar, 128);
                        if (leader) {
                            tc::tma_store_2d(stg_half + sb * STG_BYTES, &tmOut, cc * 64, (int)(m_blk * BM));
                            tc::bulk_commit();
                        }
                    }
                }
                if (EPI == TC_EPI_BIAS_RES_LN_HEAD) {            // the two half-row dot products meet in half 0
                    if (half == 1) xch[512 + r] = dot;           // own 512-byte region after the sums
                    tc::bar_sync(EPI_BAR + 2, 256);
                    if (half == 0 && valid) {
                        float sc = dot + xch[512 + r] + __ldg(p.head_b);
                        if (p.apply_sigmoid) sc = 1.0f / (1.0f + __expf(-sc));
                        p.scores_out[row] = sc;
                    }
                }
            }
        }
        if (leader) tc::bulk_wait_all<0>();                  // all output tiles have landed
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
    // the last CTA to leave hands the scheduler slot back zeroed
    if (threadIdx.x == 0 && atomicAdd(p.sched + 8, 1) == (int)(gridDim.x * gridDim.y) - 1) {
        for (int i = 0; i < 9; ++i) p.sched[i] = 0;
        __threadfence();
    }
}

template <bool TF32, int EPI, bool WRES, bool DBL = false>
int launch_variant(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmOut, const CUtensorMap &tmRes,
                   const GemmParams &p, cudaStream_t s, int cat) {
    auto kern = gemm_tc05_kernel<TF32, EPI, WRES, DBL>;
    constexpr size_t SMEM = gemm_smem(EPI, WRES, DBL);
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM)));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms = max(8, sms - scorer_sm_reserve());               // SM partition with the evaluation stream
    dim3 grid;
    if (WRES) grid = dim3((unsigned)max((int64_t)1, min(p.m_tiles, (int64_t)(sms / p.n_tiles))), (unsigned)p.n_tiles);
    else grid = dim3((unsigned)min(p.m_tiles * p.n_tiles, (int64_t)sms));
    GemmParams q = p;
    q.sched = sched_slot();
    VSUM_REQUIRE(q.sched != nullptr, VSUM_ENOMEM, "gemm_tc05: no device memory for the tile scheduler");
    ProfScope prof(cat, s);
    kern<<<grid, gemm_threads(EPI), SMEM, s>>>(tmA, tmB, tmOut, tmRes, q);
    VSUM_LAUNCH_OK("gemm_tc05_kernel");
    return VSUM_OK;
}

}  // namespace

int launch_gemm_tc05(const Tc05GemmArgs &a, cudaStream_t s) {
    if (a.M == 0) return VSUM_OK;
    const int elt = a.a_is_f32 ? 4 : 2;
    const int kelts = 128 / elt;
    VSUM_REQUIRE(a.N % BN == 0 && a.K % kelts == 0, VSUM_EUNSUPPORTED,
                 "gemm_tc05: N=%d must be a multiple of 256 and K=%d of %d", a.N, a.K, kelts);
    VSUM_REQUIRE(a.M < ((int64_t)1 << 31), VSUM_EUNSUPPORTED, "gemm_tc05: M=%lld exceeds the TMA coordinate range", (long long)a.M);
    const bool full_row = a.epi == TC_EPI_BIAS_POS || a.epi == TC_EPI_BIAS_RES_LN || a.epi == TC_EPI_BIAS_RES_LN_HEAD ||
                          a.epi == TC_EPI_BIAS_POS_F32;
    const bool out_f32 = a.epi >= TC_EPI_BIAS_F32;
    VSUM_REQUIRE(!full_row || a.N == BN, VSUM_EUNSUPPORTED, "gemm_tc05: epilogue %d needs N == 256", a.epi);
    CUtensorMap tmA, tmB;
    int rc = make_tensor_map_2d(&tmA, a.A, elt, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.K * elt, kelts, BM);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmB, a.W, elt, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * elt, kelts, BN);
    if (rc) return rc;
    CUtensorMap tmOut = tmA, tmRes = tmA;                    // unused maps alias a valid one
    if (out_f32) {
        VSUM_REQUIRE(a.out_f32 && a.a_is_f32, VSUM_EINVAL, "gemm_tc05: fp32-output epilogues need out_f32 and fp32 (tf32) operands");
        rc = make_tensor_map_2d(&tmOut, a.out_f32, 4, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.N * 4, 32, BM);
        if (rc) return rc;
    } else if (a.out) {
        rc = make_tensor_map_2d(&tmOut, a.out, 2, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.N * 2, 64, BM);
        if (rc) return rc;
    }
    if (full_row && a.epi != TC_EPI_BIAS_POS && a.epi != TC_EPI_BIAS_POS_F32) {
        VSUM_REQUIRE(a.residual && a.gamma && a.beta, VSUM_EINVAL, "gemm_tc05: LayerNorm epilogue needs residual, gamma, beta");
        rc = make_tensor_map_2d(&tmRes, a.residual, 2, (uint64_t)BN, (uint64_t)a.M, (uint64_t)BN * 2, 64, BM);
        if (rc) return rc;
    }
    VSUM_REQUIRE(a.out || out_f32 || a.epi == TC_EPI_BIAS_RES_LN_HEAD, VSUM_EINVAL, "gemm_tc05: epilogue %d needs an output", a.epi);
    GemmParams p{};
    p.M = a.M; p.N = a.N; p.num_kb = a.K / kelts; p.n_tiles = a.N / BN; p.m_tiles = ceil_div(a.M, BM);
    p.bias = a.bias; p.gamma = a.gamma; p.beta = a.beta;
    p.pos_table = a.pos_table; p.row_pos = a.row_pos; p.pos_rows = a.pos_rows; p.head_w = a.head_w; p.head_b = a.head_b;
    p.scores_out = a.scores_out; p.feats_out = a.feats_out; p.apply_sigmoid = a.apply_sigmoid;
    p.store_out = a.out != nullptr;
    const bool wres = !a.a_is_f32 && p.num_kb <= WRES_MAX_KB && p.m_tiles >= 2 * (148 / p.n_tiles);
    if (a.a_is_f32) {
        switch (a.epi) {
            case TC_EPI_BIAS_POS: return launch_variant<true, TC_EPI_BIAS_POS, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
            case TC_EPI_BIAS: return launch_variant<true, TC_EPI_BIAS, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
            case TC_EPI_BIAS_F32: return launch_variant<true, TC_EPI_BIAS_F32, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
            case TC_EPI_BIAS_RELU_F32: return launch_variant<true, TC_EPI_BIAS_RELU_F32, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
            case TC_EPI_BIAS_POS_F32: return launch_variant<true, TC_EPI_BIAS_POS_F32, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
            default: return set_error(VSUM_EUNSUPPORTED, "gemm_tc05: the tf32 path has no epilogue %d", a.epi);
        }
    }
    switch (a.epi) {
        case TC_EPI_BIAS:
            return wres ? launch_variant<false, TC_EPI_BIAS, true>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat) : launch_variant<false, TC_EPI_BIAS, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
        case TC_EPI_BIAS_POS:     // feature GEMM from bf16 features (VSUM_MODE_BF16_FEATURES): K = 1024, W streamed
            return launch_variant<false, TC_EPI_BIAS_POS, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
        case TC_EPI_BIAS_RELU:
            return wres ? launch_variant<false, TC_EPI_BIAS_RELU, true>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat) : launch_variant<false, TC_EPI_BIAS_RELU, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
        case TC_EPI_BIAS_RES_LN:
            // K = 256 (out-projection): not weight-stationary -- the 128 KB a resident W block would take double-buffers the staging tiles
            return p.num_kb <= WRES_MAX_KB ? launch_variant<false, TC_EPI_BIAS_RES_LN, false, true>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat)
                                           : launch_variant<false, TC_EPI_BIAS_RES_LN, false, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
        case TC_EPI_BIAS_RES_LN_HEAD:
            return launch_variant<false, TC_EPI_BIAS_RES_LN_HEAD, false>(tmA, tmB, tmOut, tmRes, p, s, a.prof_cat);
        default:
            return set_error(VSUM_EINVAL, "gemm_tc05: unknown epilogue %d", a.epi);
    }
}

}  // namespace vsum
