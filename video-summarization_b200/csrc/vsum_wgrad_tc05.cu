// Weight gradient of a Linear layer on the tensor cores (training path):
//     dW[N,K] += dY[M,N]^T  X[M,K]         (operands rounded to bf16, fp32 accumulation)
// The contraction runs over the FRAME dimension M, so both operands are consumed exactly as they
// lie in memory: dY rows give the A operand M-major (its "M" is the output-feature index n), X rows
// give the B operand N-major.  TMA stages [64 frames x 64 features] boxes (128-byte swizzled rows);
// one tcgen05.mma (M=128, N=256, K=16 frames) consumes two 8-frame groups of 2 + 4 such boxes.
// (MN-major operands are a bf16/fp16 feature: with kind::tf32 the same descriptors return zeros on
// sm_100a -- measured, tests/test_tc05_gpu.py history -- so the fp32 activations are converted first.)
// Grid = (N/128) x (K/256) x splits over the frames; every CTA owns one fp32 accumulator tile in
// TMEM for its whole frame range and adds it to dW with fp32 atomics at the end.
#include "vsum_kernels.cuh"
#include "vsum_tc05.cuh"

namespace vsum {
namespace {

constexpr int WG_BM = 128;            // output-feature rows of dW per CTA (UMMA M)
constexpr int WG_BN = 256;            // input-feature columns of dW per CTA (UMMA N)
constexpr int WG_STAGES = 4;
constexpr int WG_A_STAGE = WG_BM * 128;               // 16 KB: (128 / E) atoms x R frames x 128 B with E * R = 1024
constexpr int WG_B_STAGE = WG_BN * 128;               // 32 KB
constexpr size_t WG_SMEM = (size_t)WG_STAGES * (WG_A_STAGE + WG_B_STAGE) + 256;
constexpr int WG_THREADS = 256;

// BF16 = true is the product path (bf16 operands, K = 16 frames per MMA).  BF16 = false (tf32, K = 8) is kept
// only as the record of the experiment: kind::tf32 ignores MN-major descriptors and it is never launched.
template <bool BF16>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc05_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                  float *__restrict__ dW, int K, int64_t M, int64_t rows_per_split) {
    constexpr int E = BF16 ? 64 : 32;          // features per 128-byte swizzle atom
    constexpr int WG_ROWS = BF16 ? 64 : 32;    // frames per pipeline stage
    constexpr int WG_BOX = WG_ROWS * 128;      // one [R frames x E features] box
    constexpr int KSTEP = BF16 ? 2048 : 1024;  // bytes of frames consumed by one MMA (16 / 8 rows of 128 B)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WG_STAGES * (WG_A_STAGE + WG_B_STAGE));
    uint64_t *full = bars, *empty = bars + WG_STAGES, *done = bars + 2 * WG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WG_BM, k0 = blockIdx.y * WG_BN;
    const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const int iters = (int)((m_end - m_begin + WG_ROWS - 1) / WG_ROWS);

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmY); tc::tma_prefetch_desc(&tmX); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, WG_BN); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer: 4 + 8 boxes per stage (rows past M are zero-filled) =====
            for (int it = 0; it < iters; ++it) {
                const int s = it % WG_STAGES;
                tc::mbar_wait(empty + s, ((it / WG_STAGES) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(full + s, WG_A_STAGE + WG_B_STAGE);
                uint8_t *a = smem + (size_t)s * (WG_A_STAGE + WG_B_STAGE), *b = a + WG_A_STAGE;
                const int row = (int)(m_begin + (int64_t)it * WG_ROWS);
                for (int j = 0; j < WG_BM / E; ++j) tc::tma_load_2d(a + j * WG_BOX, &tmY, full + s, n0 + j * E, row);
                for (int j = 0; j < WG_BN / E; ++j) tc::tma_load_2d(b + j * WG_BOX, &tmX, full + s, k0 + j * E, row);
            }
        }
    } else if (warp == 1) {   // ===== MMA issuer (warp-uniform, elected lane issues) =====
        constexpr uint32_t IDESC = tc::make_idesc(BF16 ? 1 : 2, WG_BM, WG_BN, 1, 1);   // A and B are MN-major
        // MN-major SW128: feature atoms are one box apart (LBO), 8-frame groups 1024 B apart (SBO)
        const uint64_t base_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem), WG_BOX, 1024);
        for (int it = 0; it < iters; ++it) {
            const int s = it % WG_STAGES;
            tc::mbar_wait(full + s, (it / WG_STAGES) & 1);
            tc::tc_fence_after();
            const uint32_t a_off = (uint32_t)(s * (WG_A_STAGE + WG_B_STAGE)) >> 4, b_off = a_off + (WG_A_STAGE >> 4);
            if (tc::elect_one()) {
#pragma unroll
                for (int g = 0; g < WG_BOX / KSTEP; ++g) {   // 4 MMAs per stage
                    const uint64_t ad = base_desc + (uint64_t)(a_off + g * (KSTEP >> 4)), bd = base_desc + (uint64_t)(b_off + g * (KSTEP >> 4));
                    if (BF16) tc::mma_f16_ss(tmem_base, ad, bd, IDESC, (it | g) != 0);
                    else tc::mma_tf32_ss(tmem_base, ad, bd, IDESC, (it | g) != 0);
                }
                tc::mma_commit(empty + s);
                if (it == iters - 1) tc::mma_commit(done);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {   // ===== epilogue: thread <-> one row (output feature) of the dW tile =====
        const int q = warp - 4, r = q * 32 + lane;
        if (iters > 0) {
            tc::mbar_wait(done, 0);
            tc::tc_fence_after();
            float *dst = dW + (int64_t)(n0 + r) * K + k0;
            uint32_t v[32];
#pragma unroll 1
            for (int c = 0; c < WG_BN / 32; ++c) {
                tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4)   // 16-byte vector reductions: a quarter of the atomic instructions
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + j), "f"(__uint_as_float(v[j])),
                                 "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                 : "memory");
            }
        }
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, WG_BN); }
}

// One launch prepares both wgrad operands.  blockIdx.z == 0: one pass over dY -- db[n] += sum_m dY[m, n] (when db != NULL) and
// the bf16 copy the wgrad MMAs consume; blockIdx.z == 1: the bf16 copy of X (grid-stride, four elements per thread).
__global__ void __launch_bounds__(256)
colsum_convert_kernel(const float *__restrict__ dY, __nv_bfloat16 *__restrict__ dY16, float *__restrict__ db, int64_t M, int N,
                      int64_t rows_per_block, const float *__restrict__ X, __nv_bfloat16 *__restrict__ X16, int64_t nx) {
    if (blockIdx.z == 1) {
        const int64_t stride = (int64_t)gridDim.x * gridDim.y * 256, t0 = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        const int64_t n4 = nx >> 2;                       // X and X16 are 16-byte aligned (workspace carve) and K is a multiple of 256
        for (int64_t k = t0; k < n4; k += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(X) + k);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&lo); pk.y = *reinterpret_cast<uint32_t *>(&hi);
            reinterpret_cast<uint2 *>(X16)[k] = pk;
        }
        for (int64_t k = 4 * n4 + t0; k < nx; k += stride) X16[k] = __float2bfloat16_rn(X[k]);
        return;
    }
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const int64_t m0 = (int64_t)blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
    float acc = 0.f;
    for (int64_t m = m0; m < m1; ++m) {
        const float v = dY[m * N + n];
        dY16[m * N + n] = __float2bfloat16_rn(v);
        acc += v;
    }
    if (db) atomicAdd(db + n, acc);
}

}  // namespace

template <bool BF16>
static int launch_wgrad(const void *dY, const void *X, float *dW, int64_t M, int N, int K, cudaStream_t s) {
    constexpr int E = BF16 ? 64 : 32, R = BF16 ? 64 : 32, ELT = BF16 ? 2 : 4;
    CUtensorMap tmY, tmX;
    int rc = make_tensor_map_2d(&tmY, dY, ELT, (uint64_t)N, (uint64_t)M, (uint64_t)N * ELT, E, R);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmX, X, ELT, (uint64_t)K, (uint64_t)M, (uint64_t)K * ELT, E, R);
    if (rc) return rc;
    auto kern = wgrad_tc05_kernel<BF16>;
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM)));
    const int tiles = (N / WG_BM) * (K / WG_BN);
    int splits = (int)max((int64_t)1, min((int64_t)(2 * 148 / tiles), ceil_div(M, 4 * R)));
    const int64_t rows = ceil_div(ceil_div(M, splits), R) * R;
    splits = (int)ceil_div(M, rows);
    dim3 grid(N / WG_BM, K / WG_BN, splits);
    ProfScope prof(PROF_OTHER, s);
    kern<<<grid, WG_THREADS, WG_SMEM, s>>>(tmY, tmX, dW, K, M, rows);
    VSUM_LAUNCH_OK("wgrad_tc05_kernel");
    return VSUM_OK;
}

// dW[N,K] += dY^T X, db[N] += colsum(dY) (db may be NULL).  dW / db must be zero-initialised.
// With bf16 scratch buffers (dY16 [M,N], X16 [M,K]) the operands are rounded to bf16 first.
int launch_linear_wgrad_tc05(const float *dY, const float *X, float *dW, float *db, int64_t M, int N, int K,
                             cudaStream_t s, __nv_bfloat16 *dY16, __nv_bfloat16 *X16) {
    VSUM_REQUIRE(N % WG_BM == 0 && K % WG_BN == 0, VSUM_EUNSUPPORTED, "wgrad_tc05: N=%d (x128) K=%d (x256)", N, K);
    if (M == 0) return VSUM_OK;
    int rc;
    VSUM_REQUIRE(dY16 && X16, VSUM_EINVAL, "wgrad_tc05: bf16 scratch buffers are required");
    {   // dY -> bf16 with its column sums, and X -> bf16, in one launch
        const int64_t rpb = ceil_div(M, 512);
        const bool x_vec = (((uintptr_t)X | (uintptr_t)X16) & 15) == 0;
        dim3 g2((unsigned)ceil_div(N, 256), (unsigned)ceil_div(M, rpb), x_vec ? 2 : 1);
        colsum_convert_kernel<<<g2, 256, 0, s>>>(dY, dY16, db, M, N, rpb, X, X16, M * (int64_t)K);
        VSUM_LAUNCH_OK("colsum_convert_kernel");
        if (!x_vec && (rc = launch_f32_to_bf16(X, X16, M * K, s))) return rc;
    }
    return launch_wgrad<true>(dY16, X16, dW, M, N, K, s);
}

}  // namespace vsum
