// Kernel temporal segmentation (change-point detection) on the GPU: the step that produces the shot
// boundaries the summariser consumes.  Follows src/data/preprocess/segmentations/kts/cpd_nonlin.py:5-91
// operation for operation, so that given the same kernel matrix K the objective values are bit-identical
// to the reference's numpy/Python run and the change points (first-minimum rule of line 76) are the same:
//   * K2 = cumsum(cumsum(K, 0), 1) in FLOAT32, sequential along each axis like numpy's cumsum (lines 13-16)
//   * K1 = fp64 running sum of diag(K) (line 13)
//   * scatters J[i,j] in fp64 with the reference's expression order (lines 20-22), stored transposed so the
//     dynamic programme reads it coalesced
//   * I[k,l] = min_t I[k-1,t] + J[t,l-1], strict '<' from 1e100 -> smallest t among equal minima (lines 70-79);
//     one CTA per l, one launch per k.  The reference spends O(m n^2) Python iterations here.
#include "vsum_kernels.cuh"

namespace vsum {
namespace {

// axis 0: thread per column, sequential down the rows (coalesced across threads)
__global__ void __launch_bounds__(256) kts_cumsum_axis0_kernel(const float *__restrict__ K, float *__restrict__ C, int n) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
        const float v = K[(int64_t)i * n + j];
        acc = i == 0 ? v : acc + v;
        C[(int64_t)i * n + j] = acc;
    }
}
// axis 1: CTA = 32 rows; 32x32 tiles go through shared memory so global accesses stay coalesced
__global__ void __launch_bounds__(1024) kts_cumsum_axis1_kernel(float *__restrict__ C, int n) {
    __shared__ float tile[32][33];
    __shared__ float carry[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, r = blockIdx.x * 32 + ty;
    for (int c0 = 0; c0 < n; c0 += 32) {
        tile[ty][tx] = (r < n && c0 + tx < n) ? C[(int64_t)r * n + c0 + tx] : 0.f;
        __syncthreads();
        if (ty == 0) {                                        // thread tx scans row tx of the tile
            float acc = c0 == 0 ? 0.f : carry[tx];
            for (int c = 0; c < 32; ++c) {
                acc = (c0 == 0 && c == 0) ? tile[tx][0] : acc + tile[tx][c];
                tile[tx][c] = acc;
            }
            carry[tx] = acc;
        }
        __syncthreads();
        if (r < n && c0 + tx < n) C[(int64_t)r * n + c0 + tx] = tile[ty][tx];
        __syncthreads();
    }
}
__global__ void kts_diag_cumsum_kernel(const float *__restrict__ K, double *__restrict__ K1, int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc = 0.0;                                         // np.cumsum([0] + list(diag)): fp64, sequential
    K1[0] = 0.0;
    for (int j = 0; j < n; ++j) { acc += (double)K[(int64_t)j * n + j]; K1[j + 1] = acc; }
}
// JT[j][i] = scatters[i][j] for i <= j
__global__ void __launch_bounds__(256)
kts_scatter_kernel(const float *__restrict__ C, const double *__restrict__ K1, double *__restrict__ JT, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
    if (i > j || i >= n) return;
    auto K2 = [&](int a, int b) -> double { return (a == 0 || b == 0) ? 0.0 : (double)C[(int64_t)(a - 1) * n + (b - 1)]; };
    const double q = (((K2(j + 1, j + 1) + K2(i, i)) - K2(j + 1, i)) - K2(i, j + 1)) / (double)(j - i + 1);
    JT[(int64_t)j * n + i] = (K1[j + 1] - K1[i]) - q;
}
__global__ void __launch_bounds__(256)
kts_dp_init_kernel(const double *__restrict__ JT, double *__restrict__ I, int32_t *__restrict__ P, int n, int m, int lmin, int lmax) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)(m + 1) * (n + 1)) return;
    const int k = (int)(idx / (n + 1)), l = (int)(idx % (n + 1));
    double v = 1e101;
    if (k == 0 && l >= lmin && l < lmax) v = JT[(int64_t)(l - 1) * n + 0];      // I[0, lmin:lmax] = J[0, lmin-1:lmax-1]
    I[idx] = v;
    P[idx] = 0;
}
// row k of the programme: CTA per l
__global__ void __launch_bounds__(128)
kts_dp_step_kernel(const double *__restrict__ JT, double *__restrict__ I, int32_t *__restrict__ P, int n, int k, int lmin, int lmax) {
    __shared__ double sv[4];
    __shared__ int st[4];
    const int l = (k + 1) * lmin + blockIdx.x;
    if (l > n) return;
    const int t0 = max(k * lmin, l - lmax), t1 = l - lmin;
    const double *prev = I + (int64_t)(k - 1) * (n + 1), *col = JT + (int64_t)(l - 1) * n;
    double best = 1e100;
    int arg = 0x7fffffff;
    for (int t = t0 + threadIdx.x; t <= t1; t += 128) {
        const double c = prev[t] + col[t];
        if (c < best) { best = c; arg = t; }                 // ascending t inside the thread: first minimum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; st[threadIdx.x >> 5] = arg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w)
            if (sv[w] < best || (sv[w] == best && st[w] < arg)) { best = sv[w]; arg = st[w]; }
        I[(int64_t)k * (n + 1) + l] = best;                   // stays 1e100 when nothing beat it (line 73)
        P[(int64_t)k * (n + 1) + l] = best < 1e100 ? arg : 0;
    }
}
__global__ void kts_scores_kernel(const double *__restrict__ I, double *__restrict__ scores, int n, int m) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= m) scores[k] = I[(int64_t)k * (n + 1) + n];
}

struct KtsWs { float *C; double *K1, *JT, *I; };
size_t carve_kts(int n, int m, void *base, KtsWs &w) {
    Carver k{(uint8_t *)base};
    w.C = k.get<float>((size_t)n * n); w.K1 = k.get<double>(n + 1); w.JT = k.get<double>((size_t)n * n);
    w.I = k.get<double>((size_t)(m + 1) * (n + 1));
    return align_up(k.off, 1024);
}

}  // namespace
}  // namespace vsum

using namespace vsum;

extern "C" size_t vsum_kts_workspace_bytes(int32_t n, int32_t m) {
    if (n <= 0 || m < 0) return 0;
    KtsWs w;
    return carve_kts(n, m, nullptr, w);
}

extern "C" int vsum_kts_gram(const float *features, int32_t n, int32_t dim, float *zeros_n, float *K_out, void *stream) {
    VSUM_REQUIRE(features && zeros_n && K_out && n > 0 && dim > 0, VSUM_EINVAL, "vsum_kts_gram: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    VSUM_CUDA_OK(cudaMemsetAsync(zeros_n, 0, (size_t)n * sizeof(float), s));
    return launch_linear_f32(features, features, zeros_n, K_out, n, n, dim, EPI_BIAS, nullptr, nullptr, 0, s);   // K = X X^T in fp32
}

extern "C" int vsum_kts_dp(const float *K, int32_t n, int32_t m, int32_t lmin, int32_t lmax, void *workspace, size_t workspace_bytes,
                           double *scores_out, int32_t *prev_out, void *stream) {
    VSUM_REQUIRE(K && workspace && scores_out && prev_out, VSUM_EINVAL, "vsum_kts_dp: null pointer");
    VSUM_REQUIRE(n > 0 && m >= 0 && lmin >= 1 && lmax >= lmin, VSUM_EINVAL, "vsum_kts_dp: bad sizes (n=%d m=%d lmin=%d lmax=%d)", n, m, lmin, lmax);
    VSUM_REQUIRE(n <= 46340, VSUM_EUNSUPPORTED, "vsum_kts_dp: n=%d exceeds 46340 frames", n);
    VSUM_REQUIRE(((uintptr_t)workspace & 1023) == 0 && workspace_bytes >= vsum_kts_workspace_bytes(n, m), VSUM_ENOMEM,
                 "vsum_kts_dp: workspace unaligned or too small");
    cudaStream_t s = (cudaStream_t)stream;
    KtsWs w;
    carve_kts(n, m, workspace, w);
    ProfScope prof(PROF_OTHER, s);
    kts_cumsum_axis0_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(K, w.C, n);
    VSUM_LAUNCH_OK("kts_cumsum_axis0_kernel");
    kts_cumsum_axis1_kernel<<<(unsigned)ceil_div(n, 32), 1024, 0, s>>>(w.C, n);
    VSUM_LAUNCH_OK("kts_cumsum_axis1_kernel");
    kts_diag_cumsum_kernel<<<1, 32, 0, s>>>(K, w.K1, n);
    VSUM_LAUNCH_OK("kts_diag_cumsum_kernel");
    kts_scatter_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)n), 256, 0, s>>>(w.C, w.K1, w.JT, n);
    VSUM_LAUNCH_OK("kts_scatter_kernel");
    kts_dp_init_kernel<<<(unsigned)ceil_div((int64_t)(m + 1) * (n + 1), 256), 256, 0, s>>>(w.JT, w.I, prev_out, n, m, lmin, lmax);
    VSUM_LAUNCH_OK("kts_dp_init_kernel");
    for (int k = 1; k <= m; ++k) {
        const int first = (k + 1) * lmin;
        if (first > n) break;
        kts_dp_step_kernel<<<(unsigned)(n - first + 1), 128, 0, s>>>(w.JT, w.I, prev_out, n, k, lmin, lmax);
        VSUM_LAUNCH_OK("kts_dp_step_kernel");
    }
    kts_scores_kernel<<<(unsigned)ceil_div(m + 1, 128), 128, 0, s>>>(w.I, scores_out, n, m);
    VSUM_LAUNCH_OK("kts_scores_kernel");
    return VSUM_OK;
}
