// fp32 SIMT kernels of the scorer's training path: the exact (2e-4 vs autograd) mode of every op and the
// element-wise / LayerNorm / loss kernels all modes share (the tensor-core linears and attention live in
// vsum_gemm_tc05.cu, vsum_wgrad_tc05.cu, vsum_attn_tc05.cu, vsum_attn_bwd_tc05.cu).  Backward of everything under SimNet.forward (src/model/simnet.py:32-45) plus the
// masked MSE of src/utils/utils.py:45-56, as called by src/train.py:111-131.
// Dropout (simnet.py:107,110,159,181) is counter-based: masks are recomputed from (seed, index).
#include "vsum_kernels.cuh"

namespace vsum {

// ---------------------------------------------------------------------------------------------
// s = dropout(a) + res ; out = LayerNorm(s)          (EncoderBlock.forward, simnet.py:107,110)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
add_dropout_layernorm_f32_kernel(const float *__restrict__ a, const float *__restrict__ res,
                                 const float *__restrict__ gamma, const float *__restrict__ beta,
                                 float *__restrict__ s_out, float *__restrict__ out, int64_t M, int d,
                                 float drop_p, unsigned long long seed) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    float x[32];
    const int per = d >> 5;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) {
            const int c = i * 32 + lane;
            float av = a[m * d + c];
            if (drop_p > 0.f) av = dropout_keep(seed, (unsigned long long)(m * d + c), drop_p) ? av * ks : 0.f;
            x[i] = av + res[m * d + c];
            s_out[m * d + c] = x[i];
            sum += x[i];
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) { const float t = x[i] - mean; var = fmaf(t, t, var); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) {
            const int c = i * 32 + lane;
            out[m * d + c] = (x[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
}

// Same op for d = NV * 128: every lane owns NV groups of four consecutive columns -- 16-byte accesses and one
// dropout draw per group (the generic kernel above pays one draw per element and 4-byte accesses).
template <int NV>
__global__ void __launch_bounds__(256)
add_dropout_layernorm_v4_kernel(const float *__restrict__ a, const float *__restrict__ res,
                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                float *__restrict__ s_out, float *__restrict__ out, int64_t M, float drop_p,
                                unsigned long long seed) {
    constexpr int d = NV * 128;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const unsigned int th = dropout_thresh16(drop_p);
    float4 x[NV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int64_t e = m * d + j * 128 + lane * 4;
        float4 av = *reinterpret_cast<const float4 *>(a + e);
        if (drop_p > 0.f) {
            const unsigned long long z = dropout_bits64(seed, (unsigned long long)e >> 2);
            av.x = (unsigned int)(z & 0xffffu) >= th ? av.x * ks : 0.f;
            av.y = (unsigned int)((z >> 16) & 0xffffu) >= th ? av.y * ks : 0.f;
            av.z = (unsigned int)((z >> 32) & 0xffffu) >= th ? av.z * ks : 0.f;
            av.w = (unsigned int)(z >> 48) >= th ? av.w * ks : 0.f;
        }
        const float4 rv = *reinterpret_cast<const float4 *>(res + e);
        x[j] = make_float4(av.x + rv.x, av.y + rv.y, av.z + rv.z, av.w + rv.w);
        *reinterpret_cast<float4 *>(s_out + e) = x[j];
        sum += (x[j].x + x[j].y) + (x[j].z + x[j].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        x[j].x -= mean; x[j].y -= mean; x[j].z -= mean; x[j].w -= mean;
        var = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, fmaf(x[j].w, x[j].w, var))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)d + 1e-5f);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = j * 128 + lane * 4;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c)), bt = __ldg(reinterpret_cast<const float4 *>(beta + c));
        *reinterpret_cast<float4 *>(out + m * d + c) = make_float4(x[j].x * rstd * g.x + bt.x, x[j].y * rstd * g.y + bt.y,
                                                                   x[j].z * rstd * g.z + bt.z, x[j].w * rstd * g.w + bt.w);
    }
}

int launch_add_dropout_layernorm_f32(const float *a, const float *res, const float *gamma, const float *beta,
                                     float *s_out, float *out, int64_t M, int d, float drop_p,
                                     unsigned long long seed, cudaStream_t s) {
    VSUM_REQUIRE(d % 32 == 0 && d <= 1024, VSUM_EUNSUPPORTED, "add_dropout_layernorm_f32: d_model=%d", d);
    if (M == 0) return VSUM_OK;
    const bool aligned = ((((uintptr_t)a | (uintptr_t)res | (uintptr_t)s_out | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0);
    const unsigned blocks = (unsigned)ceil_div(M, 8);
    if (aligned && d == 256) add_dropout_layernorm_v4_kernel<2><<<blocks, 256, 0, s>>>(a, res, gamma, beta, s_out, out, M, drop_p, seed);
    else if (aligned && d == 128) add_dropout_layernorm_v4_kernel<1><<<blocks, 256, 0, s>>>(a, res, gamma, beta, s_out, out, M, drop_p, seed);
    else if (aligned && d == 512) add_dropout_layernorm_v4_kernel<4><<<blocks, 256, 0, s>>>(a, res, gamma, beta, s_out, out, M, drop_p, seed);
    else add_dropout_layernorm_f32_kernel<<<blocks, 256, 0, s>>>(a, res, gamma, beta, s_out, out, M, d, drop_p, seed);
    VSUM_LAUNCH_OK("add_dropout_layernorm_f32_kernel");
    return VSUM_OK;
}

__global__ void __launch_bounds__(256)
dropout_inplace_f32_kernel(float *__restrict__ x, int64_t n, float drop_p, unsigned long long seed) {
    const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one draw = four consecutive elements
    const int64_t i = i4 * 4;
    if (i >= n) return;
    const float ks = 1.0f / (1.0f - drop_p);
    const unsigned int th = dropout_thresh16(drop_p);
    const unsigned long long z = dropout_bits64(seed, (unsigned long long)i4);
    if (i + 3 < n && ((uintptr_t)(x + i) & 15) == 0) {
        float4 v = *reinterpret_cast<float4 *>(x + i);
        v.x = (unsigned int)(z & 0xffffu) >= th ? v.x * ks : 0.f;
        v.y = (unsigned int)((z >> 16) & 0xffffu) >= th ? v.y * ks : 0.f;
        v.z = (unsigned int)((z >> 32) & 0xffffu) >= th ? v.z * ks : 0.f;
        v.w = (unsigned int)(z >> 48) >= th ? v.w * ks : 0.f;
        *reinterpret_cast<float4 *>(x + i) = v;
    } else {
        for (int e = 0; e < 4 && i + e < n; ++e)
            x[i + e] = (unsigned int)((z >> (16 * e)) & 0xffffu) >= th ? x[i + e] * ks : 0.f;
    }
}

int launch_dropout_inplace_f32(float *x, int64_t n, float drop_p, unsigned long long seed, cudaStream_t s) {
    if (n == 0 || drop_p <= 0.f) return VSUM_OK;
    dropout_inplace_f32_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(x, n, drop_p, seed);
    VSUM_LAUNCH_OK("dropout_inplace_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (s - mean) * rstd * gamma + beta.  Each warp walks rows m, m + W, ...
// and keeps its share of dgamma / dbeta in registers until the end (one atomic per column and warp).
// ---------------------------------------------------------------------------------------------
// PER = d / 32 columns per lane (compile-time for the common widths so the four per-lane arrays live in
// exactly PER registers each).  dgamma / dbeta: per-warp register accumulation over the warp's rows, then
// one reduction over the block's warps through shared memory and ONE atomic per column and block -- with
// one atomic per column and WARP, 4736 warps hammered 512 addresses and the kernel ran at 10 % of HBM speed.
template <int PER, bool EXACT>
__global__ void __launch_bounds__(256)
layernorm_bwd_f32_kernel(const float *__restrict__ dy, const float *__restrict__ s_in,
                         const float *__restrict__ gamma, float *__restrict__ ds, float *__restrict__ d_a,
                         float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t M, int d, float drop_p,
                         unsigned long long seed) {
    __shared__ float red[2][8][PER * 32 > 256 ? 1 : PER * 32];      // [gamma|beta][warp][column] (PER <= 8 only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const int per = EXACT ? PER : (d >> 5);               // columns per lane actually present
    float g_acc[PER], b_acc[PER], gam[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) { g_acc[i] = 0.f; b_acc[i] = 0.f; gam[i] = i < per ? __ldg(gamma + i * 32 + lane) : 0.f; }
    for (int64_t m = warp0; m < M; m += nwarps) {
        float x[PER], g[PER];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { x[i] = i < per ? s_in[m * d + i * 32 + lane] : 0.f; sum += x[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)d;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { x[i] = i < per ? x[i] - mean : 0.f; var = fmaf(x[i], x[i], var); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var / (float)d + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const float dyv = i < per ? dy[m * d + i * 32 + lane] : 0.f;
            x[i] *= rstd;                                   // xhat
            g[i] = dyv * gam[i];                            // dxhat
            g_acc[i] = fmaf(dyv, x[i], g_acc[i]);
            b_acc[i] += dyv;
            s1 += g[i];
            s2 = fmaf(g[i], x[i], s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 /= (float)d; s2 /= (float)d;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (i >= per) continue;
            const int c = i * 32 + lane;
            const float v = rstd * (g[i] - s1 - x[i] * s2);
            ds[m * d + c] = v;
            if (d_a) d_a[m * d + c] = (drop_p > 0.f && !dropout_keep(seed, (unsigned long long)(m * d + c), drop_p)) ? 0.f : v * ks;
        }
    }
    if (PER <= 8) {
#pragma unroll
        for (int i = 0; i < PER; ++i) { red[0][warp][i * 32 + lane] = g_acc[i]; red[1][warp][i * 32 + lane] = b_acc[i]; }
        __syncthreads();
        for (int c = threadIdx.x; c < PER * 32; c += blockDim.x) {
            float gs = 0.f, bs = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { gs += red[0][w][c]; bs += red[1][w][c]; }
            atomicAdd(dgamma + c, gs);
            atomicAdd(dbeta + c, bs);
        }
    } else {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (i >= per) continue;
            atomicAdd(dgamma + i * 32 + lane, g_acc[i]);
            atomicAdd(dbeta + i * 32 + lane, b_acc[i]);
        }
    }
}

int launch_layernorm_bwd_f32(const float *dy, const float *s_in, const float *gamma, float *ds, float *d_a,
                             float *dgamma, float *dbeta, int64_t M, int d, float drop_p,
                             unsigned long long seed, cudaStream_t s) {
    VSUM_REQUIRE(d % 32 == 0 && d <= 1024, VSUM_EUNSUPPORTED, "layernorm_bwd_f32: d_model=%d", d);
    if (M == 0) return VSUM_OK;
    const unsigned blocks = (unsigned)max((int64_t)1, min(ceil_div(M, 8), (int64_t)(2 * 148)));
#define VSUM_LNB(P, E) layernorm_bwd_f32_kernel<P, E><<<blocks, 256, 0, s>>>(dy, s_in, gamma, ds, d_a, dgamma, dbeta, M, d, drop_p, seed)
    switch (d / 32) {
        case 2: VSUM_LNB(2, true); break;
        case 4: VSUM_LNB(4, true); break;
        case 8: VSUM_LNB(8, true); break;
        default: VSUM_LNB(32, false); break;              // any other width up to 1024: guarded generic instantiation
    }
#undef VSUM_LNB
    VSUM_LAUNCH_OK("layernorm_bwd_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// dW[N,K] += dY[M,N]^T X[M,K], db[N] += colsum(dY).  64x64 output tiles, split over M, fp32 atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_wgrad_f32_kernel(const float *__restrict__ dY, const float *__restrict__ X, float *__restrict__ dW,
                        float *__restrict__ db, int64_t M, int N, int K, int64_t rows_per_split) {
    __shared__ __align__(16) float Ys[16][64 + 4];
    __shared__ __align__(16) float Xs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    float acc[4][4] = {};
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t m0 = m_begin; m0 < m_end; m0 += 16) {
        float4 yv = make_float4(0.f, 0.f, 0.f, 0.f), xv = yv;
        if (m0 + lr < m_end) {
            yv = *reinterpret_cast<const float4 *>(dY + (m0 + lr) * N + n0 + lc);
            xv = *reinterpret_cast<const float4 *>(X + (m0 + lr) * K + k0 + lc);
        }
        *reinterpret_cast<float4 *>(&Ys[lr][lc]) = yv;
        *reinterpret_cast<float4 *>(&Xs[lr][lc]) = xv;
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&Ys[mm][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Xs[mm][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += ar[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dW + (int64_t)(n0 + ty * 4 + i) * K + k0 + tx * 4 + j, acc[i][j]);
        if (db && blockIdx.y == 0 && tx == 0) atomicAdd(db + n0 + ty * 4 + i, bsum[i]);
    }
}

int launch_linear_wgrad_f32(const float *dY, const float *X, float *dW, float *db, int64_t M, int N, int K,
                            cudaStream_t s) {
    VSUM_REQUIRE(N % 64 == 0 && K % 64 == 0, VSUM_EUNSUPPORTED, "linear_wgrad_f32: N=%d and K=%d must be multiples of 64", N, K);
    if (M == 0) return VSUM_OK;
    const int tiles = (N / 64) * (K / 64);
    int splits = (int)max((int64_t)1, min((int64_t)(2 * 148 * 2 / max(tiles, 1) + 1), ceil_div(M, 256)));
    int64_t rows = ceil_div(ceil_div(M, splits), 16) * 16;
    splits = (int)ceil_div(M, rows);
    dim3 grid(N / 64, K / 64, splits);
    linear_wgrad_f32_kernel<<<grid, 256, 0, s>>>(dY, X, dW, db, M, N, K, rows);
    VSUM_LAUNCH_OK("linear_wgrad_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// dX[M,K] (+)= dY[M,N] W[N,K]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_dgrad_f32_kernel(const float *__restrict__ dY, const float *__restrict__ W, float *__restrict__ dX,
                        int64_t M, int N, int K, int accumulate) {
    __shared__ __align__(16) float As[16][64 + 4];   // As[n][m]
    __shared__ __align__(16) float Bs[16][64 + 4];   // Bs[n][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 64;
    const int k0 = blockIdx.x * 64;
    const int ar_ = tid >> 2, an = (tid & 3) * 4;     // A loader: row ar_, n offset an
    const int br_ = tid >> 4, bk = (tid & 15) * 4;    // B loader: n row br_, k offset bk
    float acc[4][4] = {};
    for (int n0 = 0; n0 < N; n0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + ar_ < M) a = *reinterpret_cast<const float4 *>(dY + (m0 + ar_) * N + n0 + an);
        As[an + 0][ar_] = a.x; As[an + 1][ar_] = a.y; As[an + 2][ar_] = a.z; As[an + 3][ar_] = a.w;
        *reinterpret_cast<float4 *>(&Bs[br_][bk]) = *reinterpret_cast<const float4 *>(W + (int64_t)(n0 + br_) * K + k0 + bk);
        __syncthreads();
#pragma unroll
        for (int nn = 0; nn < 16; ++nn) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[nn][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[nn][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float *p = dX + m * K + k0 + tx * 4 + j;
            *p = accumulate ? *p + acc[i][j] : acc[i][j];
        }
    }
}

int launch_linear_dgrad_f32(const float *dY, const float *W, float *dX, int64_t M, int N, int K, int accumulate,
                            cudaStream_t s) {
    VSUM_REQUIRE(N % 16 == 0 && K % 64 == 0, VSUM_EUNSUPPORTED, "linear_dgrad_f32: N=%d (x16) K=%d (x64)", N, K);
    if (M == 0) return VSUM_OK;
    dim3 grid(K / 64, (unsigned)ceil_div(M, 64));
    linear_dgrad_f32_kernel<<<grid, 256, 0, s>>>(dY, W, dX, M, N, K, accumulate);
    VSUM_LAUNCH_OK("linear_dgrad_f32_kernel");
    return VSUM_OK;
}

__global__ void __launch_bounds__(256)
relu_dropout_bwd_f32_kernel(const float *__restrict__ hid, float *__restrict__ dhid, int64_t n, float ks) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 3 < n && (((uintptr_t)(hid + i) | (uintptr_t)(dhid + i)) & 15) == 0) {
        const float4 h = *reinterpret_cast<const float4 *>(hid + i);
        float4 g = *reinterpret_cast<float4 *>(dhid + i);
        g.x = h.x > 0.f ? g.x * ks : 0.f; g.y = h.y > 0.f ? g.y * ks : 0.f;
        g.z = h.z > 0.f ? g.z * ks : 0.f; g.w = h.w > 0.f ? g.w * ks : 0.f;
        *reinterpret_cast<float4 *>(dhid + i) = g;
    } else {
        for (int e = 0; e < 4 && i + e < n; ++e) dhid[i + e] = hid[i + e] > 0.f ? dhid[i + e] * ks : 0.f;
    }
}

int launch_relu_dropout_bwd_f32(const float *hid, float *dhid, int64_t n, float drop_p, cudaStream_t s) {
    if (n == 0) return VSUM_OK;
    relu_dropout_bwd_f32_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(hid, dhid, n, drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f);
    VSUM_LAUNCH_OK("relu_dropout_bwd_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// Attention backward (recompute).  Forward: P = softmax(S), Pd = dropout(P), O = Pd V.
//   delta = rowsum(dO * O);  dPd = dO V^T;  dS = P * (dropout_bwd(dPd) - delta) * scale
//   dQ = dS K;  dK = dS^T Q;  dV = Pd^T dO
// Two kernels so that no atomics are needed: one owns query tiles (dQ), one owns key tiles (dK, dV).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_delta_f32_kernel(const float *__restrict__ o, const float *__restrict__ d_o, float *__restrict__ delta,
                      int64_t T, int d, int H) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * H) return;
    const int64_t t = idx / H;
    const int h = (int)(idx % H), hd = d / H;
    float acc = 0.f;
    for (int c = 0; c < hd; ++c) acc = fmaf(o[t * d + h * hd + c], d_o[t * d + h * hd + c], acc);
    delta[idx] = acc;
}

template <int HD, bool KV_OWNER>
__global__ void __launch_bounds__(256)
attention_bwd_f32_kernel(const float *__restrict__ qkv, const float *__restrict__ d_o, const float *__restrict__ lse,
                         const float *__restrict__ delta, const int32_t *__restrict__ cu, int d, float scale,
                         float drop_p, unsigned long long seed, float *__restrict__ dqkv) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = HD + 1;
    float *Qs = smem, *Ks = Qs + 64 * LD, *Vs = Ks + 64 * LD, *Gs = Vs + 64 * LD;   // Gs = dO tile
    float *Ps = Gs + 64 * LD, *Ds = Ps + 64 * 65;                                    // Pd / dS tiles [64][65]
    constexpr int OC = HD / 16;
    const int v = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int base = __ldg(cu + v), n = __ldg(cu + v + 1) - base;
    const int own0 = blockIdx.x * 64;                       // first query (dQ kernel) or key (dK/dV kernel) of this CTA
    if (own0 >= n) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ld = 3 * d;
    const float *qp = qkv + (int64_t)base * ld + h * HD, *kp = qp + d, *vp = qp + 2 * d;
    const float *gp = d_o + (int64_t)base * d + h * HD;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;

    auto load_rows = [&](float *dst, const float *src, int row0, int stride) {
        for (int idx = tid; idx < 64 * HD; idx += 256) {
            const int r = idx / HD, c = idx % HD;
            dst[r * LD + c] = (row0 + r < n) ? src[(int64_t)(row0 + r) * stride + c] : 0.f;
        }
    };
    // the owned tile stays resident; the other side streams
    if (KV_OWNER) { load_rows(Ks, kp, own0, ld); load_rows(Vs, vp, own0, ld); }
    else { load_rows(Qs, qp, own0, ld); load_rows(Gs, gp, own0, d); }
    float acc1[4][OC], acc2[4][OC];                         // dQ | (dK, dV): rows ty*4+i, cols tx+16j
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < OC; ++j) { acc1[i][j] = 0.f; acc2[i][j] = 0.f; }

    for (int o0 = 0; o0 < n; o0 += 64) {
        __syncthreads();
        if (KV_OWNER) { load_rows(Qs, qp, o0, ld); load_rows(Gs, gp, o0, d); }
        else { load_rows(Ks, kp, o0, ld); load_rows(Vs, vp, o0, ld); }
        __syncthreads();
        const int q0 = KV_OWNER ? o0 : own0, k0 = KV_OWNER ? own0 : o0;
        // S[q][key] and dPd[q][key] for q = ty*4+i, key = tx*4+j
        float sacc[4][4] = {}, pacc[4][4] = {};
#pragma unroll 4
        for (int c = 0; c < HD; ++c) {
            float qr[4], gr[4], kr[4], vr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { qr[i] = Qs[(ty * 4 + i) * LD + c]; gr[i] = Gs[(ty * 4 + i) * LD + c]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { kr[j] = Ks[(tx * 4 + j) * LD + c]; vr[j] = Vs[(tx * 4 + j) * LD + c]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    sacc[i][j] = fmaf(qr[i], kr[j], sacc[i][j]);
                    pacc[i][j] = fmaf(gr[i], vr[j], pacc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + ty * 4 + i;
            const bool qv = q < n;
            const float l = qv ? __ldg(lse + (int64_t)(base + q) * H + h) : 0.f;
            const float dl = qv ? __ldg(delta + (int64_t)(base + q) * H + h) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = k0 + tx * 4 + j;
                float p = 0.f, pd = 0.f, dsv = 0.f;
                if (qv && key < n) {
                    p = expf(sacc[i][j] * scale - l);
                    bool keep = true;
                    if (drop_p > 0.f) keep = dropout_keep(seed, attn_drop_index(base + q, h, H, key), drop_p);
                    pd = keep ? p * ks : 0.f;
                    const float dp = keep ? pacc[i][j] * ks : 0.f;
                    dsv = p * (dp - dl) * scale;
                }
                Ps[(ty * 4 + i) * 65 + tx * 4 + j] = pd;
                Ds[(ty * 4 + i) * 65 + tx * 4 + j] = dsv;
            }
        }
        __syncthreads();
        if (KV_OWNER) {   // dK[key][c] += sum_q dS[q][key] Q[q][c];  dV[key][c] += sum_q Pd[q][key] dO[q][c]
#pragma unroll 4
            for (int q = 0; q < 64; ++q) {
                float qr[OC], gr[OC];
#pragma unroll
                for (int j = 0; j < OC; ++j) { qr[j] = Qs[q * LD + tx + 16 * j]; gr[j] = Gs[q * LD + tx + 16 * j]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dsv = Ds[q * 65 + ty * 4 + i], pd = Ps[q * 65 + ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < OC; ++j) { acc1[i][j] = fmaf(dsv, qr[j], acc1[i][j]); acc2[i][j] = fmaf(pd, gr[j], acc2[i][j]); }
                }
            }
        } else {          // dQ[q][c] += sum_key dS[q][key] K[key][c]
#pragma unroll 4
            for (int key = 0; key < 64; ++key) {
                float kr[OC];
#pragma unroll
                for (int j = 0; j < OC; ++j) kr[j] = Ks[key * LD + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dsv = Ds[(ty * 4 + i) * 65 + key];
#pragma unroll
                    for (int j = 0; j < OC; ++j) acc1[i][j] = fmaf(dsv, kr[j], acc1[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = own0 + ty * 4 + i;
        if (r >= n) continue;
        float *dst = dqkv + (int64_t)(base + r) * ld + h * HD;
#pragma unroll
        for (int j = 0; j < OC; ++j) {
            if (KV_OWNER) { dst[d + tx + 16 * j] = acc1[i][j]; dst[2 * d + tx + 16 * j] = acc2[i][j]; }
            else dst[tx + 16 * j] = acc1[i][j];
        }
    }
}

int launch_attention_bwd_f32(const float *qkv, const float *o, const float *d_o, const float *lse,
                             const int32_t *cu_seqlens, int B, int max_len, int64_t T, int d, int num_heads,
                             float scale, float drop_p, unsigned long long seed, float *delta_ws, float *dqkv,
                             cudaStream_t s) {
    if (B == 0 || T == 0) return VSUM_OK;
    const int hd = d / num_heads;
    VSUM_REQUIRE(hd * num_heads == d && (hd == 16 || hd == 32 || hd == 64), VSUM_EUNSUPPORTED,
                 "attention_bwd_f32: head_dim %d not in {16,32,64}", hd);
    VSUM_REQUIRE(B <= 65535, VSUM_EUNSUPPORTED, "attention_bwd_f32: at most 65535 videos per call");
    attn_delta_f32_kernel<<<(unsigned)ceil_div(T * num_heads, 256), 256, 0, s>>>(o, d_o, delta_ws, T, d, num_heads);
    VSUM_LAUNCH_OK("attn_delta_f32_kernel");
    dim3 grid((unsigned)ceil_div(max_len, 64), (unsigned)num_heads, (unsigned)B);
    const size_t smem = (size_t)(4 * 64 * (hd + 1) + 2 * 64 * 65) * sizeof(float);
#define VSUM_ATTB(HD, OWN)                                                                                  \
    {                                                                                                       \
        auto kern = attention_bwd_f32_kernel<HD, OWN>;                                                      \
        VSUM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        kern<<<grid, 256, smem, s>>>(qkv, d_o, lse, delta_ws, cu_seqlens, d, scale, drop_p, seed, dqkv);    \
        VSUM_LAUNCH_OK("attention_bwd_f32_kernel");                                                         \
    }
    if (hd == 16) { VSUM_ATTB(16, false) VSUM_ATTB(16, true) }
    else if (hd == 32) { VSUM_ATTB(32, false) VSUM_ATTB(32, true) }
    else { VSUM_ATTB(64, false) VSUM_ATTB(64, true) }
#undef VSUM_ATTB
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// final_layer backward (simnet.py:42): dx = d_scores W (+ d_feats); dW += d_scores^T x; db += colsum
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_bwd_f32_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ d_scores,
                    const float *__restrict__ d_feats, float *__restrict__ dx, float *__restrict__ dw,
                    float *__restrict__ db, int64_t M, int d, int C) {
    __shared__ float red[8][1024];
    __shared__ float red_b[8];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int c = 0; c < C; ++c) {
        float w_acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w_acc[i] = 0.f;
        float b_acc = 0.f;
        for (int64_t m = warp0; m < M; m += nwarps) {
            const float g = d_scores[m * C + c];
            b_acc += g;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int k = i * 32 + lane;
                if (k < d) {
                    w_acc[i] = fmaf(g, x[m * d + k], w_acc[i]);
                    const float add = g * __ldg(w + (int64_t)c * d + k);
                    if (c == 0) dx[m * d + k] = add + (d_feats ? d_feats[m * d + k] : 0.f);
                    else dx[m * d + k] += add;
                }
            }
        }
        // one atomic per column and BLOCK: the eight warps meet in shared memory first
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int k = i * 32 + lane;
            if (k < d) red[threadIdx.x >> 5][k] = w_acc[i];
        }
        if (lane == 0) red_b[threadIdx.x >> 5] = b_acc;
        __syncthreads();
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) t += red[w8][k];
            atomicAdd(dw + (int64_t)c * d + k, t);
        }
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w8 = 0; w8 < 8; ++w8) t += red_b[w8];
            atomicAdd(db + c, t);
        }
    }
}

int launch_head_bwd_f32(const float *x, const float *w, const float *d_scores, const float *d_feats, float *dx,
                        float *dw, float *db, int64_t M, int d, int C, cudaStream_t s) {
    VSUM_REQUIRE(d <= 1024, VSUM_EUNSUPPORTED, "head_bwd_f32: d_model=%d", d);
    if (M == 0) return VSUM_OK;
    const unsigned blocks = (unsigned)max((int64_t)1, min(ceil_div(M, 8), (int64_t)(2 * 148)));
    head_bwd_f32_kernel<<<blocks, 256, 0, s>>>(x, w, d_scores, d_feats, dx, dw, db, M, d, C);
    VSUM_LAUNCH_OK("head_bwd_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// masked MSE (src/utils/utils.py:45-56): padded entries contribute 0, the mean divides by `denom`
// (= bs * Nmax, the PADDED size).  d_out = grad_scale * 2 (out - tgt) keep / denom.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
masked_mse_f32_kernel(const float *__restrict__ out, const float *__restrict__ tgt, const uint8_t *__restrict__ pad,
                      int64_t n, float inv_denom, float *__restrict__ loss, float grad_scale, float *__restrict__ d_out) {
    __shared__ float red[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float keep = (pad && pad[i]) ? 0.f : 1.f;
        const float e = (out[i] - tgt[i]) * keep;              // utils.py:48-53: both sides are zeroed where padded
        acc = fmaf(e, e, acc);
        if (d_out) d_out[i] = grad_scale * 2.0f * e * inv_denom;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        acc = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
        if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * inv_denom);
    }
}

int launch_masked_mse_f32(const float *out, const float *tgt, const uint8_t *pad_mask, int64_t n, float denom,
                          float *loss, float grad_scale, float *d_out, cudaStream_t s) {
    if (n == 0) return VSUM_OK;
    const unsigned blocks = (unsigned)max((int64_t)1, min(ceil_div(n, 256), (int64_t)592));
    masked_mse_f32_kernel<<<blocks, 256, 0, s>>>(out, tgt, pad_mask, n, 1.0f / denom, loss, grad_scale, d_out);
    VSUM_LAUNCH_OK("masked_mse_f32_kernel");
    return VSUM_OK;
}

__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float *__restrict__ in, float *__restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (r0 + i < rows && c0 + threadIdx.x < cols) tile[i][threadIdx.x] = in[(int64_t)(r0 + i) * cols + c0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < cols && r0 + threadIdx.x < rows) out[(int64_t)(c0 + i) * rows + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

int launch_transpose_f32(const float *in, float *out, int rows, int cols, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
    VSUM_LAUNCH_OK("transpose_f32_kernel");
    return VSUM_OK;
}

__global__ void __launch_bounds__(256)
add_inplace_f32_kernel(float *__restrict__ dst, const float *__restrict__ src, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 a = *reinterpret_cast<float4 *>(dst + i);
        const float4 b = *reinterpret_cast<const float4 *>(src + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        *reinterpret_cast<float4 *>(dst + i) = a;
    } else {
        for (int64_t j = i; j < n; ++j) dst[j] += src[j];
    }
}

int launch_add_inplace_f32(float *dst, const float *src, int64_t n, cudaStream_t s) {
    if (n == 0) return VSUM_OK;
    add_inplace_f32_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(dst, src, n);
    VSUM_LAUNCH_OK("add_inplace_f32_kernel");
    return VSUM_OK;
}

}  // namespace vsum
