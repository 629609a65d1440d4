// fp32 "reference mode" of the frame scorer (BASELINE.json: importance scores within 1e-5 of the
// reference's fp32 PyTorch path).  Plain SIMT kernels with fp32 FMA accumulation; they are the
// accuracy anchor for the bf16 tcgen05 path, not the throughput path.
// Reference: src/model/simnet.py (Embedding 208-217, MultiAttentionNetwork 138-164, MLP 180-183,
// EncoderBlock 105-114, final_layer 42).
#include "vsum_kernels.cuh"

#include <cfloat>

namespace vsum {

// row -> (video, position inside the video) for packed rows
__global__ void __launch_bounds__(256)
row_positions_kernel(const int32_t *__restrict__ cu, int B, int64_t T, int32_t *__restrict__ row_pos,
                     int32_t *__restrict__ row_vid) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= T) return;
    const int v = find_segment(cu, B, (int)m);
    row_pos[m] = (int)m - __ldg(cu + v);
    if (row_vid) row_vid[m] = v;
}

int launch_row_positions(const int32_t *cu_seqlens, int B, int64_t T, int32_t *row_pos,
                         int32_t *row_vid, cudaStream_t s) {
    if (T == 0) return VSUM_OK;
    row_positions_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, s>>>(cu_seqlens, B, T, row_pos, row_vid);
    VSUM_LAUNCH_OK("row_positions_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = epi(A[M,K] W[N,K]^T + bias): 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float *__restrict__ A, const float *__restrict__ W,
                  const float *__restrict__ bias, float *__restrict__ C, int64_t M, int N, int K,
                  const float *__restrict__ pos_table, const int32_t *__restrict__ row_pos,
                  int pos_rows) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Ws[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    const int lr = tid >> 2, lk = (tid & 3) * 4;          // loader: row lr, k offset lk
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
        if (m0 + lr < M) a = *reinterpret_cast<const float4 *>(A + (m0 + lr) * K + k0 + lk);
        if (n0 + lr < N) w = *reinterpret_cast<const float4 *>(W + (int64_t)(n0 + lr) * K + k0 + lk);
        As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Ws[lk + 0][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4 *>(&Ws[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int pr = 0;
        if (EPI == EPI_BIAS_POS) pr = min(__ldg(row_pos + m), pos_rows - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + __ldg(bias + n);
            if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.0f);
            if (EPI == EPI_BIAS_POS) v += __ldg(pos_table + (int64_t)pr * N + n);
            C[m * N + n] = v;
        }
    }
}

int launch_linear_f32(const float *A, const float *W, const float *bias, float *C, int64_t M, int N,
                      int K, int epi, const float *pos_table, const int32_t *row_pos, int pos_rows,
                      cudaStream_t s) {
    VSUM_REQUIRE(K % 16 == 0, VSUM_EUNSUPPORTED, "linear_f32: K=%d must be a multiple of 16", K);
    if (M == 0) return VSUM_OK;
    dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
    switch (epi) {
        case EPI_BIAS: linear_f32_kernel<EPI_BIAS><<<grid, 256, 0, s>>>(A, W, bias, C, M, N, K, nullptr, nullptr, 0); break;
        case EPI_BIAS_RELU: linear_f32_kernel<EPI_BIAS_RELU><<<grid, 256, 0, s>>>(A, W, bias, C, M, N, K, nullptr, nullptr, 0); break;
        case EPI_BIAS_POS: linear_f32_kernel<EPI_BIAS_POS><<<grid, 256, 0, s>>>(A, W, bias, C, M, N, K, pos_table, row_pos, pos_rows); break;
        default: return set_error(VSUM_EINVAL, "linear_f32: unknown epilogue %d", epi);
    }
    VSUM_LAUNCH_OK("linear_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// out = LayerNorm(a + res) * gamma + beta, one warp per row (d <= 1024, d % 32 == 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
add_layernorm_f32_kernel(const float *__restrict__ a, const float *__restrict__ res,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         float *__restrict__ out, int64_t M, int d) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float x[32];
    const int per = d >> 5;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) {
            const int c = i * 32 + lane;
            x[i] = a[m * d + c] + res[m * d + c];
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

int launch_add_layernorm_f32(const float *a, const float *res, const float *gamma, const float *beta,
                             float *out, int64_t M, int d, cudaStream_t s) {
    VSUM_REQUIRE(d % 32 == 0 && d <= 1024, VSUM_EUNSUPPORTED,
                 "add_layernorm_f32: d_model=%d must be a multiple of 32 and <= 1024", d);
    if (M == 0) return VSUM_OK;
    add_layernorm_f32_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, s>>>(a, res, gamma, beta, out, M, d);
    VSUM_LAUNCH_OK("add_layernorm_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 flash-style attention over packed videos: CTA = (64 queries, head, video),
// 64-key tiles, online softmax (simnet.py:155-161, scale = d_model^-0.5).
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256)
attention_f32_kernel(const float *__restrict__ qkv, const int32_t *__restrict__ cu, int d,
                     float scale, float *__restrict__ out, float *__restrict__ lse, float drop_p,
                     unsigned long long seed) {
    extern __shared__ __align__(16) float smem[];
    float *Qt = smem;                 // [HD][64]   Qt[k][r]
    float *Kt = Qt + HD * 64;         // [HD][64]   Kt[k][c]
    float *Vs = Kt + HD * 64;         // [64][HD]
    float *Ps = Vs + 64 * HD;         // [64][65]
    constexpr int OC = HD / 16;       // output columns per thread: tx + 16*j
    const int v = blockIdx.z, h = blockIdx.y;
    const int base = __ldg(cu + v), n = __ldg(cu + v + 1) - base;
    const int q0 = blockIdx.x * 64;
    if (q0 >= n) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ld = 3 * d;
    const float *qp = qkv + (int64_t)base * ld + h * HD;
    const float *kp = qp + d, *vp = qp + 2 * d;

    for (int idx = tid; idx < 64 * HD; idx += 256) {
        const int r = idx / HD, k = idx % HD;
        Qt[k * 64 + r] = (q0 + r < n) ? qp[(int64_t)(q0 + r) * ld + k] : 0.f;
    }
    const int H = gridDim.y;
    const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    float m_run[4], l_run[4], o[4][OC];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < OC; ++j) o[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < n; k0 += 64) {
        __syncthreads();
        for (int idx = tid; idx < 64 * HD; idx += 256) {
            const int c = idx / HD, k = idx % HD;
            const bool in = k0 + c < n;
            Kt[k * 64 + c] = in ? kp[(int64_t)(k0 + c) * ld + k] : 0.f;
            Vs[c * HD + k] = in ? vp[(int64_t)(k0 + c) * ld + k] : 0.f;
        }
        __syncthreads();
        float sacc[4][4] = {};
#pragma unroll 8
        for (int k = 0; k < HD; ++k) {
            const float4 qv = *reinterpret_cast<const float4 *>(Qt + k * 64 + ty * 4);
            const float4 kv = *reinterpret_cast<const float4 *>(Kt + k * 64 + tx * 4);
            const float qr[4] = {qv.x, qv.y, qv.z, qv.w}, kr[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sacc[i][j] = fmaf(qr[i], kr[j], sacc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sacc[i][j] = (k0 + tx * 4 + j < n) ? sacc[i][j] * scale : -INFINITY;
                mx = fmaxf(mx, sacc[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);            // finite: key k0 is always valid
            const float corr = expf(m_run[i] - m_new);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(sacc[i][j] - m_new);
                ps += p;                                          // the softmax normaliser ignores dropout
                float pd = p;
                if (drop_p > 0.f)                                 // simnet.py:159: dropout on the attention weights
                    pd = dropout_keep(seed, attn_drop_index(base + q0 + ty * 4 + i, h, H, k0 + tx * 4 + j), drop_p) ? p * keep_scale : 0.f;
                Ps[(ty * 4 + i) * 65 + tx * 4 + j] = pd;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            l_run[i] = l_run[i] * corr + ps;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < OC; ++j) o[i][j] *= corr;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < 64; ++c) {
            float vr[OC];
#pragma unroll
            for (int j = 0; j < OC; ++j) vr[j] = Vs[c * HD + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = Ps[(ty * 4 + i) * 65 + c];
#pragma unroll
                for (int j = 0; j < OC; ++j) o[i][j] = fmaf(p, vr[j], o[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= n) continue;
        const float inv = 1.0f / l_run[i];
#pragma unroll
        for (int j = 0; j < OC; ++j)
            out[(int64_t)(base + r) * d + h * HD + tx + 16 * j] = o[i][j] * inv;
        if (lse && tx == 0) lse[(int64_t)(base + r) * H + h] = m_run[i] + logf(l_run[i]);
    }
}

int launch_attention_f32(const float *qkv, const int32_t *cu_seqlens, int B, int max_len, int d,
                         int num_heads, float scale, float *out, cudaStream_t s, float *lse, float drop_p,
                         unsigned long long seed) {
    if (B == 0 || max_len == 0) return VSUM_OK;
    const int hd = d / num_heads;
    VSUM_REQUIRE(hd * num_heads == d, VSUM_EINVAL, "attention: d_model %d not divisible by %d heads", d, num_heads);
    VSUM_REQUIRE(B <= 65535 && num_heads <= 65535, VSUM_EUNSUPPORTED, "attention_f32: at most 65535 videos per call");
    dim3 grid((unsigned)ceil_div(max_len, 64), (unsigned)num_heads, (unsigned)B);
    const size_t smem = (size_t)(3 * 64 * hd + 64 * 65) * sizeof(float);
#define VSUM_ATT(HD)                                                                                   \
    {                                                                                                  \
        auto kern = attention_f32_kernel<HD>;                                                          \
        if (smem > 48 * 1024)                                                                          \
            VSUM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 256, smem, s>>>(qkv, cu_seqlens, d, scale, out, lse, drop_p, seed);               \
    }
    if (hd == 16) VSUM_ATT(16)
    else if (hd == 32) VSUM_ATT(32)
    else if (hd == 64) VSUM_ATT(64)
    else if (hd == 128) VSUM_ATT(128)
    else return set_error(VSUM_EUNSUPPORTED, "attention_f32: head_dim %d not in {16,32,64,128}", hd);
#undef VSUM_ATT
    VSUM_LAUNCH_OK("attention_f32_kernel");
    return VSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// final_layer (simnet.py:42) + optional sigmoid (train.py:144): one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_f32_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                float *__restrict__ scores, int64_t M, int d, int C, int apply_sigmoid) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
        for (int k = lane; k < d; k += 32) acc = fmaf(x[m * d + k], __ldg(w + (int64_t)c * d + k), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            float y = acc + __ldg(b + c);
            if (apply_sigmoid) y = 1.0f / (1.0f + expf(-y));
            scores[m * C + c] = y;
        }
    }
}

int launch_head_f32(const float *x, const float *w, const float *b, float *scores, int64_t M, int d,
                    int num_classes, int apply_sigmoid, cudaStream_t s) {
    if (M == 0) return VSUM_OK;
    head_f32_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, s>>>(x, w, b, scores, M, d, num_classes, apply_sigmoid);
    VSUM_LAUNCH_OK("head_f32_kernel");
    return VSUM_OK;
}

__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4 *>(in + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&a);
        pk.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(out + i) = pk;
    } else {
        for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
    }
}

// out = scale * in, as bf16 (out16) and / or as fp32 (out32); used once per weight load
__global__ void __launch_bounds__(256)
scale_convert_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out16, float *__restrict__ out32, int64_t n, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = in[i] * scale;
    if (out16) out16[i] = __float2bfloat16_rn(v);
    if (out32) out32[i] = v;
}

int launch_scale_convert(const float *in, __nv_bfloat16 *out16, float *out32, int64_t n, float scale, cudaStream_t s) {
    if (n == 0) return VSUM_OK;
    scale_convert_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(in, out16, out32, n, scale);
    VSUM_LAUNCH_OK("scale_convert_kernel");
    return VSUM_OK;
}

int launch_f32_to_bf16(const float *in, __nv_bfloat16 *out, int64_t n, cudaStream_t s) {
    if (n == 0) return VSUM_OK;
    f32_to_bf16_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(in, out, n);
    VSUM_LAUNCH_OK("f32_to_bf16_kernel");
    return VSUM_OK;
}

}  // namespace vsum
