// Losses of the knowledge-distillation pretraining head (src/model/simnet_pretrain.py:35-100), forward
// and backward, on the packed (padding-free) layout:
//   repel  = mean_b (1/Nmax^2) sum_{i != j} xh_i . xh_j,  xh = x / (|x| + 1e-9)         (lines 49-69)
//          = mean_b (|sum_i xh_i|^2 - sum_i |xh_i|^2) / Nmax^2      -- no [N,N] similarity tensor
//   m      = softmax_i(score_i / t) over the video's frames                              (line 88)
//   center = mean_b (1/Nmax) sum_i (m_i + 1e-9) log(m_i + 1e-9)   or  mean_b |m_b|_2     (lines 90-94, 43-47)
//   loss   = mean_{b,c} -softmax(video_rep_b)_c log softmax(sum_i m_i x_i)_c             (lines 95-99, 35-41)
// x = video_transform(frame features) [T,512] fp32, Nmax = padded length of the reference's batch.
// All reductions run in a fixed order (no atomics): results are reproducible run to run.
#include "vsum_kernels.cuh"

namespace vsum {
namespace {

constexpr int PD = 512;          // video_transform width (simnet_pretrain.py:33)
constexpr int CHUNK = 256;       // frames per partial column sum

__device__ __forceinline__ float block_sum(float v, float *red) {   // blockDim.x multiple of 32, <= 1024
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];       // same order in every thread
    return t;
}
__device__ __forceinline__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
    return t;
}

// One block per video: mixture weights and the centering term.
__global__ void __launch_bounds__(256)
pt_mixture_kernel(const float *__restrict__ scores, const int32_t *__restrict__ cu, float inv_t, int pen_entropy,
                  float *__restrict__ mixture, float *__restrict__ center_b) {
    __shared__ float red[8];
    const int b = blockIdx.x, base = cu[b], n = cu[b + 1] - base;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, scores[base + i] * inv_t);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) sum += __expf(scores[base + i] * inv_t - mx);
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float m = __expf(scores[base + i] * inv_t - mx) * inv;
        mixture[base + i] = m;
        acc += pen_entropy ? (m + 1e-9f) * __logf(m + 1e-9f) : m * m;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) center_b[b] = pen_entropy ? acc : sqrtf(acc);
}

// One warp per frame: u = 1 / (|x| + 1e-9).
__global__ void __launch_bounds__(256)
pt_rownorm_kernel(const float *__restrict__ x, int64_t T, float *__restrict__ inv_norm) {
    const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= T) return;
    const float4 *p = reinterpret_cast<const float4 *>(x + row * PD);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < PD / 128; ++k) {
        const float4 v = p[k * 32 + lane];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) inv_norm[row] = 1.0f / (sqrtf(acc) + 1e-9f);
}

// Block (video b, chunk ch) of 512 threads = one column each: partial sums over <= 256 frames of
// xh (-> s_b), m * x (-> pooled_b); thread 0 also sums |xh_i|^2 = (|x_i| u_i)^2.
__global__ void __launch_bounds__(PD)
pt_colsum_kernel(const float *__restrict__ x, const float *__restrict__ inv_norm, const float *__restrict__ mixture,
                 const int32_t *__restrict__ cu, int chunks, float *__restrict__ partial) {
    __shared__ float su[CHUNK], sm[CHUNK];
    const int b = blockIdx.x, ch = blockIdx.y, base = cu[b], n = cu[b + 1] - base;
    const int r0 = ch * CHUNK, rows = min(CHUNK, n - r0);
    float *out = partial + ((int64_t)b * chunks + ch) * (2 * PD + 1);
    if (rows <= 0) {
        out[threadIdx.x] = 0.f; out[PD + threadIdx.x] = 0.f;
        if (threadIdx.x == 0) out[2 * PD] = 0.f;
        return;
    }
    if (threadIdx.x < rows) { su[threadIdx.x] = inv_norm[base + r0 + threadIdx.x]; sm[threadIdx.x] = mixture[base + r0 + threadIdx.x]; }
    __syncthreads();
    const float *p = x + (int64_t)(base + r0) * PD + threadIdx.x;
    float as = 0.f, ap = 0.f;
    for (int i = 0; i < rows; ++i) {
        const float v = p[(int64_t)i * PD];
        as = fmaf(v, su[i], as);
        ap = fmaf(v, sm[i], ap);
    }
    out[threadIdx.x] = as;
    out[PD + threadIdx.x] = ap;
    if (threadIdx.x == 0) {
        float q = 0.f;
        for (int i = 0; i < rows; ++i) { const float r = 1.0f / su[i] - 1e-9f; const float h = r * su[i]; q = fmaf(h, h, q); }
        out[2 * PD] = q;
    }
}

// One block of 512 threads per video: finish s_b, pooled_b; the distillation and repel terms.
__global__ void __launch_bounds__(PD)
pt_finalize_kernel(const float *__restrict__ partial, int chunks, const float *__restrict__ video_rep, int B, float inv_nmax2,
                   float *__restrict__ s_b, float *__restrict__ dpooled, float *__restrict__ loss_b, float *__restrict__ repel_b) {
    __shared__ float red[16];
    const int b = blockIdx.x, c = threadIdx.x;
    float s = 0.f, pooled = 0.f, q = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
        const float *in = partial + ((int64_t)b * chunks + ch) * (2 * PD + 1);
        s += in[c]; pooled += in[PD + c]; q += in[2 * PD];
    }
    s_b[(int64_t)b * PD + c] = s;
    const float ss = block_sum(s * s, red);
    // log-softmax of pooled, softmax of the target representation
    const float mx1 = block_max(pooled, red);
    const float z1 = block_sum(__expf(pooled - mx1), red);
    const float logp1 = pooled - mx1 - __logf(z1);
    const float vr = video_rep[(int64_t)b * PD + c];
    const float mx2 = block_max(vr, red);
    const float e2 = __expf(vr - mx2);
    const float z2 = block_sum(e2, red);
    const float p2 = e2 / z2;
    const float l = block_sum(-p2 * logp1, red);
    dpooled[(int64_t)b * PD + c] = (__expf(logp1) - p2) / (float)((int64_t)B * PD);   // d loss / d pooled
    if (c == 0) { loss_b[b] = l; repel_b[b] = (ss - q) * inv_nmax2; }
}

// losses[0..2] = (distillation, center, repel), reduced over the videos in index order.
__global__ void pt_reduce_kernel(const float *__restrict__ loss_b, const float *__restrict__ center_b,
                                 const float *__restrict__ repel_b, int B, float center_scale, float *__restrict__ losses) {
    if (threadIdx.x != 0) return;
    float a = 0.f, c = 0.f, r = 0.f;
    for (int b = 0; b < B; ++b) { a += loss_b[b]; c += center_b[b]; r += repel_b[b]; }
    losses[0] = a / (float)((int64_t)B * PD);
    losses[1] = c * center_scale;
    losses[2] = r / (float)B;
}

// ---- backward --------------------------------------------------------------------------------
// One warp per frame: d x (distillation through pooled + repel through the normalisation) and the
// gradient w.r.t. the frame's mixture weight.
__global__ void __launch_bounds__(256)
pt_bwd_rows_kernel(const float *__restrict__ x, const float *__restrict__ inv_norm, const float *__restrict__ mixture,
                   const float *__restrict__ s_b, const float *__restrict__ dpooled, const float *__restrict__ center_b,
                   const int32_t *__restrict__ cu, int B, int64_t T, const float *__restrict__ g, float repel_coef,
                   float center_coef, int pen_entropy, float *__restrict__ dx, float *__restrict__ dm) {
    const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= T) return;
    const int b = find_segment(cu, B, (int)row);
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    const float u = inv_norm[row], m = mixture[row];
    const float r = 1.0f / u - 1e-9f;
    const float4 *px = reinterpret_cast<const float4 *>(x + row * PD);
    const float4 *ps = reinterpret_cast<const float4 *>(s_b + (int64_t)b * PD);
    const float4 *pd = reinterpret_cast<const float4 *>(dpooled + (int64_t)b * PD);
    float4 xv[PD / 128], dh[PD / 128];
    float dot_pool = 0.f, dot_hx = 0.f;
    const float k2 = g2 * repel_coef;                       // 2 / (B Nmax^2)
#pragma unroll
    for (int k = 0; k < PD / 128; ++k) {
        const float4 v = px[k * 32 + lane], s = ps[k * 32 + lane], d = pd[k * 32 + lane];
        xv[k] = v;
        dot_pool += v.x * d.x + v.y * d.y + v.z * d.z + v.w * d.w;
        float4 h;                                            // d L / d xh = k2 (s_b - xh)
        h.x = k2 * (s.x - v.x * u); h.y = k2 * (s.y - v.y * u); h.z = k2 * (s.z - v.z * u); h.w = k2 * (s.w - v.w * u);
        dh[k] = h;
        dot_hx += h.x * v.x + h.y * v.y + h.z * v.z + h.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dot_pool += __shfl_xor_sync(0xffffffffu, dot_pool, o);
        dot_hx += __shfl_xor_sync(0xffffffffu, dot_hx, o);
    }
    // xh = x u(r): d x = u dh - x (dh . x) u^2 / r     (r = 0 <=> x = 0: second term vanishes)
    const float w = r > 0.f ? dot_hx * u * u / r : 0.f;
    const float gm = g0 * m;
    float4 *po = reinterpret_cast<float4 *>(dx + row * PD);
#pragma unroll
    for (int k = 0; k < PD / 128; ++k) {
        const float4 d = pd[k * 32 + lane];
        float4 o;
        o.x = u * dh[k].x - w * xv[k].x + gm * d.x; o.y = u * dh[k].y - w * xv[k].y + gm * d.y;
        o.z = u * dh[k].z - w * xv[k].z + gm * d.z; o.w = u * dh[k].w - w * xv[k].w + gm * d.w;
        po[k * 32 + lane] = o;
    }
    if (lane == 0) {
        const float dc = pen_entropy ? (__logf(m + 1e-9f) + 1.0f) : (center_b[b] > 0.f ? m / center_b[b] : 0.f);
        dm[row] = g0 * dot_pool + g1 * center_coef * dc;
    }
}

// One block per video: softmax backward, d score_i = m_i (dm_i - sum_j m_j dm_j) / t.
__global__ void __launch_bounds__(256)
pt_bwd_scores_kernel(const float *__restrict__ mixture, const float *__restrict__ dm, const int32_t *__restrict__ cu,
                     float inv_t, float *__restrict__ d_scores) {
    __shared__ float red[8];
    const int b = blockIdx.x, base = cu[b], n = cu[b + 1] - base;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) acc = fmaf(mixture[base + i], dm[base + i], acc);
    acc = block_sum(acc, red);
    for (int i = threadIdx.x; i < n; i += 256) d_scores[base + i] = mixture[base + i] * (dm[base + i] - acc) * inv_t;
}

struct Saved { float *mixture, *inv_norm, *dm, *partial, *s_b, *dpooled, *loss_b, *center_b, *repel_b; int chunks; };
size_t carve_saved(int64_t T, int B, int max_len, void *base, Saved &s) {
    Carver k{(uint8_t *)base};
    s.chunks = (int)ceil_div(max_len, CHUNK);
    s.mixture = k.get<float>(T); s.inv_norm = k.get<float>(T); s.dm = k.get<float>(T);
    s.partial = k.get<float>((size_t)B * s.chunks * (2 * PD + 1));
    s.s_b = k.get<float>((size_t)B * PD); s.dpooled = k.get<float>((size_t)B * PD);
    s.loss_b = k.get<float>(B); s.center_b = k.get<float>(B); s.repel_b = k.get<float>(B);
    return align_up(k.off, 1024);
}

}  // namespace
}  // namespace vsum

using namespace vsum;

extern "C" size_t vsum_pretrain_saved_bytes(int64_t T, int32_t B, int32_t max_len) {
    if (T <= 0 || B <= 0 || max_len <= 0) return 0;
    Saved s;
    return carve_saved(T, B, max_len, nullptr, s);
}

extern "C" int vsum_pretrain_losses_forward(const float *scores, const float *x512, const int32_t *cu_seqlens, int32_t B,
                                            int64_t T, int32_t max_len, int32_t n_pad, float sharpening_t,
                                            const float *video_rep, int32_t pen_entropy, float *losses3, void *saved,
                                            size_t saved_bytes, void *stream) {
    VSUM_REQUIRE(scores && x512 && cu_seqlens && video_rep && losses3 && saved, VSUM_EINVAL, "vsum_pretrain_losses_forward: null pointer");
    VSUM_REQUIRE(B > 0 && T > 0 && max_len > 0 && n_pad >= max_len && sharpening_t > 0.f, VSUM_EINVAL,
                 "vsum_pretrain_losses_forward: bad sizes (B=%d T=%lld max_len=%d n_pad=%d)", B, (long long)T, max_len, n_pad);
    VSUM_REQUIRE(B <= 65535 && ((uintptr_t)saved & 15) == 0 && ((uintptr_t)x512 & 15) == 0, VSUM_EINVAL,
                 "vsum_pretrain_losses_forward: at most 65535 videos; 16-byte aligned buffers");
    VSUM_REQUIRE(saved_bytes >= vsum_pretrain_saved_bytes(T, B, max_len), VSUM_ENOMEM, "vsum_pretrain_losses_forward: saved buffer too small");
    cudaStream_t s = (cudaStream_t)stream;
    Saved v;
    carve_saved(T, B, max_len, saved, v);
    pt_mixture_kernel<<<B, 256, 0, s>>>(scores, cu_seqlens, 1.0f / sharpening_t, pen_entropy, v.mixture, v.center_b);
    VSUM_LAUNCH_OK("pt_mixture_kernel");
    pt_rownorm_kernel<<<(unsigned)ceil_div(T * 32, 256), 256, 0, s>>>(x512, T, v.inv_norm);
    VSUM_LAUNCH_OK("pt_rownorm_kernel");
    pt_colsum_kernel<<<dim3(B, v.chunks), PD, 0, s>>>(x512, v.inv_norm, v.mixture, cu_seqlens, v.chunks, v.partial);
    VSUM_LAUNCH_OK("pt_colsum_kernel");
    pt_finalize_kernel<<<B, PD, 0, s>>>(v.partial, v.chunks, video_rep, B, 1.0f / ((float)n_pad * (float)n_pad), v.s_b, v.dpooled,
                                        v.loss_b, v.repel_b);
    VSUM_LAUNCH_OK("pt_finalize_kernel");
    const float center_scale = pen_entropy ? 1.0f / ((float)B * (float)n_pad) : 1.0f / (float)B;
    pt_reduce_kernel<<<1, 32, 0, s>>>(v.loss_b, v.center_b, v.repel_b, B, center_scale, losses3);
    VSUM_LAUNCH_OK("pt_reduce_kernel");
    return VSUM_OK;
}

extern "C" int vsum_pretrain_losses_backward(const float *x512, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                             int32_t max_len, int32_t n_pad, float sharpening_t, int32_t pen_entropy,
                                             const float *d_losses3, void *saved, float *d_scores, float *d_x512,
                                             void *stream) {
    VSUM_REQUIRE(x512 && cu_seqlens && d_losses3 && saved && d_scores && d_x512, VSUM_EINVAL, "vsum_pretrain_losses_backward: null pointer");
    VSUM_REQUIRE(B > 0 && T > 0 && max_len > 0 && n_pad >= max_len, VSUM_EINVAL, "vsum_pretrain_losses_backward: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    Saved v;
    carve_saved(T, B, max_len, saved, v);
    const float repel_coef = 2.0f / ((float)B * (float)n_pad * (float)n_pad);
    const float center_coef = pen_entropy ? 1.0f / ((float)B * (float)n_pad) : 1.0f / (float)B;
    pt_bwd_rows_kernel<<<(unsigned)ceil_div(T * 32, 256), 256, 0, s>>>(x512, v.inv_norm, v.mixture, v.s_b, v.dpooled, v.center_b,
                                                                       cu_seqlens, B, T, d_losses3, repel_coef, center_coef,
                                                                       pen_entropy, d_x512, v.dm);
    VSUM_LAUNCH_OK("pt_bwd_rows_kernel");
    pt_bwd_scores_kernel<<<B, 256, 0, s>>>(v.mixture, v.dm, cu_seqlens, 1.0f / sharpening_t, d_scores);
    VSUM_LAUNCH_OK("pt_bwd_scores_kernel");
    return VSUM_OK;
}
