// Scorer handle, weight packing and the forward pass (C ABI: vsum_scorer_*).
// Replaces SimNet.forward (src/model/simnet.py:32-45) over PACKED videos: rows of video v are
// [cu_seqlens[v], cu_seqlens[v+1]).  Attention never crosses a video boundary, which is exactly
// what the reference computes for batch size 1 (train.py:139-143) and, for padded batches, for
// every valid row (padded keys are masked, simnet.py:156-157).
#include "vsum_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <new>

using namespace vsum;

struct LayerOffsets {
    size_t wqkv, bqkv, wo, bo, ln1g, ln1b, fc1w, fc1b, fc2w, fc2b, ln2g, ln2b;   // fp32 blob (floats)
    size_t bqkv_s;                                                                  // fp32: bqkv with the q third pre-scaled (bf16 path)
    size_t h_wqkv, h_wo, h_fc1, h_fc2;                                              // bf16 blob (elements)
    size_t t_wqkv, t_wo, t_fc1, t_fc2;                                              // fp32 transposes [in,out] (floats)
};

struct vsum_scorer {
    vsum_scorer_config cfg;
    float *w32 = nullptr;
    __nv_bfloat16 *w16 = nullptr;
    float *pos_table = nullptr;
    int pos_rows = 0;
    size_t embed_w, embed_b, final_w, final_b, n32 = 0, n16 = 0;
    size_t h_embed = 0;            // bf16 copy of embed_w in the bf16 blob (VSUM_MODE_BF16_FEATURES)
    LayerOffsets L[VSUM_MAX_LAYERS];
    bool loaded = false;
    bool bf16_valid = false;       // the bf16 inference copies match w32 (false after a VSUM_WEIGHTS_TRAIN_ONLY refresh)
    const float *pos_src = nullptr;   // caller's table the handle's copy was taken from
    bool tc05_shape = false;
    int train_mode = 0;            // 0: fp32 SIMT linears; 1: tf32 tcgen05 linears (forward, dgrad, wgrad)
    float *zeros = nullptr;        // zero bias for the dgrad GEMMs
};

static size_t take(size_t &cursor, size_t n) {
    const size_t at = cursor;
    cursor += (n + 63) / 64 * 64;     // keep every tensor 256-byte aligned
    return at;
}

static constexpr float kQPrescale = 1.4426950408889634f;      // log2(e): folded, with d_model^-0.5, into the bf16 copy of W_q / b_q

extern "C" int vsum_scorer_create(vsum_scorer_t *out, const vsum_scorer_config *cfg) {
    VSUM_REQUIRE(out && cfg, VSUM_EINVAL, "vsum_scorer_create: null argument");
    VSUM_REQUIRE(cfg->d_model > 0 && cfg->num_heads > 0 && cfg->d_model % cfg->num_heads == 0, VSUM_EINVAL,
                 "vsum_scorer_create: d_model %d / heads %d", cfg->d_model, cfg->num_heads);
    VSUM_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= VSUM_MAX_LAYERS, VSUM_EINVAL,
                 "vsum_scorer_create: num_layers %d not in [1,%d]", cfg->num_layers, VSUM_MAX_LAYERS);
    VSUM_REQUIRE(cfg->d_model % 32 == 0 && cfg->d_model <= 1024 && cfg->d_ff % 16 == 0 && cfg->in_features % 16 == 0,
                 VSUM_EUNSUPPORTED, "vsum_scorer_create: d_model must be a multiple of 32 (<=1024), d_ff and in_features of 16");
    VSUM_REQUIRE(cfg->num_classes >= 1, VSUM_EINVAL, "vsum_scorer_create: num_classes %d", cfg->num_classes);
    int ndev = 0;
    VSUM_CUDA_OK(cudaGetDeviceCount(&ndev));
    VSUM_REQUIRE(ndev > 0, VSUM_ECUDA, "vsum_scorer_create: no CUDA device (there is no CPU fallback)");
    vsum_scorer *h = new (std::nothrow) vsum_scorer();
    VSUM_REQUIRE(h, VSUM_ENOMEM, "vsum_scorer_create: out of host memory");
    h->cfg = *cfg;
    const size_t d = cfg->d_model, ff = cfg->d_ff, in = cfg->in_features, C = cfg->num_classes;
    size_t c32 = 0, c16 = 0;
    h->embed_w = take(c32, d * in); h->embed_b = take(c32, d);
    h->final_w = take(c32, C * d); h->final_b = take(c32, C);
    for (int l = 0; l < cfg->num_layers; ++l) {
        LayerOffsets &o = h->L[l];
        o.wqkv = take(c32, 3 * d * d); o.bqkv = take(c32, 3 * d); o.bqkv_s = take(c32, 3 * d);
        o.wo = take(c32, d * d); o.bo = take(c32, d);
        o.ln1g = take(c32, d); o.ln1b = take(c32, d);
        o.fc1w = take(c32, ff * d); o.fc1b = take(c32, ff);
        o.fc2w = take(c32, d * ff); o.fc2b = take(c32, d);
        o.ln2g = take(c32, d); o.ln2b = take(c32, d);
        o.h_wqkv = take(c16, 3 * d * d); o.h_wo = take(c16, d * d);
        o.h_fc1 = take(c16, ff * d); o.h_fc2 = take(c16, d * ff);
        o.t_wqkv = take(c32, 3 * d * d); o.t_wo = take(c32, d * d);
        o.t_fc1 = take(c32, ff * d); o.t_fc2 = take(c32, d * ff);
    }
    h->h_embed = take(c16, d * in);
    h->n32 = c32; h->n16 = c16;
    h->tc05_shape = cfg->d_model == 256 && cfg->num_heads == 4 && cfg->d_ff == 1024 && cfg->num_classes == 1 &&
                    cfg->in_features % 32 == 0;
    cudaError_t e = cudaMalloc(&h->w32, c32 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->w16, c16 * sizeof(__nv_bfloat16));
    if (e == cudaSuccess) e = cudaMalloc(&h->zeros, 4096 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(h->zeros, 0, 4096 * sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(h->w32); cudaFree(h->w16); delete h;
        return set_error(VSUM_ENOMEM, "vsum_scorer_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return VSUM_OK;
}

extern "C" int vsum_scorer_destroy(vsum_scorer_t h) {
    if (!h) return VSUM_OK;
    cudaFree(h->w32); cudaFree(h->w16); cudaFree(h->pos_table); cudaFree(h->zeros);
    delete h;
    return VSUM_OK;
}

namespace {
// One launch copies a group of parameter tensors into the handle's fp32 blob (a training step refreshes 69 tensors: one
// cudaMemcpyAsync each made the weight refresh the largest block of launches of the launch-bound finetune step).
struct GatherArgs { const float *src[20]; float *dst[20]; float *tdst[20]; int n[20]; int cols[20]; int tstride[20]; int count; };
__global__ void __launch_bounds__(256) gather_copy_kernel(const GatherArgs a) {
    __shared__ float tile[32][33];
    const int i = blockIdx.y;
    if (i >= a.count) return;
    const float *__restrict__ src = a.src[i];
    float *__restrict__ dst = a.dst[i];
    const int n = a.n[i];
    if (a.tdst[i]) {   // weight matrix [rows, cols] (both multiples of 32): straight copy + [cols, rows] transpose for the dgrad GEMMs
        float *__restrict__ tdst = a.tdst[i];
        const int cols = a.cols[i], rows = n / cols, tc = cols >> 5, n_tiles = (rows >> 5) * tc, ts = a.tstride[i];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int r0 = (t / tc) << 5, c0 = (t % tc) << 5;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r0 + ty + 8 * k;
                const float v = __ldg(src + (size_t)r * cols + c0 + tx);
                dst[(size_t)r * cols + c0 + tx] = v;
                tile[ty + 8 * k][tx] = v;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) tdst[(size_t)(c0 + ty + 8 * k) * ts + r0 + tx] = tile[tx][ty + 8 * k];
            __syncthreads();
        }
        return;
    }
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const int n4 = n >> 2;
        for (int k = t0; k < n4; k += stride) reinterpret_cast<float4 *>(dst)[k] = __ldg(reinterpret_cast<const float4 *>(src) + k);
        for (int k = 4 * n4 + t0; k < n; k += stride) dst[k] = __ldg(src + k);
    } else {
        for (int k = t0; k < n; k += stride) dst[k] = __ldg(src + k);
    }
}
struct Gather {
    GatherArgs a{};
    // tdst != nullptr: also write the [cols, n / cols] transpose there (rows and cols must be multiples of 32, else tdst is ignored by the caller)
    int add(const float *src, float *dst, size_t n, const char *what, float *tdst = nullptr, int cols = 0, int tstride = 0) {
        VSUM_REQUIRE(src != nullptr, VSUM_EINVAL, "vsum_scorer_load_weights: null tensor %s", what);
        VSUM_REQUIRE(a.count < 20 && n < ((size_t)1 << 31), VSUM_EINVAL, "vsum_scorer_load_weights: gather group overflow");
        a.src[a.count] = src; a.dst[a.count] = dst; a.tdst[a.count] = tdst; a.n[a.count] = (int)n; a.cols[a.count] = cols; a.tstride[a.count] = tstride; ++a.count;
        return VSUM_OK;
    }
    int run(cudaStream_t s) {
        if (a.count == 0) return VSUM_OK;
        gather_copy_kernel<<<dim3(32, (unsigned)a.count), 256, 0, s>>>(a);
        VSUM_LAUNCH_OK("gather_copy_kernel");
        a.count = 0;
        return VSUM_OK;
    }
};
}  // namespace

extern "C" int vsum_scorer_load_weights_ex(vsum_scorer_t h, const vsum_scorer_weights *w, int32_t flags, void *stream) {
    VSUM_REQUIRE(h && w, VSUM_EINVAL, "vsum_scorer_load_weights: null argument");
    VSUM_REQUIRE((flags & ~VSUM_WEIGHTS_TRAIN_ONLY) == 0, VSUM_EINVAL, "vsum_scorer_load_weights_ex: unknown flags %d", flags);
    const bool train_only = (flags & VSUM_WEIGHTS_TRAIN_ONLY) != 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t d = h->cfg.d_model, ff = h->cfg.d_ff, in = h->cfg.in_features, C = h->cfg.num_classes;
    int rc;
    Gather g;
#define CP(dst_off, src, n) do { if ((rc = g.add((src), h->w32 + (dst_off), (n), #src))) return rc; } while (0)
    CP(h->embed_w, w->embed_w, d * in); CP(h->embed_b, w->embed_b, d);
    CP(h->final_w, w->final_w, C * d); CP(h->final_b, w->final_b, C);
    if ((rc = g.run(s))) return rc;
    if (h->cfg.use_pos) {
        VSUM_REQUIRE(w->pos_table && w->pos_rows > 0, VSUM_EINVAL, "vsum_scorer_load_weights: use_pos needs pos_table");
        const bool same_table = train_only && h->pos_src == w->pos_table && w->pos_rows == h->pos_rows;   // a training step does not touch the table
        if (w->pos_rows != h->pos_rows) {
            cudaFree(h->pos_table); h->pos_table = nullptr; h->pos_rows = 0;
            VSUM_CUDA_OK(cudaMalloc(&h->pos_table, (size_t)w->pos_rows * d * sizeof(float)));
            h->pos_rows = w->pos_rows;
        }
        if (!same_table)
            VSUM_CUDA_OK(cudaMemcpyAsync(h->pos_table, w->pos_table, (size_t)w->pos_rows * d * sizeof(float),
                                         cudaMemcpyDeviceToDevice, s));
        h->pos_src = w->pos_table;
    }
    if (!train_only && (rc = launch_f32_to_bf16(h->w32 + h->embed_w, h->w16 + h->h_embed, d * in, s))) return rc;
    for (int l = 0; l < h->cfg.num_layers; ++l) {
        const vsum_layer_weights &lw = w->layers[l];
        const LayerOffsets &o = h->L[l];
        // weight matrices go out straight and transposed ([out,in] -> [in,out]: the dgrad GEMMs dX = dY W run as dY (W^T)^T on the
        // K-major kernel) in the same launch when their sides are multiples of 32
        const bool fuse_t = d % 32 == 0 && ff % 32 == 0;
#define CPT(dst_off, src, rows, cols, t_off, tstride) do { if ((rc = g.add((src), h->w32 + (dst_off), (size_t)(rows) * (cols), #src, \
                                                                           fuse_t ? h->w32 + (t_off) : nullptr, (int)(cols), (int)(tstride)))) return rc; } while (0)
        CPT(o.wqkv, lw.q_w, d, d, o.t_wqkv, 3 * d); CPT(o.wqkv + d * d, lw.k_w, d, d, o.t_wqkv + d, 3 * d);
        CPT(o.wqkv + 2 * d * d, lw.v_w, d, d, o.t_wqkv + 2 * d, 3 * d);
        CP(o.bqkv, lw.q_b, d); CP(o.bqkv + d, lw.k_b, d); CP(o.bqkv + 2 * d, lw.v_b, d);
        CPT(o.wo, lw.o_w, d, d, o.t_wo, d); CP(o.bo, lw.o_b, d);
        CP(o.ln1g, lw.ln1_g, d); CP(o.ln1b, lw.ln1_b, d);
        CPT(o.fc1w, lw.fc1_w, ff, d, o.t_fc1, ff); CP(o.fc1b, lw.fc1_b, ff);
        CPT(o.fc2w, lw.fc2_w, d, ff, o.t_fc2, d); CP(o.fc2b, lw.fc2_b, d);
        CP(o.ln2g, lw.ln2_g, d); CP(o.ln2b, lw.ln2_b, d);
#undef CPT
        if ((rc = g.run(s))) return rc;
        if (!train_only) {
            // bf16 inference copy of [Wq; Wk; Wv]: the softmax scale d_model^-0.5 (simnet.py:126) and log2(e) are folded into the q
            // rows and the q bias, so that a score Q K^T already is the base-2 exponent the attention kernel needs
            const float qs = kQPrescale / sqrtf((float)d);
            if ((rc = launch_scale_convert(h->w32 + o.wqkv, h->w16 + o.h_wqkv, nullptr, d * d, qs, s))) return rc;
            if ((rc = launch_f32_to_bf16(h->w32 + o.wqkv + d * d, h->w16 + o.h_wqkv + d * d, 2 * d * d, s))) return rc;
            if ((rc = launch_scale_convert(h->w32 + o.bqkv, nullptr, h->w32 + o.bqkv_s, d, qs, s))) return rc;
            if ((rc = launch_scale_convert(h->w32 + o.bqkv + d, nullptr, h->w32 + o.bqkv_s + d, 2 * d, 1.0f, s))) return rc;
            if ((rc = launch_f32_to_bf16(h->w32 + o.wo, h->w16 + o.h_wo, d * d, s))) return rc;
            if ((rc = launch_f32_to_bf16(h->w32 + o.fc1w, h->w16 + o.h_fc1, ff * d, s))) return rc;
            if ((rc = launch_f32_to_bf16(h->w32 + o.fc2w, h->w16 + o.h_fc2, d * ff, s))) return rc;
        }
        if (!fuse_t) {
            if ((rc = launch_transpose_f32(h->w32 + o.wqkv, h->w32 + o.t_wqkv, (int)(3 * d), (int)d, s))) return rc;
            if ((rc = launch_transpose_f32(h->w32 + o.wo, h->w32 + o.t_wo, (int)d, (int)d, s))) return rc;
            if ((rc = launch_transpose_f32(h->w32 + o.fc1w, h->w32 + o.t_fc1, (int)ff, (int)d, s))) return rc;
            if ((rc = launch_transpose_f32(h->w32 + o.fc2w, h->w32 + o.t_fc2, (int)d, (int)ff, s))) return rc;
        }
    }
#undef CP
    h->loaded = true;
    h->bf16_valid = !train_only;
    return VSUM_OK;
}

extern "C" int vsum_scorer_load_weights(vsum_scorer_t h, const vsum_scorer_weights *w, void *stream) {
    return vsum_scorer_load_weights_ex(h, w, 0, stream);
}

namespace {

struct Ws32 { int32_t *row_pos; float *xa, *xb, *qkv, *att, *tmp, *hid; };
struct Ws16 { int32_t *row_pos, *tile_video, *tile_q0, *n_tiles, *a2; __nv_bfloat16 *xa, *xb, *qkv, *att, *hid; };

size_t carve32(const vsum_scorer_config &c, int64_t T, void *base, Ws32 &w) {
    Carver k{(uint8_t *)base};
    const size_t t = (size_t)T, d = c.d_model;
    w.row_pos = k.get<int32_t>(t);
    w.xa = k.get<float>(t * d); w.xb = k.get<float>(t * d); w.qkv = k.get<float>(t * 3 * d);
    w.att = k.get<float>(t * d); w.tmp = k.get<float>(t * d); w.hid = k.get<float>(t * c.d_ff);
    return align_up(k.off, 1024);
}
size_t carve16(const vsum_scorer_config &c, int64_t T, int max_tiles, void *base, Ws16 &w) {
    Carver k{(uint8_t *)base};
    const size_t t = (size_t)T, d = c.d_model;
    w.row_pos = k.get<int32_t>(t);
    w.tile_video = k.get<int32_t>(max_tiles); w.tile_q0 = k.get<int32_t>(max_tiles); w.n_tiles = k.get<int32_t>(1);
    w.a2 = k.get<int32_t>(6 * (size_t)max_tiles + 8);      // >= attention2_scratch_ints(T, B): work list + exact-pass flags of the two-tile attention kernel
    w.xa = k.get<__nv_bfloat16>(t * d); w.xb = k.get<__nv_bfloat16>(t * d);
    w.qkv = k.get<__nv_bfloat16>(t * 3 * d); w.att = k.get<__nv_bfloat16>(t * d);
    w.hid = k.get<__nv_bfloat16>(t * c.d_ff);
    return align_up(k.off, 1024);
}
int max_attn_tiles(int64_t T, int32_t B) { return (int)(T / 128 + B); }
}  // namespace

extern "C" size_t vsum_scorer_workspace_bytes(vsum_scorer_t h, int64_t T, int32_t B, int32_t mode) {
    if (!h || T <= 0 || B <= 0) return 0;
    if (mode == VSUM_MODE_FP32) { Ws32 w; return carve32(h->cfg, T, nullptr, w); }
    Ws16 w;
    return carve16(h->cfg, T, max_attn_tiles(T, B), nullptr, w);
}

static int forward_fp32(vsum_scorer_t h, const float *x, const int32_t *cu, int B, int64_t T, int max_len,
                        int sigm, float *scores, float *feats, void *ws, cudaStream_t s) {
    const vsum_scorer_config &c = h->cfg;
    const int d = c.d_model;
    Ws32 w;
    carve32(c, T, ws, w);
    int rc;
#define RUN(call) do { if ((rc = (call))) return rc; } while (0)
    RUN(launch_row_positions(cu, B, T, w.row_pos, nullptr, s));
    RUN(launch_linear_f32(x, h->w32 + h->embed_w, h->w32 + h->embed_b, w.xa, T, d, c.in_features,
                          c.use_pos ? EPI_BIAS_POS : EPI_BIAS, h->pos_table, w.row_pos, h->pos_rows, s));
    const float scale = 1.0f / sqrtf((float)d);                         // simnet.py:126: d_model ** -0.5
    for (int l = 0; l < c.num_layers; ++l) {
        const LayerOffsets &o = h->L[l];
        RUN(launch_linear_f32(w.xa, h->w32 + o.wqkv, h->w32 + o.bqkv, w.qkv, T, 3 * d, d, EPI_BIAS, nullptr, nullptr, 0, s));
        RUN(launch_attention_f32(w.qkv, cu, B, max_len, d, c.num_heads, scale, w.att, s));
        RUN(launch_linear_f32(w.att, h->w32 + o.wo, h->w32 + o.bo, w.tmp, T, d, d, EPI_BIAS, nullptr, nullptr, 0, s));
        RUN(launch_add_layernorm_f32(w.tmp, w.xa, h->w32 + o.ln1g, h->w32 + o.ln1b, w.xb, T, d, s));
        RUN(launch_linear_f32(w.xb, h->w32 + o.fc1w, h->w32 + o.fc1b, w.hid, T, c.d_ff, d, EPI_BIAS_RELU, nullptr, nullptr, 0, s));
        RUN(launch_linear_f32(w.hid, h->w32 + o.fc2w, h->w32 + o.fc2b, w.tmp, T, d, c.d_ff, EPI_BIAS, nullptr, nullptr, 0, s));
        RUN(launch_add_layernorm_f32(w.tmp, w.xb, h->w32 + o.ln2g, h->w32 + o.ln2b, w.xa, T, d, s));
    }
    RUN(launch_head_f32(w.xa, h->w32 + h->final_w, h->w32 + h->final_b, scores, T, d, c.num_classes, sigm, s));
    if (feats) VSUM_CUDA_OK(cudaMemcpyAsync(feats, w.xa, (size_t)T * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return VSUM_OK;
}

// x: fp32 features (tf32 feature GEMM) or, with x_is_bf16, bf16 features (bf16 feature GEMM on the bf16 copy of the weight)
static int forward_bf16(vsum_scorer_t h, const void *x, bool x_is_bf16, const int32_t *cu, int B, int64_t T, int sigm,
                        float *scores, float *feats, void *ws, cudaStream_t s) {
    const vsum_scorer_config &c = h->cfg;
    VSUM_REQUIRE(h->tc05_shape, VSUM_EUNSUPPORTED,
                 "the sm_100a tcgen05 scorer is built for d_model=256, heads=4, d_ff=1024, num_classes=1 "
                 "(got d_model=%d heads=%d d_ff=%d classes=%d); use VSUM_MODE_FP32",
                 c.d_model, c.num_heads, c.d_ff, c.num_classes);
    VSUM_REQUIRE(h->bf16_valid, VSUM_EINVAL, "vsum_scorer_forward: the weights were last refreshed with VSUM_WEIGHTS_TRAIN_ONLY; "
                 "call vsum_scorer_load_weights before the bf16 inference path");
    const int max_tiles = max_attn_tiles(T, B);
    Ws16 w;
    carve16(c, T, max_tiles, ws, w);
    int rc;
    RUN(launch_row_positions(cu, B, T, w.row_pos, nullptr, s));
    const bool attn2 = attention_kernel_version() >= 2;
    if (attn2) RUN(launch_attn2_schedule(cu, B, T, w.a2, s));
    else RUN(launch_attn_schedule(cu, B, w.tile_video, w.tile_q0, w.n_tiles, max_tiles, s));
    Tc05GemmArgs g{};
    g.A = x; g.M = T; g.N = 256; g.K = c.in_features; g.a_is_f32 = x_is_bf16 ? 0 : 1;
    g.W = x_is_bf16 ? (const void *)(h->w16 + h->h_embed) : (const void *)(h->w32 + h->embed_w);
    g.epi = c.use_pos ? TC_EPI_BIAS_POS : TC_EPI_BIAS; g.bias = h->w32 + h->embed_b; g.out = w.xa;
    g.pos_table = h->pos_table; g.row_pos = w.row_pos; g.pos_rows = h->pos_rows; g.prof_cat = PROF_EMBED;
    RUN(launch_gemm_tc05(g, s));
    const float scale = 1.0f / kQPrescale;                              // 256 ** -0.5 and log2(e) already sit in q: the kernels see scale * log2(e) = 1
    for (int l = 0; l < c.num_layers; ++l) {
        const LayerOffsets &o = h->L[l];
        const bool last = l == c.num_layers - 1;
        Tc05GemmArgs q{};
        q.A = w.xa; q.W = h->w16 + o.h_wqkv; q.M = T; q.N = 768; q.K = 256; q.epi = TC_EPI_BIAS;
        q.bias = h->w32 + o.bqkv_s; q.out = w.qkv; q.prof_cat = PROF_QKV;      // q arrives pre-scaled (vsum_scorer_load_weights)
        RUN(launch_gemm_tc05(q, s));
        if (attn2) RUN(launch_attention2_tc05(w.qkv, cu, B, T, scale, w.att, w.a2, s));
        else RUN(launch_attention_tc05(w.qkv, cu, w.tile_video, w.tile_q0, w.n_tiles, max_tiles, T, scale, w.att, s));
        Tc05GemmArgs p{};
        p.A = w.att; p.W = h->w16 + o.h_wo; p.M = T; p.N = 256; p.K = 256; p.epi = TC_EPI_BIAS_RES_LN;
        p.bias = h->w32 + o.bo; p.residual = w.xa; p.gamma = h->w32 + o.ln1g; p.beta = h->w32 + o.ln1b; p.out = w.xb;
        p.prof_cat = PROF_OPROJ_LN;
        RUN(launch_gemm_tc05(p, s));
        if (ffn_kernel_version() == 2) {   // fc1 + ReLU + fc2 + residual + LayerNorm (+ head) in one kernel: the hidden rows stay on the SM
            Tc05FfnArgs f{};
            f.x = w.xb; f.w1 = h->w16 + o.h_fc1; f.w2 = h->w16 + o.h_fc2; f.b1 = h->w32 + o.fc1b; f.b2 = h->w32 + o.fc2b;
            f.gamma = h->w32 + o.ln2g; f.beta = h->w32 + o.ln2b; f.M = T; f.out = last ? nullptr : w.xa;
            if (last) { f.head_w = h->w32 + h->final_w; f.head_b = h->w32 + h->final_b; f.scores_out = scores; f.feats_out = feats; }
            f.apply_sigmoid = sigm;
            RUN(launch_ffn_tc05(f, s));
            continue;
        }
        Tc05GemmArgs f1{};
        f1.A = w.xb; f1.W = h->w16 + o.h_fc1; f1.M = T; f1.N = 1024; f1.K = 256; f1.epi = TC_EPI_BIAS_RELU;
        f1.bias = h->w32 + o.fc1b; f1.out = w.hid; f1.prof_cat = PROF_FC1;
        RUN(launch_gemm_tc05(f1, s));
        Tc05GemmArgs f2{};
        f2.A = w.hid; f2.W = h->w16 + o.h_fc2; f2.M = T; f2.N = 256; f2.K = 1024;
        f2.epi = last ? TC_EPI_BIAS_RES_LN_HEAD : TC_EPI_BIAS_RES_LN;
        f2.bias = h->w32 + o.fc2b; f2.residual = w.xb; f2.gamma = h->w32 + o.ln2g; f2.beta = h->w32 + o.ln2b;
        f2.out = last ? nullptr : w.xa;
        f2.head_w = h->w32 + h->final_w; f2.head_b = h->w32 + h->final_b; f2.scores_out = scores; f2.feats_out = feats;
        f2.apply_sigmoid = sigm; f2.prof_cat = PROF_FC2_LN;
        RUN(launch_gemm_tc05(f2, s));
    }
#undef RUN
    return VSUM_OK;
}

extern "C" int vsum_scorer_forward(vsum_scorer_t h, const void *features, const int32_t *cu_seqlens,
                                   int32_t B, int64_t T, int32_t max_len, int32_t mode, int32_t apply_sigmoid,
                                   float *scores_out, float *feats_out, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    VSUM_REQUIRE(h, VSUM_EINVAL, "vsum_scorer_forward: null handle");
    VSUM_REQUIRE(h->loaded, VSUM_EINVAL, "vsum_scorer_forward: call vsum_scorer_load_weights first");
    VSUM_REQUIRE(B >= 0 && T >= 0 && max_len >= 0, VSUM_EINVAL, "vsum_scorer_forward: negative sizes");
    if (B == 0 || T == 0) return VSUM_OK;
    VSUM_REQUIRE(features && cu_seqlens && scores_out && workspace, VSUM_EINVAL, "vsum_scorer_forward: null pointer");
    VSUM_REQUIRE(mode == VSUM_MODE_FP32 || mode == VSUM_MODE_BF16 || mode == VSUM_MODE_BF16_FEATURES, VSUM_EINVAL,
                 "vsum_scorer_forward: unknown mode %d", mode);
    VSUM_REQUIRE(mode != VSUM_MODE_BF16_FEATURES || h->cfg.in_features % 64 == 0, VSUM_EUNSUPPORTED,
                 "vsum_scorer_forward: bf16 features need in_features to be a multiple of 64 (got %d)", h->cfg.in_features);
    VSUM_REQUIRE(T < ((int64_t)1 << 31), VSUM_EUNSUPPORTED, "vsum_scorer_forward: T=%lld frames exceed one call", (long long)T);
    VSUM_REQUIRE(!h->cfg.use_pos || max_len <= h->pos_rows, VSUM_EINVAL,
                 "vsum_scorer_forward: video of %d frames exceeds the positional table (%d rows); the reference "
                 "fails the same way past its 2000-row table (simnet.py:188)", max_len, h->pos_rows);
    const size_t need = vsum_scorer_workspace_bytes(h, T, B, mode);
    VSUM_REQUIRE(workspace_bytes >= need, VSUM_ENOMEM, "vsum_scorer_forward: workspace %zu < %zu bytes", workspace_bytes, need);
    VSUM_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)features & 15) == 0, VSUM_EINVAL,
                 "vsum_scorer_forward: workspace must be 1024-byte and features 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == VSUM_MODE_FP32)
        return forward_fp32(h, static_cast<const float *>(features), cu_seqlens, B, T, max_len, apply_sigmoid, scores_out, feats_out,
                            workspace, s);
    return forward_bf16(h, features, mode == VSUM_MODE_BF16_FEATURES, cu_seqlens, B, T, apply_sigmoid, scores_out, feats_out, workspace, s);
}

// ---- training (fp32) ---------------------------------------------------------------------------
namespace {
struct TapeLayer { float *qkv, *lse, *att, *s1, *xmid, *hid, *s2, *xout; };
struct Tape { float *x0; TapeLayer L[VSUM_MAX_LAYERS]; };
size_t carve_tape(const vsum_scorer_config &c, int64_t T, void *base, Tape &t) {
    Carver k{(uint8_t *)base};
    const size_t n = (size_t)T, d = c.d_model;
    t.x0 = k.get<float>(n * d);
    for (int l = 0; l < c.num_layers; ++l) {
        TapeLayer &L = t.L[l];
        L.qkv = k.get<float>(n * 3 * d); L.lse = k.get<float>(n * c.num_heads); L.att = k.get<float>(n * d);
        L.s1 = k.get<float>(n * d); L.xmid = k.get<float>(n * d); L.hid = k.get<float>(n * c.d_ff);
        L.s2 = k.get<float>(n * d); L.xout = k.get<float>(n * d);
    }
    return align_up(k.off, 1024);
}
struct TrainWs {
    int32_t *row_pos; float *a, *b, *c, *dd, *dhid, *dqkv, *delta, *dwqkv, *dbqkv; __nv_bfloat16 *y16, *x16;
    int32_t *tile_video, *tile_q0, *n_tiles; int max_tiles;       // tile list of the tcgen05 attention kernels
    int32_t *a2;                                                  // work list of the two-tile forward kernel
};
size_t carve_train_ws(const vsum_scorer_config &c, int64_t T, int32_t B, void *base, TrainWs &w) {
    Carver k{(uint8_t *)base};
    const size_t n = (size_t)T, d = c.d_model;
    w.row_pos = k.get<int32_t>(n);
    w.a = k.get<float>(n * d); w.b = k.get<float>(n * d); w.c = k.get<float>(n * d); w.dd = k.get<float>(n * d);
    w.dhid = k.get<float>(n * c.d_ff); w.dqkv = k.get<float>(n * 3 * d); w.delta = k.get<float>(n * c.num_heads);
    w.dwqkv = k.get<float>(3 * d * d); w.dbqkv = k.get<float>(3 * d);
    const size_t widest = std::max<size_t>(std::max<size_t>(c.d_ff, 3 * d), c.in_features);
    w.y16 = k.get<__nv_bfloat16>(n * widest); w.x16 = k.get<__nv_bfloat16>(n * widest);   // bf16 operands of the tcgen05 wgrad / attention
    w.max_tiles = (int)(T / 128 + B);
    w.tile_video = k.get<int32_t>(w.max_tiles); w.tile_q0 = k.get<int32_t>(w.max_tiles); w.n_tiles = k.get<int32_t>(1);
    w.a2 = k.get<int32_t>(6 * (size_t)w.max_tiles + 8);
    return align_up(k.off, 1024);
}
}  // namespace

extern "C" size_t vsum_scorer_tape_bytes(vsum_scorer_t h, int64_t T) {
    if (!h || T <= 0) return 0;
    Tape t;
    return carve_tape(h->cfg, T, nullptr, t);
}
extern "C" size_t vsum_scorer_train_workspace_bytes(vsum_scorer_t h, int64_t T, int32_t B) {
    if (!h || T <= 0 || B <= 0) return 0;
    TrainWs w;
    return carve_train_ws(h->cfg, T, B, nullptr, w);
}

// Linear layers of the training path: fp32 SIMT (mode 0) or tf32 tcgen05 (mode 1).
static int lin_fwd(vsum_scorer_t h, const float *A, const float *W, const float *bias, float *C, int64_t M, int N, int K,
                   int epi, const int32_t *row_pos, cudaStream_t s) {
    if (h->train_mode == 0 || N % 256 != 0 || K % 32 != 0)
        return launch_linear_f32(A, W, bias, C, M, N, K, epi, h->pos_table, row_pos, h->pos_rows, s);
    Tc05GemmArgs g{};
    g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.a_is_f32 = 1; g.bias = bias; g.out_f32 = C; g.prof_cat = PROF_OTHER;
    g.epi = epi == EPI_BIAS_RELU ? TC_EPI_BIAS_RELU_F32 : (epi == EPI_BIAS_POS ? TC_EPI_BIAS_POS_F32 : TC_EPI_BIAS_F32);
    g.pos_table = h->pos_table; g.row_pos = row_pos; g.pos_rows = h->pos_rows;
    return launch_gemm_tc05(g, s);
}
// dX[M,K] (+)= dY[M,N] W[N,K];  WT is W transposed ([K,N]); tmp [M,K] is scratch for the accumulating form
static int lin_dgrad(vsum_scorer_t h, const float *dY, const float *W, const float *WT, float *dX, float *tmp, int64_t M,
                     int N, int K, int accumulate, cudaStream_t s) {
    if (h->train_mode == 0 || K % 256 != 0 || N % 32 != 0 || K > 4096)
        return launch_linear_dgrad_f32(dY, W, dX, M, N, K, accumulate, s);
    Tc05GemmArgs g{};
    g.A = dY; g.W = WT; g.M = M; g.N = K; g.K = N; g.a_is_f32 = 1; g.bias = h->zeros; g.epi = TC_EPI_BIAS_F32;
    g.out_f32 = accumulate ? tmp : dX; g.prof_cat = PROF_OTHER;
    int rc = launch_gemm_tc05(g, s);
    if (rc || !accumulate) return rc;
    return launch_add_inplace_f32(dX, tmp, M * K, s);
}
static int lin_wgrad(vsum_scorer_t h, const float *dY, const float *X, float *dW, float *db, int64_t M, int N, int K,
                     __nv_bfloat16 *y16, __nv_bfloat16 *x16, cudaStream_t s) {
    if (h->train_mode == 0 || N % 128 != 0 || K % 256 != 0) return launch_linear_wgrad_f32(dY, X, dW, db, M, N, K, s);
    return launch_linear_wgrad_tc05(dY, X, dW, db, M, N, K, s, y16, x16);
}

extern "C" int vsum_scorer_set_train_mode(vsum_scorer_t h, int32_t mode) {
    VSUM_REQUIRE(h && mode >= 0 && mode <= 2, VSUM_EINVAL, "vsum_scorer_set_train_mode: mode %d", mode);
    VSUM_REQUIRE(mode < 2 || (h->cfg.d_model == 256 && h->cfg.num_heads == 4), VSUM_EUNSUPPORTED,
                 "vsum_scorer_set_train_mode: the tcgen05 attention needs d_model 256 with 4 heads");
    h->train_mode = mode;
    return VSUM_OK;
}

#define RUN(call) do { if ((rc = (call))) return rc; } while (0)
extern "C" int vsum_scorer_forward_train(vsum_scorer_t h, const float *x, const int32_t *cu, int32_t B, int64_t T,
                                         int32_t max_len, float p, uint64_t seed, float *scores, float *feats,
                                         void *tape_mem, size_t tape_bytes, void *ws_mem, size_t ws_bytes, void *stream) {
    VSUM_REQUIRE(h && h->loaded, VSUM_EINVAL, "vsum_scorer_forward_train: handle without weights");
    VSUM_REQUIRE(B > 0 && T > 0 && x && cu && scores && tape_mem && ws_mem, VSUM_EINVAL, "vsum_scorer_forward_train: bad argument");
    VSUM_REQUIRE(p >= 0.f && p < 1.f, VSUM_EINVAL, "vsum_scorer_forward_train: dropout %f", p);
    VSUM_REQUIRE(!h->cfg.use_pos || max_len <= h->pos_rows, VSUM_EINVAL, "vsum_scorer_forward_train: positional table too short");
    VSUM_REQUIRE(tape_bytes >= vsum_scorer_tape_bytes(h, T) && ws_bytes >= vsum_scorer_train_workspace_bytes(h, T, B),
                 VSUM_ENOMEM, "vsum_scorer_forward_train: tape or workspace too small");
    const vsum_scorer_config &c = h->cfg;
    const int d = c.d_model;
    cudaStream_t s = (cudaStream_t)stream;
    Tape t; TrainWs w;
    carve_tape(c, T, tape_mem, t);
    carve_train_ws(c, T, B, ws_mem, w);
    int rc;
    RUN(launch_row_positions(cu, B, T, w.row_pos, nullptr, s));
    RUN(lin_fwd(h, x, h->w32 + h->embed_w, h->w32 + h->embed_b, t.x0, T, d, c.in_features,
                c.use_pos ? EPI_BIAS_POS : EPI_BIAS, w.row_pos, s));
    const float scale = 1.0f / sqrtf((float)d);
    const float *xin = t.x0;
    const bool tc_attn = h->train_mode == 2;
    const bool attn2 = tc_attn && attention_kernel_version() >= 2;
    if (attn2) RUN(launch_attn2_schedule(cu, B, T, w.a2, s));
    else if (tc_attn) RUN(launch_attn_schedule(cu, B, w.tile_video, w.tile_q0, w.n_tiles, w.max_tiles, s));
    for (int l = 0; l < c.num_layers; ++l) {
        const LayerOffsets &o = h->L[l];
        TapeLayer &L = t.L[l];
        RUN(lin_fwd(h, xin, h->w32 + o.wqkv, h->w32 + o.bqkv, L.qkv, T, 3 * d, d, EPI_BIAS, nullptr, s));
        if (tc_attn) {   // bf16 operands on tcgen05; L.lse holds log2-domain values in this mode
            RUN(launch_f32_to_bf16(L.qkv, w.x16, (int64_t)T * 3 * d, s));
            if (attn2) RUN(launch_attention2_tc05(w.x16, cu, B, T, scale, L.att, w.a2, s, L.lse, p, site_seed(seed, SITE_ATTN, l)));
            else RUN(launch_attention_tc05(w.x16, cu, w.tile_video, w.tile_q0, w.n_tiles, w.max_tiles, T, scale, L.att, s, L.lse, p,
                                      site_seed(seed, SITE_ATTN, l)));
        } else
        RUN(launch_attention_f32(L.qkv, cu, B, max_len, d, c.num_heads, scale, L.att, s, L.lse, p, site_seed(seed, SITE_ATTN, l)));
        RUN(lin_fwd(h, L.att, h->w32 + o.wo, h->w32 + o.bo, w.a, T, d, d, EPI_BIAS, nullptr, s));
        RUN(launch_add_dropout_layernorm_f32(w.a, xin, h->w32 + o.ln1g, h->w32 + o.ln1b, L.s1, L.xmid, T, d, p, site_seed(seed, SITE_PROJ, l), s));
        RUN(lin_fwd(h, L.xmid, h->w32 + o.fc1w, h->w32 + o.fc1b, L.hid, T, c.d_ff, d, EPI_BIAS_RELU, nullptr, s));
        RUN(launch_dropout_inplace_f32(L.hid, (int64_t)T * c.d_ff, p, site_seed(seed, SITE_HIDDEN, l), s));
        RUN(lin_fwd(h, L.hid, h->w32 + o.fc2w, h->w32 + o.fc2b, w.a, T, d, c.d_ff, EPI_BIAS, nullptr, s));
        RUN(launch_add_dropout_layernorm_f32(w.a, L.xmid, h->w32 + o.ln2g, h->w32 + o.ln2b, L.s2, L.xout, T, d, p, site_seed(seed, SITE_MLP, l), s));
        xin = L.xout;
    }
    RUN(launch_head_f32(xin, h->w32 + h->final_w, h->w32 + h->final_b, scores, T, d, c.num_classes, 0, s));
    if (feats) VSUM_CUDA_OK(cudaMemcpyAsync(feats, xin, (size_t)T * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return VSUM_OK;
}

static int scorer_backward_impl(vsum_scorer_t h, const float *x, const int32_t *cu, int32_t B, int64_t T,
                                int32_t max_len, float p, uint64_t seed, const float *d_scores, const float *d_feats,
                                const void *tape_mem, const vsum_scorer_grads *g, void *ws_mem, size_t ws_bytes,
                                void *stream, vsum_grad_bucket_hook hook, void *hook_user) {
    VSUM_REQUIRE(h && h->loaded && g, VSUM_EINVAL, "vsum_scorer_backward: bad handle / grads");
    VSUM_REQUIRE(B > 0 && T > 0 && x && cu && d_scores && tape_mem && ws_mem, VSUM_EINVAL, "vsum_scorer_backward: bad argument");
    VSUM_REQUIRE(ws_bytes >= vsum_scorer_train_workspace_bytes(h, T, B), VSUM_ENOMEM, "vsum_scorer_backward: workspace too small");
    const vsum_scorer_config &c = h->cfg;
    const size_t d = c.d_model, ff = c.d_ff, C = c.num_classes;
    cudaStream_t s = (cudaStream_t)stream;
    Tape t; TrainWs w;
    carve_tape(c, T, const_cast<void *>(tape_mem), t);
    carve_train_ws(c, T, B, ws_mem, w);
    int rc;
#define ZERO(ptr, n) do { VSUM_REQUIRE((ptr) != nullptr, VSUM_EINVAL, "vsum_scorer_backward: null gradient " #ptr); \
                          if (!pre_zeroed) VSUM_CUDA_OK(cudaMemsetAsync((ptr), 0, (n) * sizeof(float), s)); } while (0)
    const bool pre_zeroed = g->pre_zeroed != 0;
    ZERO(g->embed_w, d * c.in_features); ZERO(g->embed_b, d); ZERO(g->final_w, C * d); ZERO(g->final_b, C);
    const float scale = 1.0f / sqrtf((float)d);
    const float *x_last = t.L[c.num_layers - 1].xout;
    const bool tc_attn = h->train_mode == 2;
    if (tc_attn) RUN(launch_attn_schedule(cu, B, w.tile_video, w.tile_q0, w.n_tiles, w.max_tiles, s));
    RUN(launch_head_bwd_f32(x_last, h->w32 + h->final_w, d_scores, d_feats, w.a, g->final_w, g->final_b, T, (int)d, (int)C, s));
    for (int l = c.num_layers - 1; l >= 0; --l) {
        const LayerOffsets &o = h->L[l];
        const TapeLayer &L = t.L[l];
        const vsum_layer_grads &gl = g->layers[l];
        const float *xin = l == 0 ? t.x0 : t.L[l - 1].xout;
        ZERO(gl.ln2_g, d); ZERO(gl.ln2_b, d); ZERO(gl.fc2_w, d * ff); ZERO(gl.fc2_b, d); ZERO(gl.fc1_w, ff * d); ZERO(gl.fc1_b, ff);
        ZERO(gl.ln1_g, d); ZERO(gl.ln1_b, d); ZERO(gl.o_w, d * d); ZERO(gl.o_b, d);
        VSUM_REQUIRE(gl.q_w && gl.q_b && gl.k_w && gl.k_b && gl.v_w && gl.v_b, VSUM_EINVAL, "vsum_scorer_backward: null q/k/v gradient");
        // q|k|v gradients adjacent in memory (flat gradient buffer): the fused [3d,d] wgrad lands in place
        const bool qkv_in_place = gl.k_w == gl.q_w + d * d && gl.v_w == gl.k_w + d * d && gl.k_b == gl.q_b + d && gl.v_b == gl.k_b + d;
        if (qkv_in_place) { ZERO(gl.q_w, 3 * d * d); ZERO(gl.q_b, 3 * d); }
        else { VSUM_CUDA_OK(cudaMemsetAsync(w.dwqkv, 0, 3 * d * d * sizeof(float), s)); VSUM_CUDA_OK(cudaMemsetAsync(w.dbqkv, 0, 3 * d * sizeof(float), s)); }
        // xout = LN2(s2), s2 = dropout(mlp) + xmid          a: d_xout -> b: d_s2 (residual path), c: d_mlp
        RUN(launch_layernorm_bwd_f32(w.a, L.s2, h->w32 + o.ln2g, w.b, w.c, gl.ln2_g, gl.ln2_b, T, (int)d, p, site_seed(seed, SITE_MLP, l), s));
        // mlp = hid W2^T + b2
        RUN(lin_wgrad(h, w.c, L.hid, gl.fc2_w, gl.fc2_b, T, (int)d, (int)ff, w.y16, w.x16, s));
        RUN(lin_dgrad(h, w.c, h->w32 + o.fc2w, h->w32 + o.t_fc2, w.dhid, nullptr, T, (int)d, (int)ff, 0, s));
        RUN(launch_relu_dropout_bwd_f32(L.hid, w.dhid, (int64_t)T * ff, p, s));
        // h1 = xmid W1^T + b1                               b: d_xmid += dh1 W1
        RUN(lin_wgrad(h, w.dhid, L.xmid, gl.fc1_w, gl.fc1_b, T, (int)ff, (int)d, w.y16, w.x16, s));
        RUN(lin_dgrad(h, w.dhid, h->w32 + o.fc1w, h->w32 + o.t_fc1, w.b, w.dd, T, (int)ff, (int)d, 1, s));
        // xmid = LN1(s1), s1 = dropout(proj) + xin          b: d_xmid -> a: d_s1 (residual path), c: d_proj
        RUN(launch_layernorm_bwd_f32(w.b, L.s1, h->w32 + o.ln1g, w.a, w.c, gl.ln1_g, gl.ln1_b, T, (int)d, p, site_seed(seed, SITE_PROJ, l), s));
        // proj = att Wo^T + bo                              dd: d_att
        RUN(lin_wgrad(h, w.c, L.att, gl.o_w, gl.o_b, T, (int)d, (int)d, w.y16, w.x16, s));
        RUN(lin_dgrad(h, w.c, h->w32 + o.wo, h->w32 + o.t_wo, w.dd, nullptr, T, (int)d, (int)d, 0, s));
        if (tc_attn) {
            RUN(launch_f32_to_bf16(L.qkv, w.x16, (int64_t)T * 3 * d, s));
            RUN(launch_f32_to_bf16(w.dd, w.y16, (int64_t)T * d, s));
            RUN(launch_attn_delta_bf16(L.att, w.y16, w.delta, T, s));
            RUN(launch_attention_bwd_tc05(w.x16, w.y16, L.lse, w.delta, cu, w.tile_video, w.tile_q0, w.n_tiles, w.max_tiles, T,
                                          scale, p, site_seed(seed, SITE_ATTN, l), w.dqkv, s));
        } else
        RUN(launch_attention_bwd_f32(L.qkv, L.att, w.dd, L.lse, cu, B, max_len, T, (int)d, c.num_heads, scale, p,
                                     site_seed(seed, SITE_ATTN, l), w.delta, w.dqkv, s));
        // qkv = xin Wqkv^T + bqkv                           a: d_xin += dqkv Wqkv
        RUN(lin_wgrad(h, w.dqkv, xin, qkv_in_place ? gl.q_w : w.dwqkv, qkv_in_place ? gl.q_b : w.dbqkv, T, (int)(3 * d), (int)d,
                      w.y16, w.x16, s));
        RUN(lin_dgrad(h, w.dqkv, h->w32 + o.wqkv, h->w32 + o.t_wqkv, w.a, w.dd, T, (int)(3 * d), (int)d, 1, s));
        float *wdst[3] = {gl.q_w, gl.k_w, gl.v_w}, *bdst[3] = {gl.q_b, gl.k_b, gl.v_b};
        for (int i = 0; i < 3 && !qkv_in_place; ++i) {
            VSUM_CUDA_OK(cudaMemcpyAsync(wdst[i], w.dwqkv + i * d * d, d * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
            VSUM_CUDA_OK(cudaMemcpyAsync(bdst[i], w.dbqkv + i * d, d * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        if (hook) hook(hook_user, l);          // every gradient of layer l (and, for the last layer, of the head) is queued on `stream`
    }
    // x0 = features We^T + be (+ positions)
    RUN(lin_wgrad(h, w.a, x, g->embed_w, g->embed_b, T, (int)d, c.in_features, w.y16, w.x16, s));
    if (hook) hook(hook_user, -1);
#undef ZERO
    return VSUM_OK;
}

extern "C" int vsum_scorer_backward(vsum_scorer_t h, const float *x, const int32_t *cu, int32_t B, int64_t T,
                                    int32_t max_len, float p, uint64_t seed, const float *d_scores, const float *d_feats,
                                    const void *tape_mem, const vsum_scorer_grads *g, void *ws_mem, size_t ws_bytes,
                                    void *stream) {
    return scorer_backward_impl(h, x, cu, B, T, max_len, p, seed, d_scores, d_feats, tape_mem, g, ws_mem, ws_bytes, stream, nullptr, nullptr);
}

extern "C" int vsum_scorer_backward_hooked(vsum_scorer_t h, const float *x, const int32_t *cu, int32_t B, int64_t T,
                                           int32_t max_len, float p, uint64_t seed, const float *d_scores, const float *d_feats,
                                           const void *tape_mem, const vsum_scorer_grads *g, void *ws_mem, size_t ws_bytes,
                                           void *stream, vsum_grad_bucket_hook hook, void *hook_user) {
    return scorer_backward_impl(h, x, cu, B, T, max_len, p, seed, d_scores, d_feats, tape_mem, g, ws_mem, ws_bytes, stream, hook, hook_user);
}

// Data-parallel step, last stage (after the SUM all-reduces): ext = [sum of squared errors, sum of batch sizes, Nmax of rank 0,
// ..., Nmax of rank world-1] -> D = (sum of batch sizes) * (max Nmax), the padded size mse_with_mask_loss divides by
// (src/utils/utils.py:55) for the GLOBAL batch; grads *= 1 / D, loss_out = ext[0] / D.
namespace {
__global__ void dp_finalize_kernel(float *__restrict__ grads, int64_t n, const float *__restrict__ ext, int world, float *__restrict__ loss_out) {
    float nmax = 0.f;
    for (int r = 0; r < world; ++r) nmax = fmaxf(nmax, ext[2 + r]);
    const float inv = 1.0f / (ext[1] * nmax);
    if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) *loss_out = ext[0] * inv;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) grads[i] *= inv;
}
}  // namespace

namespace {
__global__ void dp_extras_kernel(float *__restrict__ ext, const float *__restrict__ loss_sum, float batch, float nmax, int rank, int world) {
    const int i = threadIdx.x;
    if (i < 2 + world) ext[i] = i == 0 ? *loss_sum : (i == 1 ? batch : (i - 2 == rank ? nmax : 0.f));
}
}  // namespace

// ext[2 + world] = [*loss_sum, batch, Nmax one-hot at `rank`]: the operand of the extras' SUM all-reduce (scalars travel as
// kernel arguments, so a host that runs steps ahead of the device cannot overwrite them)
extern "C" int vsum_dp_extras(float *ext, const float *loss_sum, int32_t batch, int32_t nmax, int32_t rank, int32_t world, void *stream) {
    VSUM_REQUIRE(ext && loss_sum && world >= 1 && world <= 1022 && rank >= 0 && rank < world, VSUM_EINVAL, "vsum_dp_extras: bad argument");
    dp_extras_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ext, loss_sum, (float)batch, (float)nmax, rank, world);
    VSUM_LAUNCH_OK("dp_extras_kernel");
    return VSUM_OK;
}

extern "C" int vsum_dp_finalize(float *grads, int64_t n, const float *ext, int32_t world, float *loss_out, void *stream) {
    VSUM_REQUIRE(grads && ext && n >= 0 && world >= 1, VSUM_EINVAL, "vsum_dp_finalize: bad argument");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 4, (n + 255) / 256));
    dp_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grads, n, ext, world, loss_out);
    VSUM_LAUNCH_OK("dp_finalize_kernel");
    return VSUM_OK;
}
#undef RUN

// ---- stand-alone Linear layer (PretrainModel.video_transform, simnet_pretrain.py:33,80) ---------------
namespace {
struct LinWs { float *wt, *zeros; __nv_bfloat16 *y16, *x16; };
size_t carve_lin_ws(int64_t M, int N, int K, void *base, LinWs &w) {
    Carver k{(uint8_t *)base};
    w.wt = k.get<float>((size_t)N * K); w.zeros = k.get<float>(K);
    w.y16 = k.get<__nv_bfloat16>((size_t)M * N); w.x16 = k.get<__nv_bfloat16>((size_t)M * K);
    return align_up(k.off, 1024);
}
bool lin_tc_ok(int N, int K) { return N % 256 == 0 && K % 256 == 0 && N <= 4096 && K <= 4096; }
}  // namespace

extern "C" size_t vsum_linear_workspace_bytes(int64_t M, int32_t N, int32_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    LinWs w;
    return carve_lin_ws(M, N, K, nullptr, w);
}

extern "C" int vsum_linear_forward(const float *x, const float *w, const float *bias, float *y, int64_t M, int32_t N,
                                   int32_t K, int32_t mode, void *stream) {
    VSUM_REQUIRE(x && w && bias && y && M >= 0 && N > 0 && K > 0 && (mode == 0 || mode == 1), VSUM_EINVAL, "vsum_linear_forward: bad argument");
    if (M == 0) return VSUM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 0 || !lin_tc_ok(N, K)) return launch_linear_f32(x, w, bias, y, M, N, K, EPI_BIAS, nullptr, nullptr, 0, s);
    Tc05GemmArgs g{};
    g.A = x; g.W = w; g.M = M; g.N = N; g.K = K; g.a_is_f32 = 1; g.bias = bias; g.out_f32 = y; g.epi = TC_EPI_BIAS_F32;
    g.prof_cat = PROF_OTHER;
    return launch_gemm_tc05(g, s);
}

// dx [M,K] = dy W (dx may be NULL);  dw [N,K] = dy^T x;  db [N] = colsum(dy)   (dw, db overwritten)
extern "C" int vsum_linear_backward(const float *dy, const float *x, const float *w, float *dx, float *dw, float *db,
                                    int64_t M, int32_t N, int32_t K, int32_t mode, void *ws_mem, size_t ws_bytes, void *stream) {
    VSUM_REQUIRE(dy && x && w && dw && db && M >= 0 && N > 0 && K > 0 && (mode == 0 || mode == 1), VSUM_EINVAL, "vsum_linear_backward: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    VSUM_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), s));
    VSUM_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), s));
    if (M == 0) return VSUM_OK;
    int rc;
    if (mode == 0 || !lin_tc_ok(N, K)) {
        if ((rc = launch_linear_wgrad_f32(dy, x, dw, db, M, N, K, s))) return rc;
        return dx ? launch_linear_dgrad_f32(dy, w, dx, M, N, K, 0, s) : VSUM_OK;
    }
    VSUM_REQUIRE(ws_mem && ((uintptr_t)ws_mem & 1023) == 0 && ws_bytes >= vsum_linear_workspace_bytes(M, N, K), VSUM_ENOMEM,
                 "vsum_linear_backward: workspace missing, unaligned or too small");
    LinWs k;
    carve_lin_ws(M, N, K, ws_mem, k);
    if ((rc = launch_linear_wgrad_tc05(dy, x, dw, db, M, N, K, s, k.y16, k.x16))) return rc;
    if (!dx) return VSUM_OK;
    if ((rc = launch_transpose_f32(w, k.wt, N, K, s))) return rc;
    VSUM_CUDA_OK(cudaMemsetAsync(k.zeros, 0, (size_t)K * sizeof(float), s));
    Tc05GemmArgs g{};
    g.A = dy; g.W = k.wt; g.M = M; g.N = K; g.K = N; g.a_is_f32 = 1; g.bias = k.zeros; g.out_f32 = dx; g.epi = TC_EPI_BIAS_F32;
    g.prof_cat = PROF_OTHER;
    return launch_gemm_tc05(g, s);
}

extern "C" int vsum_masked_mse(const float *out, const float *tgt, const uint8_t *pad_mask, int64_t n, float denom,
                               float *loss_out, float grad_scale, float *d_out, void *stream) {
    VSUM_REQUIRE(out && tgt && n >= 0 && denom > 0.f, VSUM_EINVAL, "vsum_masked_mse: bad argument");
    return launch_masked_mse_f32(out, tgt, pad_mask, n, denom, loss_out, grad_scale, d_out, (cudaStream_t)stream);
}

// ---- diagnostics -----------------------------------------------------------------------------
extern "C" int vsum_debug_gemm_tc05(const void *A, const void *W, const float *bias, const void *residual,
                                    const float *gamma, const float *beta, void *out, int64_t M, int32_t N,
                                    int32_t K, int32_t a_is_f32, int32_t epi, void *stream) {
    VSUM_REQUIRE(A && W && bias && out, VSUM_EINVAL, "vsum_debug_gemm_tc05: null pointer");
    VSUM_REQUIRE(epi == TC_EPI_BIAS || epi == TC_EPI_BIAS_RELU || epi == TC_EPI_BIAS_RES_LN || epi == TC_EPI_BIAS_F32 ||
                 epi == TC_EPI_BIAS_RELU_F32, VSUM_EINVAL, "vsum_debug_gemm_tc05: epi %d", epi);
    Tc05GemmArgs g{};
    g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.a_is_f32 = a_is_f32; g.epi = epi; g.bias = bias;
    g.prof_cat = PROF_OTHER;
    if (epi >= TC_EPI_BIAS_F32) g.out_f32 = (float *)out; else
    g.out = (__nv_bfloat16 *)out; g.residual = (const __nv_bfloat16 *)residual; g.gamma = gamma; g.beta = beta;
    return launch_gemm_tc05(g, (cudaStream_t)stream);
}

extern "C" int vsum_debug_ffn_tc05(const void *x, const void *w1, const float *b1, const void *w2, const float *b2,
                                   const float *gamma, const float *beta, void *out, int64_t M, void *stream) {
    VSUM_REQUIRE(x && w1 && b1 && w2 && b2 && gamma && beta && out, VSUM_EINVAL, "vsum_debug_ffn_tc05: null pointer");
    Tc05FfnArgs f{};
    f.x = (const __nv_bfloat16 *)x; f.w1 = (const __nv_bfloat16 *)w1; f.w2 = (const __nv_bfloat16 *)w2; f.b1 = b1; f.b2 = b2;
    f.gamma = gamma; f.beta = beta; f.M = M; f.out = (__nv_bfloat16 *)out;
    return launch_ffn_tc05(f, (cudaStream_t)stream);
}

extern "C" int vsum_debug_wgrad_tc05(const float *dY, const float *X, float *dW, float *db, int64_t M, int32_t N,
                                     int32_t K, void *scratch_bf16, void *stream) {
    VSUM_REQUIRE(dY && X && dW, VSUM_EINVAL, "vsum_debug_wgrad_tc05: null pointer");
    __nv_bfloat16 *y16 = (__nv_bfloat16 *)scratch_bf16, *x16 = y16 ? y16 + (size_t)M * N : nullptr;
    return launch_linear_wgrad_tc05(dY, X, dW, db, M, N, K, (cudaStream_t)stream, y16, x16);
}

extern "C" int vsum_debug_attention_train_tc05(const void *qkv, const int32_t *cu_seqlens, int32_t B, int64_t T, float *out,
                                               float *lse2, float drop_p, uint64_t seed, int32_t *scratch, void *stream) {
    VSUM_REQUIRE(qkv && cu_seqlens && out && lse2 && scratch, VSUM_EINVAL, "vsum_debug_attention_train_tc05: null pointer");
    const int max_tiles = (int)(T / 128 + B);
    cudaStream_t s = (cudaStream_t)stream;
    if (attention_kernel_version() >= 2) {
        int rc2 = launch_attn2_schedule(cu_seqlens, B, T, scratch, s);
        if (rc2) return rc2;
        return launch_attention2_tc05((const __nv_bfloat16 *)qkv, cu_seqlens, B, T, 1.0f / 16.0f, out, scratch, s, lse2, drop_p, seed);
    }
    int rc = launch_attn_schedule(cu_seqlens, B, scratch, scratch + max_tiles, scratch + 2 * max_tiles, max_tiles, s);
    if (rc) return rc;
    return launch_attention_tc05((const __nv_bfloat16 *)qkv, cu_seqlens, scratch, scratch + max_tiles, scratch + 2 * max_tiles,
                                 max_tiles, T, 1.0f / 16.0f, out, s, lse2, drop_p, seed);
}

extern "C" int vsum_debug_attention_bwd_tc05(const void *qkv, const void *d_out, const float *lse2, const float *delta,
                                             const int32_t *cu_seqlens, int32_t B, int64_t T, float drop_p, uint64_t seed,
                                             float *dqkv, int32_t *scratch, void *stream) {
    VSUM_REQUIRE(qkv && d_out && lse2 && delta && cu_seqlens && dqkv && scratch, VSUM_EINVAL,
                 "vsum_debug_attention_bwd_tc05: null pointer");
    const int max_tiles = (int)(T / 128 + B);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = launch_attn_schedule(cu_seqlens, B, scratch, scratch + max_tiles, scratch + 2 * max_tiles, max_tiles, s);
    if (rc) return rc;
    return launch_attention_bwd_tc05((const __nv_bfloat16 *)qkv, (const __nv_bfloat16 *)d_out, lse2, delta, cu_seqlens, scratch,
                                     scratch + max_tiles, scratch + 2 * max_tiles, max_tiles, T, 1.0f / 16.0f, drop_p, seed, dqkv, s);
}

extern "C" int vsum_debug_attention_scaled_tc05(const void *qkv, const int32_t *cu_seqlens, int32_t B, int64_t T, float scale,
                                                void *out, int32_t *scratch, void *stream) {
    VSUM_REQUIRE(qkv && cu_seqlens && out && scratch, VSUM_EINVAL, "vsum_debug_attention_scaled_tc05: null pointer");
    const int max_tiles = (int)(T / 128 + B);
    cudaStream_t s = (cudaStream_t)stream;
    if (attention_kernel_version() >= 2) {
        int rc2 = launch_attn2_schedule(cu_seqlens, B, T, scratch, s);
        if (rc2) return rc2;
        return launch_attention2_tc05((const __nv_bfloat16 *)qkv, cu_seqlens, B, T, scale, out, scratch, s);
    }
    int rc = launch_attn_schedule(cu_seqlens, B, scratch, scratch + max_tiles, scratch + 2 * max_tiles, max_tiles, s);
    if (rc) return rc;
    return launch_attention_tc05((const __nv_bfloat16 *)qkv, cu_seqlens, scratch, scratch + max_tiles,
                                 scratch + 2 * max_tiles, max_tiles, T, scale, (__nv_bfloat16 *)out, s);
}

extern "C" int vsum_debug_attention_tc05(const void *qkv, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                         void *out, int32_t *scratch, void *stream) {
    return vsum_debug_attention_scaled_tc05(qkv, cu_seqlens, B, T, 1.0f / 16.0f, out, scratch, stream);
}
