// Internal launcher declarations shared between the translation units of libvsum_b200.
#pragma once

#include <cuda_bf16.h>
#include "vsum_common.cuh"

namespace vsum {

// ---- fp32 SIMT path (vsum_scorer_fp32.cu) ----------------------------------------------------
enum LinearEpilogue { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_POS = 2 };

int launch_row_positions(const int32_t *cu_seqlens, int B, int64_t T, int32_t *row_pos,
                         int32_t *row_vid, cudaStream_t s);
// C[M,N] = epi(A[M,K] * W[N,K]^T + bias).  EPI_BIAS_POS adds pos_table[row_pos[m], n].
int launch_linear_f32(const float *A, const float *W, const float *bias, float *C, int64_t M, int N,
                      int K, int epi, const float *pos_table, const int32_t *row_pos, int pos_rows,
                      cudaStream_t s);
// out[m,:] = LayerNorm(a[m,:] + res[m,:]) * gamma + beta   (eps 1e-5, biased variance)
int launch_add_layernorm_f32(const float *a, const float *res, const float *gamma, const float *beta,
                             float *out, int64_t M, int d, cudaStream_t s);
// qkv [T, 3d] (q | k | v, head h = columns [h*hd, (h+1)*hd) of each third) -> out [T, d]
int launch_attention_f32(const float *qkv, const int32_t *cu_seqlens, int B, int max_len, int d,
                         int num_heads, float scale, float *out, cudaStream_t s, float *lse = nullptr,
                         float drop_p = 0.f, unsigned long long seed = 0);

// ---- fp32 training path (vsum_train_fp32.cu) --------------------------------------------------
// Counter-based dropout: the keep decision is a pure function of (seed, element index), so the
// backward pass recomputes the forward's masks instead of storing them.
__host__ __device__ __forceinline__ unsigned long long attn_drop_index(long long q_row, int h, int H, int key) {
    return ((unsigned long long)(q_row * H + h) << 20) ^ (unsigned long long)key ^ 0xA5A5000000000000ULL;
}
// Tensor-core attention (training): one 64-bit draw decides FOUR consecutive keys of a query row, 16 bits
// each (drop iff bits < thresh16 = round(p * 65536)), so forward and backward hash once per 4 elements.
// The draw is four Philox-2x32 rounds (mulhi/mullo by 0xD256D193, key schedule += 0x9E3779B9) keyed by the
// 64-bit seed: one IMAD.WIDE + one LOP3 per round, ~2 instructions per attention element -- the splitmix64
// used by the other dropout sites costs ~7 on 32-bit ALUs and doubled the forward kernel's time.
__host__ __device__ __forceinline__ unsigned long long dropout_bits64(unsigned long long seed, unsigned long long idx) {
    unsigned int x0 = (unsigned int)idx, x1 = (unsigned int)(idx >> 32);
    unsigned int k = (unsigned int)seed ^ ((unsigned int)(seed >> 32) * 0x85EBCA6Bu);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned long long m = (unsigned long long)x0 * 0xD256D193ull;
        x0 = (unsigned int)(m >> 32) ^ k ^ x1;
        x1 = (unsigned int)m;
        k += 0x9E3779B9u;
    }
    return ((unsigned long long)x0 << 32) | x1;
}
// Element idx is lane (idx & 3) of draw (idx >> 2): kernels that walk four consecutive elements per thread pay
// for one draw, and the decision is the same whatever the access pattern (forward and backward agree).
__host__ __device__ __forceinline__ unsigned int dropout_thresh16(float p) { return p <= 0.f ? 0u : (unsigned int)(p * 65536.0f + 0.5f); }
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, float p) {
    const unsigned long long z = dropout_bits64(seed, idx >> 2);
    return (unsigned int)((z >> (16 * (idx & 3))) & 0xffffu) >= dropout_thresh16(p);
}
__host__ __device__ __forceinline__ unsigned long long attn_drop_group_index(long long q_row, int h, int H, int key_group) {
    return ((unsigned long long)(q_row * H + h) << 20) ^ (unsigned long long)key_group ^ 0x5A5A000000000000ULL;
}
__host__ __device__ __forceinline__ unsigned long long site_seed(unsigned long long seed, int site, int layer) {
    return seed ^ ((unsigned long long)(site + 1) << 56) ^ ((unsigned long long)(layer + 1) << 48);
}
enum DropSite { SITE_ATTN = 0, SITE_PROJ = 1, SITE_HIDDEN = 2, SITE_MLP = 3 };

// s = dropout(a) + res (written to s_out), out = LayerNorm(s)
int launch_add_dropout_layernorm_f32(const float *a, const float *res, const float *gamma, const float *beta,
                                     float *s_out, float *out, int64_t M, int d, float drop_p,
                                     unsigned long long seed, cudaStream_t s);
int launch_dropout_inplace_f32(float *x, int64_t n, float drop_p, unsigned long long seed, cudaStream_t s);
// y = LayerNorm(s): (dy, s, gamma) -> ds, d_a = dropout_bwd(ds) (optional), dgamma += , dbeta +=
int launch_layernorm_bwd_f32(const float *dy, const float *s_in, const float *gamma, float *ds, float *d_a,
                             float *dgamma, float *dbeta, int64_t M, int d, float drop_p,
                             unsigned long long seed, cudaStream_t s);
// dW[N,K] += dY[M,N]^T X[M,K];  db[N] += colsum(dY)   (dW, db must be zeroed by the caller)
int launch_linear_wgrad_f32(const float *dY, const float *X, float *dW, float *db, int64_t M, int N, int K,
                            cudaStream_t s);
// dX[M,K] = dY[M,N] W[N,K]  (+= when accumulate)
int launch_linear_dgrad_f32(const float *dY, const float *W, float *dX, int64_t M, int N, int K, int accumulate,
                            cudaStream_t s);
// dh = (hid > 0) ? dhid * keep_scale : 0   (hid is the saved post-ReLU, post-dropout activation)
int launch_relu_dropout_bwd_f32(const float *hid, float *dhid, int64_t n, float drop_p, cudaStream_t s);
int launch_attention_bwd_f32(const float *qkv, const float *o, const float *d_o, const float *lse,
                             const int32_t *cu_seqlens, int B, int max_len, int64_t T, int d, int num_heads,
                             float scale, float drop_p, unsigned long long seed, float *delta_ws, float *dqkv,
                             cudaStream_t s);
int launch_head_bwd_f32(const float *x, const float *w, const float *d_scores, const float *d_feats, float *dx,
                        float *dw, float *db, int64_t M, int d, int C, cudaStream_t s);
// masked MSE (src/utils/utils.py:45-56): loss = sum(((out - tgt) * keep)^2) / denom; d_out optional
int launch_masked_mse_f32(const float *out, const float *tgt, const uint8_t *pad_mask, int64_t n, float denom,
                          float *loss, float grad_scale, float *d_out, cudaStream_t s);
// scores[m, c] = x[m,:] . w[c,:] + b[c], optional sigmoid
int launch_head_f32(const float *x, const float *w, const float *b, float *scores, int64_t M, int d,
                    int num_classes, int apply_sigmoid, cudaStream_t s);

// ---- bf16 tcgen05 path (vsum_gemm_tc05.cu, vsum_attn_tc05.cu) --------------------------------
struct Tc05Layer {
    const __nv_bfloat16 *w_qkv;  // [768,256]
    const float *b_qkv;          // [768]
    const __nv_bfloat16 *w_o;    // [256,256]
    const float *b_o, *ln1_g, *ln1_b;
    const __nv_bfloat16 *w_fc1;  // [1024,256]
    const float *b_fc1;
    const __nv_bfloat16 *w_fc2;  // [256,1024]
    const float *b_fc2, *ln2_g, *ln2_b;
};

enum Tc05Epilogue {
    TC_EPI_BIAS = 0,         // out bf16 = acc + bias
    TC_EPI_BIAS_RELU = 1,    // out bf16 = relu(acc + bias)
    TC_EPI_BIAS_POS = 2,     // out bf16 = acc + bias + pos_table[row_pos[m]]       (N == 256)
    TC_EPI_BIAS_RES_LN = 3,  // out bf16 = LN(acc + bias + residual) * g + b        (N == 256)
    TC_EPI_BIAS_RES_LN_HEAD = 4, // ... plus feats fp32 and score = sigmoid?(y . w_head + b_head)
    TC_EPI_BIAS_F32 = 5,         // fp32 output variants (training path, tf32 operands): out_f32 = acc + bias
    TC_EPI_BIAS_RELU_F32 = 6,
    TC_EPI_BIAS_POS_F32 = 7
};

struct Tc05GemmArgs {
    const void *A;            // [M,K] bf16 (or fp32 when a_is_f32)
    const void *W;            // [N,K] same element type as A
    int64_t M;
    int N, K;
    int a_is_f32;             // 1: tf32 MMA on fp32 operands (feature embedding)
    int epi;
    const float *bias;        // [N]
    __nv_bfloat16 *out;       // [M,N] bf16 (may be NULL for the HEAD epilogue)
    float *out_f32;           // [M,N] fp32 for the *_F32 epilogues
    const __nv_bfloat16 *residual;   // [M,256]
    const float *gamma, *beta;       // [256]
    const float *pos_table;   // [pos_rows,256]
    const int32_t *row_pos;   // [M]
    int pos_rows;
    const float *head_w;      // [256]
    const float *head_b;      // [1]
    float *scores_out;        // [M]
    float *feats_out;         // [M,256] fp32 or NULL
    int apply_sigmoid;
    int prof_cat;             // ProfCategory for vsum_profile_*
};
int launch_gemm_tc05(const Tc05GemmArgs &a, cudaStream_t s);

// Fused feed-forward block (vsum_ffn_tc05.cu): out = LN(relu(x W1^T + b1) W2^T + b2 + x) * gamma + beta, d_model 256,
// d_ff 1024; head_w/head_b/scores_out (+ feats_out) as in the HEAD epilogue above; out may be NULL then.
struct Tc05FfnArgs {
    const __nv_bfloat16 *x;      // [M,256] bf16 (also the residual)
    const __nv_bfloat16 *w1;     // [1024,256]
    const __nv_bfloat16 *w2;     // [256,1024]
    const float *b1, *b2, *gamma, *beta;
    int64_t M;
    __nv_bfloat16 *out;          // [M,256] or NULL
    const float *head_w, *head_b;
    float *scores_out, *feats_out;
    int apply_sigmoid;
};
int launch_ffn_tc05(const Tc05FfnArgs &a, cudaStream_t s);

// qkv [T,768] bf16 -> out [T,256] bf16, d_model 256, 4 heads of 64, scale = 1/16
// lse2 != NULL: training variant (log2-domain log-sum-exp [T,4] out, dropout on P with the grouped hash,
// `out` is then fp32 [T,256] instead of bf16)
int launch_attention_tc05(const __nv_bfloat16 *qkv, const int32_t *cu_seqlens, const int32_t *tile_video,
                          const int32_t *tile_q0, const int32_t *n_tiles_ptr, int max_tiles, int64_t T,
                          float scale, void *out, cudaStream_t s, float *lse2 = nullptr, float drop_p = 0.f,
                          unsigned long long seed = 0);
// Second-generation forward (vsum_attn2_tc05.cu): persistent, two 128-query tiles per CTA, P in tensor memory.  Same
// arguments; `scratch` holds attention2_scratch_ints(T, B) int32 and is filled once per batch by launch_attn2_schedule.
size_t attention2_scratch_ints(int64_t T, int B);
int launch_attn2_schedule(const int32_t *cu_seqlens, int B, int64_t T, int32_t *scratch, cudaStream_t s);
int launch_attention2_tc05(const __nv_bfloat16 *qkv, const int32_t *cu_seqlens, int B, int64_t T, float scale, void *out,
                           int32_t *scratch, cudaStream_t s, float *lse2 = nullptr, float drop_p = 0.f,
                           unsigned long long seed = 0);
inline uint32_t attn_drop_thresh16(float p) { return p <= 0.f ? 0u : (uint32_t)(p * 65536.0f + 0.5f); }
// Backward of the above on tcgen05 (vsum_attn_bwd_tc05.cu): qkv16 [T,768], dO16 [T,256] bf16, lse2 / delta [T,4]
// fp32 -> dqkv [T,768] fp32 (overwritten).  Tiles = the forward's schedule (video, first key row).
int launch_attention_bwd_tc05(const __nv_bfloat16 *qkv16, const __nv_bfloat16 *dO16, const float *lse2, const float *delta,
                              const int32_t *cu_seqlens, const int32_t *tile_video, const int32_t *tile_k0,
                              const int32_t *n_tiles_ptr, int max_tiles, int64_t T, float scale, float drop_p,
                              unsigned long long seed, float *dqkv, cudaStream_t s);
// delta[t, h] = sum_c o[t, h*64 + c] * dO16[t, h*64 + c]: the SAME rounded dO the backward MMAs consume, so
// that every row of dS sums to zero up to fp32 rounding (d_model 256, 4 heads)
int launch_attn_delta_bf16(const float *o, const __nv_bfloat16 *dO16, float *delta, int64_t T, cudaStream_t s);
int launch_attn_schedule(const int32_t *cu_seqlens, int B, int32_t *tile_video, int32_t *tile_q0,
                         int32_t *n_tiles_out, int max_tiles, cudaStream_t s);

// dW[N,K] += dY[M,N]^T X[M,K] on tcgen05 (bf16 operands), db[N] += colsum(dY)   (vsum_wgrad_tc05.cu)
int launch_linear_wgrad_tc05(const float *dY, const float *X, float *dW, float *db, int64_t M, int N, int K,
                             cudaStream_t s, __nv_bfloat16 *dY16 = nullptr, __nv_bfloat16 *X16 = nullptr);
int launch_transpose_f32(const float *in, float *out, int rows, int cols, cudaStream_t s);   // out[c][r] = in[r][c]
int launch_add_inplace_f32(float *dst, const float *src, int64_t n, cudaStream_t s);         // dst += src

int launch_f32_to_bf16(const float *in, __nv_bfloat16 *out, int64_t n, cudaStream_t s);
// out16 / out32 (either may be NULL) = scale * in
int launch_scale_convert(const float *in, __nv_bfloat16 *out16, float *out32, int64_t n, float scale, cudaStream_t s);

}  // namespace vsum
