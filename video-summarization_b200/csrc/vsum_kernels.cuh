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
                         int num_heads, float scale, float *out, cudaStream_t s);
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
    TC_EPI_BIAS_RES_LN_HEAD = 4  // ... plus feats fp32 and score = sigmoid?(y . w_head + b_head)
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

// qkv [T,768] bf16 -> out [T,256] bf16, d_model 256, 4 heads of 64, scale = 1/16
int launch_attention_tc05(const __nv_bfloat16 *qkv, const int32_t *cu_seqlens, const int32_t *tile_video,
                          const int32_t *tile_q0, const int32_t *n_tiles_ptr, int max_tiles, int64_t T,
                          float scale, __nv_bfloat16 *out, cudaStream_t s);
int launch_attn_schedule(const int32_t *cu_seqlens, int B, int32_t *tile_video, int32_t *tile_q0,
                         int32_t *n_tiles_out, int max_tiles, cudaStream_t s);

int launch_f32_to_bf16(const float *in, __nv_bfloat16 *out, int64_t n, cudaStream_t s);

}  // namespace vsum
