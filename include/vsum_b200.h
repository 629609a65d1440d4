/* vsum_b200 -- C ABI of the B200-native video-summarisation hot path.
 *
 * The reference (BerserkerMother/Video-Summarization) is pure Python and has no FFI layer; its
 * boundary for this path is the Python call surface of `src/model` and `src/evaluation`
 * (SURVEY.md section 8(b)).  Each entry point below names the reference function it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VSUM_E* code; the message for the
 *     calling thread is available from vsum_last_error().  No C++ exceptions cross the ABI.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*.  Nothing synchronises the stream.
 *   - videos are PACKED (no padding): rows of video v are [cu[v], cu[v+1]) of the flat arrays.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef VSUM_B200_H
#define VSUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VSUM_API __attribute__((visibility("default")))
#else
#define VSUM_API
#endif

#define VSUM_OK            0
#define VSUM_EINVAL       -1   /* bad argument / unsupported shape */
#define VSUM_ECUDA        -2   /* CUDA runtime or driver error */
#define VSUM_ENOMEM       -3   /* workspace too small / allocation failed */
#define VSUM_EUNSUPPORTED -4   /* configuration outside what the sm_100a kernels are built for */

#define VSUM_MAX_LAYERS 16

#define VSUM_MODE_FP32 0       /* fp32 SIMT kernels: <=1e-5 vs the reference's fp32 path */
#define VSUM_MODE_BF16 1       /* bf16 tcgen05/TMEM/TMA kernels (tf32 for the fp32 feature GEMM) */
#define VSUM_MODE_BF16_FEATURES 2   /* the same kernels fed with bf16 FEATURES (a pack written with features_bf16: half the
                                       bytes over PCIe and out of HBM); the feature GEMM then runs in bf16 as well */

#define VSUM_FSCORE_AVG 0      /* evaluation_metrics.py:32-33 ('avg', TVSum) */
#define VSUM_FSCORE_MAX 1      /* evaluation_metrics.py:30-31 ('max', SumMe) */

VSUM_API int         vsum_abi_version(void);
VSUM_API const char *vsum_last_error(void);
/* Number of kernels this library has launched from the calling process (bench accounting). */
VSUM_API int64_t     vsum_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Frame scorer: replaces SimNet.forward (src/model/simnet.py:32-45) and everything under it
 * (Embedding 208-217, PositionalEncoding 236-238, Encoder 77-83, EncoderBlock 105-114,
 * MultiAttentionNetwork 138-164, MLP 180-183, final_layer 42) in eval mode, plus the sigmoid
 * the callers apply (src/train.py:144).
 * ------------------------------------------------------------------------------------------ */
typedef struct vsum_scorer *vsum_scorer_t;

typedef struct {
    int32_t d_model;       /* simnet.py:10 */
    int32_t num_heads;
    int32_t num_layers;
    int32_t d_ff;          /* 4 * d_model (simnet.py:99,175) */
    int32_t in_features;   /* 1024 (simnet.py:22) */
    int32_t num_classes;   /* 1 */
    int32_t use_pos;       /* simnet.py:212 */
    int32_t reserved;
} vsum_scorer_config;

/* fp32 device arrays laid out exactly as the reference state_dict tensors (nn.Linear weight is
 * [out,in] row-major).  pos_table is [pos_rows, d_model] (simnet.py:224-233). */
typedef struct {
    const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;   /* encoder.module_list.i.sa.* */
    const float *ln1_g, *ln1_b;                                    /* .norm1 */
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;                    /* .mlp.fc1 / .mlp.fc2 */
    const float *ln2_g, *ln2_b;                                    /* .norm2 */
} vsum_layer_weights;

typedef struct {
    const float *embed_w, *embed_b;      /* embedding_layer.feature_transform */
    const float *pos_table;              /* embedding_layer.positional_encoding.pos_embedding */
    int32_t      pos_rows;
    int32_t      reserved;
    const float *final_w, *final_b;      /* final_layer */
    vsum_layer_weights layers[VSUM_MAX_LAYERS];
} vsum_scorer_weights;

VSUM_API int vsum_scorer_create(vsum_scorer_t *out, const vsum_scorer_config *cfg_host);
VSUM_API int vsum_scorer_destroy(vsum_scorer_t h);
/* Copies (and for the bf16 path converts/packs) the weights into handle-owned device memory. */
VSUM_API int vsum_scorer_load_weights(vsum_scorer_t h, const vsum_scorer_weights *w_host, void *stream);
/* The same with flags.  VSUM_WEIGHTS_TRAIN_ONLY: refresh only what the training entry points read (fp32 copies and their
 * transposes) -- the per-step refresh after optimizer.step() (src/train.py:127); the bf16 inference path then refuses to
 * run until a full vsum_scorer_load_weights. */
#define VSUM_WEIGHTS_TRAIN_ONLY 1
VSUM_API int vsum_scorer_load_weights_ex(vsum_scorer_t h, const vsum_scorer_weights *w_host, int32_t flags, void *stream);
/* Bytes of scratch vsum_scorer_forward needs for T packed frames in B videos. */
VSUM_API size_t vsum_scorer_workspace_bytes(vsum_scorer_t h, int64_t T, int32_t B, int32_t mode);
/* features [T,in_features] fp32 (bf16 with VSUM_MODE_BF16_FEATURES), cu_seqlens int32[B+1].  Key masking follows simnet.py:156-157:
 * with packed videos no key is ever padded.  scores_out [T,num_classes] fp32 logits (or their
 * sigmoid when apply_sigmoid != 0, src/train.py:144); feats_out [T,d_model] fp32 or NULL
 * (the second element of SimNet.forward's tuple).  max_len = longest video in the batch. */
VSUM_API int vsum_scorer_forward(vsum_scorer_t h, const void *features, const int32_t *cu_seqlens,
                        int32_t B, int64_t T, int32_t max_len, int32_t mode, int32_t apply_sigmoid,
                        float *scores_out, float *feats_out, void *workspace,
                        size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Training step of the scorer: replaces autograd through SimNet.forward as used by
 * src/train.py:111-131 and src/pretrain.py:49-86, including the four dropout sites of
 * simnet.py:107,110,159,181 (counter-based masks recomputed from `seed` in the backward pass).
 *   forward_train keeps every activation the backward needs in `tape` (vsum_scorer_tape_bytes);
 *   backward writes d(loss)/d(parameter) for each tensor named in the vsum_scorer_grads struct [same fields
 *   as vsum_scorer_weights, fp32 device arrays of the parameter's shape; they are overwritten].  When
 *   q_w|k_w|v_w (and q_b|k_b|v_b) of a layer are adjacent in memory the fused QKV weight gradient is
 *   written in place, otherwise it goes through a scratch copy.
 *   d_feats [T,d_model] may be NULL (gradient w.r.t. the second forward output).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b, *ln1_g, *ln1_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b,
          *ln2_g, *ln2_b;
} vsum_layer_grads;
typedef struct {
    float *embed_w, *embed_b, *final_w, *final_b;
    vsum_layer_grads layers[VSUM_MAX_LAYERS];
    int32_t pre_zeroed;   /* non-zero: the caller has already zeroed every array (e.g. one memset of a flat buffer) */
} vsum_scorer_grads;

/* Training path arithmetic: mode 0 = fp32 SIMT kernels (reference accuracy); mode 1 = linear layers on
 * tcgen05 (tf32 MMA for forward and dgrad, bf16 MMA with fp32 accumulation for wgrad; needs d_model and
 * d_ff multiples of 256); mode 2 = mode 1 plus attention forward and backward on tcgen05 with bf16
 * operands (d_model 256, 4 heads) -- the counterpart of the reference's autocast (src/train.py:120). */
VSUM_API int vsum_scorer_set_train_mode(vsum_scorer_t h, int32_t mode);
VSUM_API size_t vsum_scorer_tape_bytes(vsum_scorer_t h, int64_t T);
VSUM_API size_t vsum_scorer_train_workspace_bytes(vsum_scorer_t h, int64_t T, int32_t B);
VSUM_API int vsum_scorer_forward_train(vsum_scorer_t h, const float *features, const int32_t *cu_seqlens,
                              int32_t B, int64_t T, int32_t max_len, float dropout_p, uint64_t seed,
                              float *scores_out, float *feats_out, void *tape, size_t tape_bytes,
                              void *workspace, size_t workspace_bytes, void *stream);
VSUM_API int vsum_scorer_backward(vsum_scorer_t h, const float *features, const int32_t *cu_seqlens, int32_t B,
                         int64_t T, int32_t max_len, float dropout_p, uint64_t seed, const float *d_scores,
                         const float *d_feats, const void *tape, const vsum_scorer_grads *grads_host,
                         void *workspace, size_t workspace_bytes, void *stream);
/* Data-parallel training (src/train.py:111-131 on several GPUs): the same backward, calling `hook(user, bucket)` on the
 * calling host thread as soon as every kernel that writes one group of gradients has been QUEUED on `stream` -- bucket =
 * num_layers-1 ... 0 for an encoder layer (the last layer's call also covers final_w / final_b), then -1 for the embedding
 * (everything is queued).  The caller records an event there and all-reduces that slice of its gradient buffer on another
 * stream while the backward of the earlier layers runs.
 *   vsum_dp_finalize: grads[n] *= 1 / D and loss_out = ext[0] / D with D = ext[1] * max(ext[2 .. 2 + world)), i.e.
 *   (sum of batch sizes) * (largest Nmax): the padded size mse_with_mask_loss (src/utils/utils.py:55) divides by for the
 *   global batch, from SUM-reduced extras -- no host round trip for the denominator.  vsum_dp_extras writes this rank's
 *   extras, ext[2 + world] = [*loss_sum, batch, Nmax one-hot at rank], before that all-reduce. */
typedef void (*vsum_grad_bucket_hook)(void *user, int32_t bucket);
VSUM_API int vsum_scorer_backward_hooked(vsum_scorer_t h, const float *features, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                         int32_t max_len, float drop_p, uint64_t seed, const float *d_scores, const float *d_feats,
                                         const void *tape, const vsum_scorer_grads *grads_host, void *workspace, size_t workspace_bytes,
                                         void *stream, vsum_grad_bucket_hook hook, void *hook_user);
VSUM_API int vsum_dp_extras(float *ext, const float *loss_sum, int32_t batch, int32_t nmax, int32_t rank, int32_t world, void *stream);
VSUM_API int vsum_dp_finalize(float *grads, int64_t n, const float *ext, int32_t world, float *loss_out, void *stream);
/* Masked MSE of src/utils/utils.py:45-56 over n = bs*Nmax entries: *loss_out (device, pre-zeroed) +=
 * sum(((out - tgt) * !pad)^2) / denom; d_out (optional) = grad_scale * d(loss)/d(out). */
VSUM_API int vsum_masked_mse(const float *out, const float *tgt, const uint8_t *pad_mask, int64_t n, float denom,
                    float *loss_out, float grad_scale, float *d_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Shot pooling: replaces generate_summary.py:25-46 (upsample by picks, per-shot float32 mean in
 * numpy's pairwise order widened to fp64, shot lengths, 15 % capacity).
 *   scores int32-indexed by cu_steps[B+1]; picks by cu_picks[B+1]; change points cps[S_total][2]
 *   (inclusive) by cu_shots[B+1]; n_frames[B].
 *   Outputs: val_out fp64[S_total], wt_out int32[S_total], cap_out int32[B].
 * ------------------------------------------------------------------------------------------ */
VSUM_API int vsum_shot_mean(const float *scores, const int32_t *cu_steps, const int32_t *picks,
                   const int32_t *cu_picks, const int32_t *n_frames, const int32_t *cps,
                   const int32_t *cu_shots, int32_t B, int32_t S_total, double *val_out,
                   int32_t *wt_out, int32_t *cap_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * 0/1 knapsack: replaces knapSack (src/evaluation/knapsack_implementation.py:1-30) for B videos
 * at once, one video per CTA.  fp64 DP row in shared memory, strict-greater decision bits,
 * back-track from w = capacity; ties go to the lower-indexed shot exactly as line 26 does.
 *   take_bits: scratch of sum(vsum_knapsack_scratch_words(S_v, cap_v)) 32-bit words; bit_offsets
 *   int64[B+1] are word offsets of each video's S x (class width / 32) decision matrix.
 *   order int32[B] (optional, may be NULL): video processing order (heaviest first).
 *   selected_out uint8[S_total]: 1 iff the shot is in the summary.
 * ------------------------------------------------------------------------------------------ */
/* Width (in capacities) of the kernel class that solves a video of this capacity, or -1 when the
 * fp64 row would not fit shared memory.  All videos of ONE vsum_knapsack call must belong to the
 * class of `max_cap`; decision-bit rows are padded to the class width. */
VSUM_API int32_t vsum_knapsack_class_width(int32_t capacity);
VSUM_API int64_t vsum_knapsack_scratch_words(int32_t n_shots, int32_t capacity);
VSUM_API int vsum_knapsack(const double *val, const int32_t *wt, const int32_t *cu_shots,
                  const int32_t *cap, const int64_t *bit_offsets, const int32_t *order, int32_t B,
                  int32_t max_cap, uint32_t *take_bits, uint8_t *selected_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Summary mask + keyshot F-score: replaces generate_summary.py:51-53 and evaluate_summary
 * (src/evaluation/evaluation_metrics.py:4-33).
 *   user_summary: float32 (as the h5 files hold it) or uint8 (the packed dataset's lossless form of 0/1 rows),
 *   per user_summary_dtype; video v holds n_users[v] rows of us_cols[v] columns starting at
 *   element us_offsets[v] (int64[B+1]).  summary_out int8: video v occupies exactly
 *   [sum_offsets[v], sum_offsets[v+1]) (int64[B+1]), i.e. last_shot_end+1 entries.  With
 *   selected == NULL, summary_out is an INPUT holding the masks (evaluate_summary on its own).  counts_ws: scratch of 3*sum(n_users) int64.
 *   f_out fp64[B]; per_user_out fp64[sum(n_users)] or NULL.  cu_users int32[B+1].
 * ------------------------------------------------------------------------------------------ */
enum { VSUM_USER_SUMMARY_F32 = 0, VSUM_USER_SUMMARY_U8 = 1 };   /* the values vsum_pack_info.user_summary_dtype takes */
VSUM_API int vsum_summary_fscore(const uint8_t *selected, const int32_t *cps, const int32_t *cu_shots,
                        const void *user_summary, int32_t user_summary_dtype, const int64_t *us_offsets,
                        const int32_t *cu_users, const int32_t *us_cols, int32_t B,
                        int32_t total_users, int32_t method, int8_t *summary_out,
                        const int64_t *sum_offsets, int64_t summary_total, int64_t *counts_ws,
                        double *f_out, double *per_user_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * SM partition for pipelined execution (scorer of batch i+1 on one stream, evaluation of batch i
 * on another): the knapsack / overlap kernels run as at most `eval_sms` SMs' worth of persistent
 * CTAs, and the persistent scorer GEMMs launch on (SM count - scorer_reserved_sms) CTAs, so that
 * neither stage waits for an SM the other holds.  0 / 0 (the default) removes both limits.
 * ------------------------------------------------------------------------------------------ */
VSUM_API int vsum_set_sm_partition(int32_t eval_sms, int32_t scorer_reserved_sms);

/* ------------------------------------------------------------------------------------------
 * Per-kernel timing (bench.py's roofline leg): between begin and end every kernel launch of
 * this library is bracketed by CUDA events on its own stream; end() synchronises those events
 * and returns summed milliseconds and launch counts per category.
 * ------------------------------------------------------------------------------------------ */
VSUM_API int vsum_profile_begin(void);
VSUM_API int vsum_profile_end(float *ms_out_host, int32_t *count_out_host, int32_t ncat);
VSUM_API int32_t vsum_profile_num_categories(void);
VSUM_API const char *vsum_profile_category_name(int32_t i);

/* Frame numbers of the summary (src/generate_summary_image.py:74-76, the lists written to summary.json):
 * ascending indices of the frames inside selected shots.  frames_out holds video v at
 * [out_offsets[v], out_offsets[v+1]) (int64[B+1]; the knapsack capacity of the video is always enough
 * room), counts_out[v] = how many were written. */
VSUM_API int vsum_summary_frames(const uint8_t *selected, const int32_t *change_points, const int32_t *cu_shots,
                                 const int64_t *out_offsets, int32_t B, int32_t *frames_out, int32_t *counts_out,
                                 void *stream);

/* ------------------------------------------------------------------------------------------
 * Rank correlations (src/evaluation/compute_correlation.py:4-15, called from compute_metrics.py:82-85):
 * per video the mean over users of Kendall tau-b and Spearman rho between the predicted frame scores
 * (the sub-sampled scores repeated up to the next pick, compute_metrics.py:19-39 -- never materialised)
 * and each user's frame scores, with scipy's tie handling.  tau is bit-exact with scipy.stats.kendalltau
 * (exact integer pair counts, scipy's expression order); rho comes from exact integer rank sums and
 * agrees with scipy.stats.spearmanr to ~1e-15.  Constant inputs give NaN like scipy.
 *   scores float32[T] packed, cu_steps int32[B+1]; picks int32 (one per step), n_frames int32[B];
 *   user_scores float32: video v holds n_users[v] rows of us_cols[v] (== n_frames[v]) columns starting at
 *   element us_offsets[v] (int64[B+1]); cu_users int32[B+1].  At most 8191 steps per video.
 *   kendall_out / spearman_out fp64[B]; per_user_* fp64[sum(n_users)] or NULL.
 *   workspace: 1024-byte aligned device memory of vsum_rank_correlation_workspace_bytes() bytes
 *   (16 bytes per user-score element).
 * ------------------------------------------------------------------------------------------ */
VSUM_API size_t vsum_rank_correlation_workspace_bytes(int64_t total_user_elems, int64_t T, int32_t B,
                                                      int32_t total_users);
VSUM_API int vsum_rank_correlation(const float *scores, const int32_t *cu_steps, const int32_t *picks,
                                   const int32_t *n_frames, const float *user_scores, const int64_t *us_offsets,
                                   const int32_t *cu_users, const int32_t *us_cols, int32_t B, int64_t T,
                                   int32_t max_steps, int32_t total_users, int64_t total_user_elems,
                                   void *workspace, size_t workspace_bytes, double *kendall_out,
                                   double *spearman_out, double *per_user_tau, double *per_user_rho, void *stream);

/* ------------------------------------------------------------------------------------------
 * Pretraining head (src/model/simnet_pretrain.py:33, 35-100; src/pretrain.py:59-67).
 *
 * vsum_linear_*: a stand-alone Linear layer on packed rows (PretrainModel.video_transform,
 *   simnet_pretrain.py:33,80): y[M,N] = x[M,K] w[N,K]^T + bias; backward overwrites dw, db and, when
 *   non-NULL, dx.  mode 0 = fp32 SIMT, mode 1 = tcgen05 (tf32 forward / dgrad, bf16 wgrad; N, K
 *   multiples of 256, else it falls back to mode 0).  The backward workspace (1024-byte aligned,
 *   the bytes vsum_linear_workspace_bytes returns) is only needed in mode 1.
 *
 * vsum_pretrain_losses_*: the three losses of PretrainModel.forward (simnet_pretrain.py:82-100) and
 *   their gradients on the packed layout.  scores [T] = encoder logits, x512 [T,512] = video_transform
 *   output, video_rep [B,512] = the target representation (no gradient), n_pad = padded length Nmax of
 *   the reference's batch (its means divide by it), pen_entropy 1 = "entropy" centering (line 90), 0 =
 *   the norm variant (line 94).  losses3 (device) = {distillation, center, repel}.  `saved` keeps the
 *   mixture weights and per-video sums for the backward, which takes d_losses3 (device, the three
 *   upstream gradients) and writes d_scores [T] and d_x512 [T,512].  Reductions are order-fixed
 *   (no atomics); the repel term uses |sum xh|^2 - sum |xh|^2, so no [N,N] tensor exists.
 * ------------------------------------------------------------------------------------------ */
VSUM_API size_t vsum_linear_workspace_bytes(int64_t M, int32_t N, int32_t K);
VSUM_API int vsum_linear_forward(const float *x, const float *w, const float *bias, float *y, int64_t M, int32_t N,
                                 int32_t K, int32_t mode, void *stream);
VSUM_API int vsum_linear_backward(const float *dy, const float *x, const float *w, float *dx, float *dw, float *db,
                                  int64_t M, int32_t N, int32_t K, int32_t mode, void *workspace,
                                  size_t workspace_bytes, void *stream);
VSUM_API size_t vsum_pretrain_saved_bytes(int64_t T, int32_t B, int32_t max_len);
VSUM_API int vsum_pretrain_losses_forward(const float *scores, const float *x512, const int32_t *cu_seqlens,
                                          int32_t B, int64_t T, int32_t max_len, int32_t n_pad, float sharpening_t,
                                          const float *video_rep, int32_t pen_entropy, float *losses3,
                                          void *saved, size_t saved_bytes, void *stream);
VSUM_API int vsum_pretrain_losses_backward(const float *x512, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                           int32_t max_len, int32_t n_pad, float sharpening_t, int32_t pen_entropy,
                                           const float *d_losses3, void *saved, float *d_scores, float *d_x512,
                                           void *stream);

/* ------------------------------------------------------------------------------------------
 * Kernel temporal segmentation (src/data/preprocess/segmentations/kts/cpd_nonlin.py:5-91, cpd_auto.py:5-44):
 * the change-point detection that produces the shot boundaries.  vsum_kts_gram: K = X X^T in fp32
 * (create_segments.py:44; zeros_n = n floats of scratch).  vsum_kts_dp: scatters + dynamic programme for
 * 0..m change points; scores_out fp64[m+1] = I[:, n], prev_out int32[(m+1)(n+1)] = the back-pointer table p
 * (both on the device).  Given the same K, scores and back-pointers are bit-identical to the reference
 * (float32 cumulative sums, fp64 scatters in its expression order, first-minimum rule); the host wrapper
 * applies the penalty of cpd_auto.py:30-36 and back-tracks.
 * ------------------------------------------------------------------------------------------ */
VSUM_API size_t vsum_kts_workspace_bytes(int32_t n, int32_t m);
VSUM_API int vsum_kts_gram(const float *features, int32_t n, int32_t dim, float *zeros_n, float *K_out, void *stream);
VSUM_API int vsum_kts_dp(const float *K, int32_t n, int32_t m, int32_t lmin, int32_t lmax, void *workspace,
                         size_t workspace_bytes, double *scores_out, int32_t *prev_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Data layer (host side; replaces the h5py / pad_sequence input path of src/data/dataset.py:64-168 and
 * src/train.py:115-118): a packed, memory-mapped dataset file (written by vsum_b200/data/packed.py) and a
 * multi-threaded padding-free collate that gathers a batch straight into one caller buffer (pinned
 * host memory) as packed rows + cu_seqlens -- the layout vsum_scorer_forward consumes.
 *   vsum_pack_array returns a zero-copy view into the mapping (NULL / 0 bytes when the array is absent).
 *   vsum_pack_collate: features_out [sum N, feature_dim] in the pack's feature dtype (vsum_pack_feature_dtype:
 *   float32 as in the h5 files, or bfloat16 for VSUM_MODE_BF16_FEATURES), gtscore_out [sum N] or NULL,
 *   cu_seqlens_out int32[n+1]; `threads` host threads split the bytes evenly.
 * ------------------------------------------------------------------------------------------ */
enum { VSUM_PACK_FEATURES = 0, VSUM_PACK_GTSCORE, VSUM_PACK_PICKS, VSUM_PACK_CHANGE_POINTS, VSUM_PACK_USER_SUMMARY,
       VSUM_PACK_USER_SCORES, VSUM_PACK_VIDEO_REP, VSUM_PACK_NUM_ARRAYS };
enum { VSUM_FEATURES_F32 = 0, VSUM_FEATURES_BF16 = 1 };
typedef struct vsum_pack *vsum_pack_t;
typedef struct {
    char name[96];
    int32_t n_steps, n_frames, n_shots, n_users, rep_dim, has_user_scores;
    int32_t user_summary_dtype;   /* 0 = float32 (as in the h5 files), 1 = uint8 */
} vsum_pack_info;
VSUM_API int vsum_pack_open(const char *path, vsum_pack_t *out);
VSUM_API void vsum_pack_close(vsum_pack_t pack);
VSUM_API int32_t vsum_pack_num_videos(vsum_pack_t pack);
VSUM_API int32_t vsum_pack_feature_dim(vsum_pack_t pack);
VSUM_API int32_t vsum_pack_feature_dtype(vsum_pack_t pack);
VSUM_API int vsum_pack_video_info(vsum_pack_t pack, int32_t video, vsum_pack_info *out);
VSUM_API int vsum_pack_array(vsum_pack_t pack, int32_t video, int32_t kind, const void **ptr, uint64_t *bytes);
VSUM_API int vsum_pack_collate(vsum_pack_t pack, const int32_t *videos, int32_t n, int32_t threads,
                               void *features_out, float *gtscore_out, int32_t *cu_seqlens_out);

/* Evaluation-side loader (replaces the val DataLoader + the per-video loop of src/train.py:139-148 and the h5 reads of
 * src/data/dataset.py:89-103,127-135 for MANY videos per step).
 *   vsum_pack_open_ex(..., VSUM_PACK_PINNED): the file is read ONCE into page-locked host memory instead of being
 *     mapped, so every later batch is DMA'd straight out of the dataset -- no host-side collate copy of the big arrays.
 *   vsum_pack_eval_collate: everything the evaluation kernels need for a batch EXCEPT the two big arrays, written into
 *     one caller blob (pinned host memory; one H2D copy moves it): the videos are ordered longest first (stable), then
 *     cu_steps, picks, cu_picks, n_frames, change points, cu_shots, knapsack decision-bit offsets, the knapsack order
 *     (largest capacity first) and its per-class launches, summary offsets, user-summary offsets / rows / columns.
 *     `blob_host` == NULL: only the layout (sizes, offsets) is computed.  Offsets are bytes from the blob's start,
 *     every array 256-byte aligned.  Capacity = (int)((last_end + 1) * 0.15) as generate_summary.py:45-46.
 *   vsum_pack_h2d: one cudaMemcpyAsync per video and array: features to features_dev + cu_steps[k] rows, user
 *     summaries to user_summary_dev + us_offsets[k] elements (either may be NULL), in the blob's video order.
 *     Needs a pack opened with VSUM_PACK_PINNED (fails otherwise: a pageable source would serialise the stream). */
enum { VSUM_PACK_MMAP = 0, VSUM_PACK_PINNED = 1 };
typedef struct {
    int32_t B, total_users, total_shots, max_steps, max_cap, n_launches, user_summary_dtype, reserved;
    int64_t T, total_picks, summary_frames, bit_words, us_elems, blob_bytes;
    int64_t off_video_ids, off_cu_steps, off_picks, off_cu_picks, off_n_frames, off_cps, off_cu_shots, off_bit_offsets,
            off_order, off_sum_offsets, off_us_offsets, off_cu_users, off_us_cols;
    int32_t launch_first[8], launch_count[8], launch_max_cap[8];
} vsum_eval_batch_layout;
VSUM_API int vsum_pack_open_ex(const char *path, int32_t residency, vsum_pack_t *out);
VSUM_API int32_t vsum_pack_residency(vsum_pack_t pack);
VSUM_API int vsum_pack_eval_collate(vsum_pack_t pack, const int32_t *videos, int32_t n, void *blob_host, size_t blob_bytes,
                                    vsum_eval_batch_layout *layout_out);
VSUM_API int vsum_pack_h2d(vsum_pack_t pack, const void *blob_host, const vsum_eval_batch_layout *layout,
                           void *features_dev, void *user_summary_dev, void *stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics: the two tcgen05 kernels on their own, so tests can pin them individually.
 *   vsum_debug_gemm_tc05: out[M,N] bf16 = epi(A[M,K] W[N,K]^T + bias); A/W bf16, or fp32 when
 *     a_is_f32 (tf32 MMA).  epi: 0 bias, 1 bias+ReLU, 3 bias+residual+LayerNorm (N == 256),
 *     5 / 6: bias (+ReLU) with an fp32 output (tf32 operands only).
 *   vsum_debug_attention_tc05: qkv [T,768] bf16 -> out [T,256] bf16 (4 heads of 64, scale 1/16);
 *     scratch_i32 holds vsum_attention_scratch_ints(T, B) int32 (all three attention entry points).
 *   vsum_set_attention_kernel: which forward kernel the scorer and these entry points run --
 *     2 (default): persistent kernel, two 128-query tiles per CTA, probabilities in tensor memory
 *     (csrc/vsum_attn2_tc05.cu); 3: the same kernel with TWO softmax threads per query row (16 softmax warps) on the
 *     pre-scaled inference fast pass; 1: one 128-query tile per CTA (csrc/vsum_attn_tc05.cu).  All replace
 *     src/model/simnet.py:155-161.  The environment variable VSUM_ATTN_KERNEL sets the initial value.
 * ------------------------------------------------------------------------------------------ */
VSUM_API int vsum_set_attention_kernel(int32_t version);
/* Feed-forward block of an encoder layer (src/model/simnet.py:180-183 + 109-110): 2 (default) = ONE kernel for
 * fc1 + ReLU + fc2 + residual + LayerNorm (+ regression head on the last layer), the [T,1024] hidden rows never leave the
 * SM (csrc/vsum_ffn_tc05.cu); 1 = two GEMM launches with the hidden rows in HBM.  VSUM_FFN_KERNEL sets the initial value.
 *   vsum_debug_ffn_tc05: out[M,256] bf16 = LayerNorm(relu(x W1^T + b1) W2^T + b2 + x) * gamma + beta; x [M,256],
 *   W1 [1024,256], W2 [256,1024] bf16, biases / gamma / beta fp32. */
VSUM_API int vsum_set_ffn_kernel(int32_t version);
VSUM_API int vsum_debug_ffn_tc05(const void *x_bf16, const void *w1_bf16, const float *b1, const void *w2_bf16, const float *b2,
                                 const float *gamma, const float *beta, void *out_bf16, int64_t M, void *stream);
VSUM_API size_t vsum_attention_scratch_ints(int64_t T, int32_t B);
VSUM_API int vsum_debug_gemm_tc05(const void *A, const void *W, const float *bias, const void *residual_bf16,
                         const float *gamma, const float *beta, void *out_bf16, int64_t M, int32_t N,
                         int32_t K, int32_t a_is_f32, int32_t epi, void *stream);
/* dW[N,K] += dY[M,N]^T X[M,K] on tcgen05 (operands rounded to bf16 into scratch_bf16, M*(N+K) elements,
 * both MN-major; fp32 accumulation), db[N] += colsum(dY) or NULL. */
VSUM_API int vsum_debug_wgrad_tc05(const float *dY, const float *X, float *dW, float *db, int64_t M, int32_t N,
                          int32_t K, void *scratch_bf16, void *stream);
/* Training variant of the tcgen05 attention: FP32 output [T,256], also writes the log2-domain
 * log-sum-exp [T,4] and applies dropout to P; and its backward (qkv, d_out bf16; lse2, delta [T,4] fp32;
 * delta = rowsum per head of out * d_out) -> dqkv [T,768] fp32. */
VSUM_API int vsum_debug_attention_train_tc05(const void *qkv_bf16, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                             float *out_f32, float *lse2, float drop_p, uint64_t seed,
                                             int32_t *scratch, void *stream);
VSUM_API int vsum_debug_attention_bwd_tc05(const void *qkv_bf16, const void *d_out_bf16, const float *lse2,
                                           const float *delta, const int32_t *cu_seqlens, int32_t B, int64_t T,
                                           float drop_p, uint64_t seed, float *dqkv, int32_t *scratch, void *stream);
VSUM_API int vsum_debug_attention_tc05(const void *qkv_bf16, const int32_t *cu_seqlens, int32_t B, int64_t T,
                              void *out_bf16, int32_t *scratch_i32, void *stream);
/* Same with an explicit softmax scale: P = softmax(scale * Q K^T).  scale = 1 / log2(e) is the form the scorer runs
 * (it folds d_model^-0.5 * log2(e) into its bf16 copy of W_q, so Q K^T already is the base-2 exponent): the two-tile
 * kernel then exponentiates the scores as they come out of tensor memory. */
VSUM_API int vsum_debug_attention_scaled_tc05(const void *qkv_bf16, const int32_t *cu_seqlens, int32_t B, int64_t T, float scale,
                                              void *out_bf16, int32_t *scratch_i32, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VSUM_B200_H */
