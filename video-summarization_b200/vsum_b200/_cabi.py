"""ctypes binding of `libvsum_b200.so` (declared in `include/vsum_b200.h`).

The library is built in-tree by `__graft_entry__.build()` / `make -C video-summarization_b200/csrc`.
There is no fallback of any kind: if the shared object is missing, or a call fails, this module
raises -- the product path never routes through PyTorch eager ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSUM_LIB", os.path.join(_HERE, "libvsum_b200.so"))   # VSUM_LIB: debug builds only

VSUM_MAX_LAYERS = 16
MODE_FP32, MODE_BF16, MODE_BF16_FEATURES = 0, 1, 2
WEIGHTS_TRAIN_ONLY = 1
FSCORE_AVG, FSCORE_MAX = 0, 1
USER_SUMMARY_F32, USER_SUMMARY_U8 = 0, 1
FEATURES_F32, FEATURES_BF16 = 0, 1

# every symbol include/vsum_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = (
    "vsum_abi_version", "vsum_last_error", "vsum_launch_count", "vsum_set_sm_partition",
    "vsum_scorer_create", "vsum_scorer_destroy", "vsum_scorer_load_weights", "vsum_scorer_load_weights_ex",
    "vsum_scorer_workspace_bytes", "vsum_scorer_forward",
    "vsum_shot_mean", "vsum_knapsack_class_width", "vsum_knapsack_scratch_words", "vsum_knapsack", "vsum_summary_fscore",
    "vsum_scorer_set_train_mode", "vsum_scorer_tape_bytes", "vsum_scorer_train_workspace_bytes", "vsum_scorer_forward_train",
    "vsum_scorer_backward", "vsum_scorer_backward_hooked", "vsum_dp_extras", "vsum_dp_finalize", "vsum_masked_mse",
    "vsum_debug_gemm_tc05", "vsum_debug_wgrad_tc05", "vsum_debug_attention_tc05", "vsum_debug_attention_scaled_tc05", "vsum_set_attention_kernel", "vsum_attention_scratch_ints",
    "vsum_set_ffn_kernel", "vsum_debug_ffn_tc05",
    "vsum_debug_attention_train_tc05", "vsum_debug_attention_bwd_tc05",
    "vsum_linear_workspace_bytes", "vsum_linear_forward", "vsum_linear_backward",
    "vsum_kts_workspace_bytes", "vsum_kts_gram", "vsum_kts_dp",
    "vsum_pack_open", "vsum_pack_close", "vsum_pack_num_videos", "vsum_pack_feature_dim", "vsum_pack_feature_dtype", "vsum_pack_video_info",
    "vsum_pack_array", "vsum_pack_collate", "vsum_pack_open_ex", "vsum_pack_residency", "vsum_pack_eval_collate", "vsum_pack_h2d",
    "vsum_summary_frames", "vsum_rank_correlation_workspace_bytes", "vsum_rank_correlation",
    "vsum_pretrain_saved_bytes", "vsum_pretrain_losses_forward", "vsum_pretrain_losses_backward",
    "vsum_profile_begin", "vsum_profile_end", "vsum_profile_num_categories", "vsum_profile_category_name",
)


class VsumError(RuntimeError):
    pass


class ScorerConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("d_model", "num_heads", "num_layers", "d_ff", "in_features",
                                          "num_classes", "use_pos", "reserved")]


_LAYER_FIELDS = ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b", "ln1_g", "ln1_b",
                 "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ln2_g", "ln2_b")


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class ScorerWeights(C.Structure):
    _fields_ = [("embed_w", C.c_void_p), ("embed_b", C.c_void_p), ("pos_table", C.c_void_p),
                ("pos_rows", C.c_int32), ("reserved", C.c_int32),
                ("final_w", C.c_void_p), ("final_b", C.c_void_p),
                ("layers", LayerWeights * VSUM_MAX_LAYERS)]


class LayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class ScorerGrads(C.Structure):
    _fields_ = [("embed_w", C.c_void_p), ("embed_b", C.c_void_p), ("final_w", C.c_void_p), ("final_b", C.c_void_p),
                ("layers", LayerGrads * VSUM_MAX_LAYERS), ("pre_zeroed", C.c_int32)]


GRAD_BUCKET_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32)      # vsum_grad_bucket_hook


class PackInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("n_steps", C.c_int32), ("n_frames", C.c_int32), ("n_shots", C.c_int32),
                ("n_users", C.c_int32), ("rep_dim", C.c_int32), ("has_user_scores", C.c_int32), ("user_summary_dtype", C.c_int32)]


class EvalBatchLayout(C.Structure):
    """vsum_eval_batch_layout (include/vsum_b200.h)."""
    _fields_ = ([(n, C.c_int32) for n in ("B", "total_users", "total_shots", "max_steps", "max_cap", "n_launches",
                                           "user_summary_dtype", "reserved")] +
                [(n, C.c_int64) for n in ("T", "total_picks", "summary_frames", "bit_words", "us_elems", "blob_bytes",
                                           "off_video_ids", "off_cu_steps", "off_picks", "off_cu_picks", "off_n_frames", "off_cps",
                                           "off_cu_shots", "off_bit_offsets", "off_order", "off_sum_offsets", "off_us_offsets",
                                           "off_cu_users", "off_us_cols")] +
                [("launch_first", C.c_int32 * 8), ("launch_count", C.c_int32 * 8), ("launch_max_cap", C.c_int32 * 8)])


PACK_MMAP, PACK_PINNED = 0, 1
PACK_FEATURES, PACK_GTSCORE, PACK_PICKS, PACK_CHANGE_POINTS, PACK_USER_SUMMARY, PACK_USER_SCORES, PACK_VIDEO_REP = range(7)

_lib = None


def load():
    """Load the shared object once.  Raises VsumError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VsumError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C video-summarization_b200/csrc`.  vsum_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.vsum_abi_version.restype = C.c_int
    L.vsum_last_error.restype = C.c_char_p
    L.vsum_launch_count.restype = i64
    L.vsum_set_sm_partition.argtypes = [i32, i32]
    L.vsum_scorer_create.argtypes = [C.POINTER(vp), C.POINTER(ScorerConfig)]
    L.vsum_scorer_destroy.argtypes = [vp]
    L.vsum_scorer_load_weights.argtypes = [vp, C.POINTER(ScorerWeights), vp]
    L.vsum_scorer_load_weights_ex.argtypes = [vp, C.POINTER(ScorerWeights), i32, vp]
    L.vsum_scorer_workspace_bytes.restype = C.c_size_t
    L.vsum_scorer_workspace_bytes.argtypes = [vp, i64, i32, i32]
    L.vsum_scorer_forward.argtypes = [vp, vp, vp, i32, i64, i32, i32, i32, vp, vp, vp, C.c_size_t, vp]
    L.vsum_scorer_set_train_mode.argtypes = [vp, i32]
    L.vsum_debug_wgrad_tc05.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, vp]
    L.vsum_scorer_tape_bytes.restype = C.c_size_t
    L.vsum_scorer_tape_bytes.argtypes = [vp, i64]
    L.vsum_scorer_train_workspace_bytes.restype = C.c_size_t
    L.vsum_scorer_train_workspace_bytes.argtypes = [vp, i64, i32]
    L.vsum_scorer_forward_train.argtypes = [vp, vp, vp, i32, i64, i32, C.c_float, C.c_uint64, vp, vp, vp, C.c_size_t,
                                            vp, C.c_size_t, vp]
    L.vsum_scorer_backward.argtypes = [vp, vp, vp, i32, i64, i32, C.c_float, C.c_uint64, vp, vp, vp,
                                       C.POINTER(ScorerGrads), vp, C.c_size_t, vp]
    L.vsum_scorer_backward_hooked.argtypes = [vp, vp, vp, i32, i64, i32, C.c_float, C.c_uint64, vp, vp, vp,
                                              C.POINTER(ScorerGrads), vp, C.c_size_t, vp, GRAD_BUCKET_HOOK, vp]
    L.vsum_dp_extras.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.vsum_dp_finalize.argtypes = [vp, i64, vp, i32, vp, vp]
    L.vsum_masked_mse.argtypes = [vp, vp, vp, i64, C.c_float, vp, C.c_float, vp, vp]
    L.vsum_shot_mean.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.vsum_knapsack_class_width.restype = i32
    L.vsum_knapsack_class_width.argtypes = [i32]
    L.vsum_knapsack_scratch_words.restype = i64
    L.vsum_knapsack_scratch_words.argtypes = [i32, i32]
    L.vsum_knapsack.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.vsum_summary_fscore.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, vp, vp, i64, vp, vp, vp, vp]
    L.vsum_debug_gemm_tc05.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    L.vsum_debug_attention_tc05.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    L.vsum_debug_attention_scaled_tc05.argtypes = [vp, vp, i32, i64, C.c_float, vp, vp, vp]
    L.vsum_set_attention_kernel.argtypes = [i32]
    L.vsum_set_ffn_kernel.argtypes = [i32]
    L.vsum_debug_ffn_tc05.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.vsum_attention_scratch_ints.argtypes = [i64, i32]
    L.vsum_attention_scratch_ints.restype = C.c_size_t
    L.vsum_debug_attention_train_tc05.argtypes = [vp, vp, i32, i64, vp, vp, C.c_float, C.c_uint64, vp, vp]
    L.vsum_debug_attention_bwd_tc05.argtypes = [vp, vp, vp, vp, vp, i32, i64, C.c_float, C.c_uint64, vp, vp, vp]
    L.vsum_linear_workspace_bytes.restype = C.c_size_t
    L.vsum_linear_workspace_bytes.argtypes = [i64, i32, i32]
    L.vsum_linear_forward.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, vp]
    L.vsum_linear_backward.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, C.c_size_t, vp]
    L.vsum_pretrain_saved_bytes.restype = C.c_size_t
    L.vsum_pretrain_saved_bytes.argtypes = [i64, i32, i32]
    L.vsum_pretrain_losses_forward.argtypes = [vp, vp, vp, i32, i64, i32, i32, C.c_float, vp, i32, vp, vp, C.c_size_t, vp]
    L.vsum_pretrain_losses_backward.argtypes = [vp, vp, i32, i64, i32, i32, C.c_float, i32, vp, vp, vp, vp, vp]
    L.vsum_kts_workspace_bytes.restype = C.c_size_t
    L.vsum_kts_workspace_bytes.argtypes = [i32, i32]
    L.vsum_kts_gram.argtypes = [vp, i32, i32, vp, vp, vp]
    L.vsum_kts_dp.argtypes = [vp, i32, i32, i32, i32, vp, C.c_size_t, vp, vp, vp]
    L.vsum_pack_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.vsum_pack_open_ex.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.vsum_pack_residency.argtypes = [vp]
    L.vsum_pack_residency.restype = i32
    L.vsum_pack_eval_collate.argtypes = [vp, vp, i32, vp, C.c_size_t, C.POINTER(EvalBatchLayout)]
    L.vsum_pack_h2d.argtypes = [vp, vp, C.POINTER(EvalBatchLayout), vp, vp, vp]
    L.vsum_pack_close.argtypes = [vp]
    L.vsum_pack_close.restype = None
    L.vsum_pack_num_videos.argtypes = [vp]
    L.vsum_pack_num_videos.restype = i32
    L.vsum_pack_feature_dim.argtypes = [vp]
    L.vsum_pack_feature_dim.restype = i32
    L.vsum_pack_feature_dtype.argtypes = [vp]
    L.vsum_pack_feature_dtype.restype = i32
    L.vsum_pack_video_info.argtypes = [vp, i32, C.POINTER(PackInfo)]
    L.vsum_pack_array.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.vsum_pack_collate.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.vsum_summary_frames.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    L.vsum_rank_correlation_workspace_bytes.restype = C.c_size_t
    L.vsum_rank_correlation_workspace_bytes.argtypes = [i64, i64, i32, i32]
    L.vsum_rank_correlation.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i64, vp, C.c_size_t, vp, vp, vp, vp, vp]
    L.vsum_profile_end.argtypes = [vp, vp, i32]
    L.vsum_profile_num_categories.restype = i32
    L.vsum_profile_category_name.restype = C.c_char_p
    L.vsum_profile_category_name.argtypes = [i32]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("vsum_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vsum_last_error()
        raise VsumError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().vsum_launch_count())


def profile_begin() -> None:
    check(load().vsum_profile_begin(), "vsum_profile_begin")


def profile_end() -> dict:
    """{category: (total milliseconds, launches)} for every kernel launched since profile_begin()."""
    L = load()
    n = int(L.vsum_profile_num_categories())
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    check(L.vsum_profile_end(ms, cnt, n), "vsum_profile_end")
    return {L.vsum_profile_category_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
