"""ctypes prototypes for every symbol declared in include/mmb200.h."""
from __future__ import annotations

import ctypes as C

vp, ll, i32, f32, f64 = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_double

PROTOTYPES = {
    "mmb_version": (i32, []),
    "mmb_gemm_bf16": (i32, [vp, ll, i32, vp, ll, i32, vp, ll, vp, ll, i32, i32, i32, i32, i32, f32, vp, vp, ll, i32, i32, vp, vp]),
    "mmb_gemm_set_mode": (i32, [i32, i32]),
    "mmb_gemm_ce_num_parts": (i32, [i32]),
    "mmb_gemm_ce_stats": (i32, [vp, ll, vp, ll, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp]),
    "mmb_gemm_ce_stats_labels": (i32, [vp, ll, vp, ll, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp]),
    "mmb_ce_labels_reduce": (i32, [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp]),
    "mmb_ce_stats_reduce": (i32, [vp, i32, i32, vp, i32, i32, f32, f32, vp, vp, vp, vp, vp]),
    "mmb_gemm_ce_grad": (i32, [vp, ll, vp, ll, i32, i32, i32, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp, i32, i32, vp, ll, vp]),
    "mmb_cast_f32_to_bf16": (i32, [vp, vp, ll, vp]),
    "mmb_cast_bf16_to_f32": (i32, [vp, vp, ll, vp]),
    "mmb_im2col_patches": (i32, [vp, vp, ll, i32, i32, i32, i32, vp]),
    "mmb_add_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    "mmb_vit_embed_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    "mmb_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "mmb_vit_embed_ln_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "mmb_batch_sum": (i32, [vp, vp, i32, ll, i32, vp]),
    "mmb_colsum_bf16": (i32, [vp, vp, i32, i32, ll, vp]),
    "mmb_text_embed_fwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmb_text_embed_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "mmb_argmax_tokens": (i32, [vp, vp, i32, i32, vp]),
    "mmb_l2norm_fwd": (i32, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "mmb_l2norm_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "mmb_adamw_step": (i32, [vp, vp, vp, vp, vp, ll, f32, f32, f32, f32, f32, i32, f32, i32, vp]),
    "mmb_anyprecision_adamw_step": (i32, [vp, vp, vp, i32, vp, i32, vp, i32, vp, ll, f64, f64, f64, f64, f64, i32, f32, i32, vp]),
    "mmb_act_fwd": (i32, [vp, vp, ll, i32, vp]),
    "mmb_memset_async": (i32, [vp, i32, ll, vp]),
    "mmb_attention_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "mmb_attention_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "mmb_attention_bwd_launches": (i32, [i32]),
    "mmb_attention_fwd_kmask": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "mmb_attention_probs": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "mmb_bert_embed_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, i32, i32, i32, i32, f32, vp]),
    "mmb_vit_assemble_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "mmb_gather_rows_cast": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "mmb_gather_rows_idx_cast": (i32, [vp, ll, vp, vp, i32, i32, vp]),
    "mmb_tanh_inplace": (i32, [vp, ll, vp]),
    "mmb_concat_tokens": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmb_coca_text_embed_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmb_attention_fwd_generic": (i32, [vp, ll, ll, vp, ll, ll, vp, ll, ll, vp, ll, ll, vp, ll, ll, i32, i32, i32, i32, i32,
                                        i32, f32, vp]),
    "mmb_ce_labels": (i32, [vp, ll, vp, ll, ll, i32, i32, vp, vp, vp]),
    "mmb_contrastive_ce_stats": (i32, [vp, ll, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp, ll, vp, vp]),
    "mmb_contrastive_ce_grad": (i32, [vp, ll, vp, i32, i32, i32, f32, f32, vp, vp, i32, i32, vp, vp, ll, vp, vp, vp]),
    "mmb_matmul_f32": (i32, [vp, ll, i32, vp, ll, i32, vp, ll, i32, i32, i32, f32, i32, vp]),
    "mmb_symm_alloc": (i32, [ll, vp]),
    "mmb_symm_free": (i32, [vp]),
    "mmb_symm_get_handle": (i32, [vp, vp]),
    "mmb_symm_open_handle": (i32, [vp, vp]),
    "mmb_symm_close_handle": (i32, [vp]),
    "mmb_symm_signal_wait": (i32, [vp, vp, i32, i32, i32, vp]),
    "mmb_sum_scale": (i32, [vp, i32, f32, vp, i32, vp]),
    "mmb_attention_bwd_kmask": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "mmb_bert_embed_ln_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "mmb_vit_assemble_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mmb_split_tokens_cast": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mmb_tanh_bwd": (i32, [vp, vp, vp, vp, ll, vp]),
    "mmb_scatter_rows_add": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "mmb_scatter_rows_idx_add": (i32, [vp, vp, vp, ll, i32, i32, vp]),
    "mmb_ce_labels_bwd": (i32, [vp, ll, vp, ll, ll, i32, i32, vp, f32, vp, vp, ll, vp]),
    "mmb_act_bwd": (i32, [vp, vp, vp, ll, i32, vp]),
    "mmb_clip_image_transform_max_taps": (i32, []),
    "mmb_clip_image_transform": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, vp]),
    "mmb_attention_bwd_generic": (i32, [vp, ll, ll, vp, ll, ll, vp, ll, ll, vp, ll, ll, vp, ll, ll, vp, vp, ll, vp, vp, vp,
                                        i32, i32, i32, i32, i32, i32, f32, vp]),
}


def declare(lib: C.CDLL) -> None:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
