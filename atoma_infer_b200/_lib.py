"""ctypes binding of libflashattention.so (include/atoma_flash.h + include/atoma_ops.h).

The library is the product; this module only loads it and declares signatures.  There is no
fallback: if the shared library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("ATOMA_FLASH_LIB", _PKG / "lib" / "libflashattention.so"))

MAX_RANK = 5
DEVICE_CPU = -1
F16, BF16, F32, U32, I64, U8, F64 = range(7)


class Tensor(C.Structure):
    """struct atoma_tensor"""

    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("shape", C.c_int64 * MAX_RANK),
        ("stride", C.c_int64 * MAX_RANK),
    ]


class MhaArgs(C.Structure):
    """struct atoma_mha_args — the run_mha argument tuple (reference csrc/src/ffi.rs:4-66)."""

    _fields_ = [
        ("q_ptr", C.c_void_p), ("k_ptr", C.c_void_p), ("v_ptr", C.c_void_p), ("o_ptr", C.c_void_p),
        ("softmax_lse_ptr", C.c_void_p), ("alibi_slopes_ptr", C.c_void_p),
        ("cu_seqlens_q_ptr", C.c_void_p), ("cu_seqlens_k_ptr", C.c_void_p),
        ("is_seqlens_k_cumulative", C.c_bool),
        ("q_batch_stride", C.c_uint32), ("k_batch_stride", C.c_uint32),
        ("v_batch_stride", C.c_uint32), ("o_batch_stride", C.c_uint32),
        ("alibi_slopes_batch_stride", C.c_uint32),
        ("q_row_stride", C.c_uint32), ("k_row_stride", C.c_uint32),
        ("v_row_stride", C.c_uint32), ("o_row_stride", C.c_uint32),
        ("q_head_stride", C.c_uint32), ("k_head_stride", C.c_uint32),
        ("v_head_stride", C.c_uint32), ("o_head_stride", C.c_uint32),
        ("num_splits", C.c_uint32),
        ("b", C.c_uint32), ("h", C.c_uint32), ("h_k", C.c_uint32), ("d", C.c_uint32),
        ("d_rounded", C.c_uint32),
        ("softmax_scale", C.c_float), ("scale_softmax_log2", C.c_float),
        ("block_table", C.c_void_p), ("block_table_batch_stride", C.c_uint32),
        ("page_block_size", C.c_int),
        ("seqused_k", C.c_void_p),
        ("seqlen_q", C.c_uint32), ("seqlen_k", C.c_uint32),
        ("seqlen_q_rounded", C.c_uint32), ("seqlen_k_rounded", C.c_uint32),
        ("is_bf16", C.c_int), ("is_causal", C.c_int),
        ("window_size_left", C.c_int), ("window_size_right", C.c_int),
        ("softcap", C.c_float),
        ("unpadded_lse", C.c_bool), ("force_split_kernel", C.c_bool),
        ("softmax_lseaccum_ptr", C.c_void_p), ("oaccum_ptr", C.c_void_p),
    ]


class MhaPlan(C.Structure):
    """struct atoma_mha_plan"""

    _fields_ = [
        ("args", MhaArgs),
        ("out_rank", C.c_int32),
        ("out_shape", C.c_int64 * 4),
        ("lse_elems", C.c_int64),
        ("lseaccum_elems", C.c_int64),
        ("oaccum_elems", C.c_int64),
        ("seqlenq_ngroups_swapped", C.c_int32),
    ]


class Sequence(C.Structure):
    """struct atoma_sequence"""

    _fields_ = [("token_ids", C.POINTER(C.c_uint32)), ("length", C.c_int64), ("num_computed_tokens", C.c_int64),
                ("token_chunk_size", C.c_int64), ("block_table", C.POINTER(C.c_uint32)), ("num_blocks", C.c_int64),
                ("is_prompt", C.c_int32)]


class ModelInputLayout(C.Structure):
    """struct atoma_model_input_layout"""

    _fields_ = [(n, C.c_int64) for n in (
        "total_bytes", "input_tokens_off", "input_positions_off", "slot_mapping_off", "sequence_lengths_off",
        "context_lengths_off", "sequence_start_loc_off", "query_start_loc_off", "block_tables_off",
        "num_tokens", "num_slots", "num_sequences", "max_block_table_len",
        "num_prefills", "num_prefill_tokens", "num_decode_tokens",
        "max_query_len", "max_prefill_seq_len", "max_decode_seq_len")]


class AttentionConfig(C.Structure):
    _fields_ = [
        ("num_heads", C.c_int64), ("num_kv_heads", C.c_int64), ("head_dim", C.c_int64),
        ("softmax_scale", C.c_float), ("sliding_window", C.c_int64),
    ]


class AttentionMetadata(C.Structure):
    _fields_ = [
        ("slot_mapping", C.POINTER(Tensor)),
        ("num_prefill_tokens", C.c_int64), ("num_decoding_tokens", C.c_int64),
        ("has_prefill", C.c_int32),
        ("prefill_block_tables", C.POINTER(Tensor)),
        ("max_prefill_sequence_length", C.c_int64),
        ("query_start_locations", C.POINTER(Tensor)),
        ("sequence_start_locations", C.POINTER(Tensor)),
        ("max_prefill_sequence_length_k", C.c_int64),
        ("has_decode", C.c_int32),
        ("decode_block_tables", C.POINTER(Tensor)),
        ("decode_sequence_lengths", C.POINTER(Tensor)),
    ]


class LlamaLayer(C.Structure):
    """struct atoma_llama_layer (include/atoma_llama.h)"""

    _fields_ = [(n, C.c_void_p) for n in ("wqkv", "wo", "w_gate_up", "w_down", "rms_attn", "rms_mlp",
                                          "key_cache", "value_cache")]


class LlamaModel(C.Structure):
    """struct atoma_llama_model"""

    _fields_ = ([(n, C.c_int64) for n in ("hidden", "inter", "num_layers", "num_heads", "num_kv_heads", "head_dim",
                                           "vocab", "block_size", "block_stride", "rot_stride")] +
                [("rms_eps", C.c_float), ("softmax_scale", C.c_float), ("dtype", C.c_uint32),
                 ("embed", C.c_void_p), ("norm_f", C.c_void_p), ("lm_head", C.c_void_p),
                 ("cos", C.c_void_p), ("sin", C.c_void_p), ("layers", C.POINTER(LlamaLayer)),
                 ("tp_rank", C.c_int32), ("tp_world", C.c_int32), ("p2p_regions", C.POINTER(C.c_void_p)),
                 ("p2p_slot_bytes", C.c_int64), ("lm_head_rows", C.c_int64), ("lm_head_row0", C.c_int64)])


class DecodeBatch(C.Structure):
    """struct atoma_decode_batch"""

    _fields_ = [("num_tokens", C.c_int64), ("tokens", C.c_void_p), ("hidden_in", C.c_void_p),
                ("positions", C.c_void_p), ("slot_mapping", C.c_void_p), ("block_tables", C.c_void_p),
                ("sequence_lengths", C.c_void_p), ("max_blocks", C.c_int64), ("scratch", C.c_void_p),
                ("scratch_bytes", C.c_int64), ("hidden_out", C.c_void_p), ("logits", C.c_void_p),
                ("logits_f32", C.c_void_p), ("next_tokens", C.c_void_p), ("next_logits", C.c_void_p)]


# Exact C parameter lists, in header order.
_RUN_MHA_ARGS = [
    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,  # q k v o lse alibi
    C.c_void_p, C.c_void_p,  # cu_seqlens_q, cu_seqlens_k
    C.c_bool,
    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,  # batch strides + alibi
    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,  # row strides
    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,  # head strides
    C.c_uint32,  # num_splits
    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,  # b h h_k d d_rounded
    C.c_float, C.c_float,
    C.c_void_p, C.c_uint32, C.c_int,  # block_table, stride, page
    C.c_void_p,  # seqused_k
    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,  # seqlens
    C.c_int, C.c_int, C.c_int, C.c_int,  # is_bf16 is_causal wl wr
    C.c_float, C.c_bool, C.c_bool,
    C.c_void_p, C.c_void_p,
]

# symbol -> (restype, argtypes); every symbol include/*.h declares is listed here and checked by
# tests/test_abi.py against the headers.
SIGNATURES = {
    # atoma_flash.h, part 1 (reference ABI)
    "run_mha": (None, _RUN_MHA_ARGS),
    "copy_blocks_f16": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "copy_blocks_bf16": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "reshape_and_cache_flash": (None, [C.c_void_p] * 5 + [C.c_int64] * 7 + [C.c_uint32, C.c_void_p]),
    # part 2 (extensions)
    "run_mha_stream": (None, _RUN_MHA_ARGS + [C.c_void_p]),
    "atoma_rope_reshape_and_cache": (None, [C.c_void_p] * 9 + [C.c_int64] * 10 + [C.c_uint32, C.c_void_p]),
    "swap_blocks_batched": (None, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "atoma_gemm_bf16_tn": (None, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 6 + [C.c_void_p]),
    "atoma_p2p_region_bytes": (C.c_int64, [C.c_int64]),
    "atoma_p2p_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "atoma_p2p_free": (C.c_int, [C.c_void_p]),
    "atoma_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "atoma_p2p_import": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "atoma_p2p_close": (C.c_int, [C.c_void_p]),
    "atoma_p2p_slot": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int]),
    "atoma_p2p_error": (C.c_int, [C.c_void_p, C.c_void_p]),
    "atoma_allreduce_sum_p2p": (None, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                       C.c_int64, C.c_uint32, C.c_void_p]),
    "atoma_debug_allreduce_algo": (None, [C.c_int]),
    "atoma_debug_decode_trace": (None, [C.c_void_p]),
    "atoma_debug_decode_merge": (None, [C.c_int]),
    "atoma_debug_step_trace": (None, [C.c_void_p, C.c_int]),
    "atoma_debug_step_trace_names": (C.c_int, [C.c_char_p, C.c_int]),
    "atoma_debug_gemm_trace": (None, [C.c_void_p]),
    "atoma_debug_gemm_stages": (None, [C.c_int]),
    "atoma_debug_gemm_algo": (None, [C.c_int]),
    "atoma_set_launch_mode": (None, [C.c_int]),
    "atoma_get_launch_mode": (C.c_int, []),
    "atoma_last_status": (C.c_int, []),
    "atoma_last_error": (C.c_char_p, []),
    "atoma_last_mha_path": (C.c_char_p, []),
    "atoma_last_decode_mode": (C.c_int, []),
    "atoma_launch_count": (C.c_int64, []),
    "atoma_reset_launch_count": (None, []),
    "atoma_version": (C.c_char_p, []),
    "atoma_source_hash": (C.c_char_p, []),
    "atoma_mha_call_flags": (None, [C.c_int]),
    # atoma_llama.h
    "atoma_gemm_tn": (None, [C.c_void_p] * 3 + [C.c_int64] * 6 + [C.c_uint32, C.c_void_p]),
    "atoma_qkv_rope_cache": (None, [C.c_void_p] * 3 + [C.c_int64] * 8 + [C.c_void_p] * 3 + [C.c_int64] +
                             [C.c_void_p] * 3 + [C.c_int64] * 2 + [C.c_uint32, C.c_void_p]),
    "atoma_gate_up_silu": (None, [C.c_void_p] * 3 + [C.c_int64] * 6 + [C.c_uint32, C.c_void_p]),
    "atoma_add_rmsnorm": (None, [C.c_void_p] * 5 + [C.c_int64] * 6 + [C.c_float, C.c_uint32, C.c_void_p]),
    "atoma_embedding": (None, [C.c_void_p] * 3 + [C.c_int64] * 5 + [C.c_uint32, C.c_void_p]),
    "atoma_sample_greedy": (None, [C.c_void_p] + [C.c_int64] * 3 + [C.c_uint32] + [C.c_void_p] * 3 +
                            [C.c_int64, C.c_void_p]),
    "atoma_gemm_tn_push": (None, [C.c_void_p] * 2 + [C.c_int64] * 5 + [C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                                                       C.c_int64, C.c_uint32, C.c_void_p]),
    "atoma_allreduce_recv_add_rmsnorm": (None, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 4 +
                                         [C.c_int64] * 5 + [C.c_float, C.c_uint32, C.c_void_p]),
    "atoma_llama_decode_scratch_bytes": (C.c_int64, [C.POINTER(LlamaModel), C.c_int64]),
    "atoma_llama_decode_step": (None, [C.POINTER(LlamaModel), C.POINTER(DecodeBatch), C.c_void_p]),
    "atoma_debug_skinny_trace": (None, [C.c_void_p]),
    "atoma_debug_skinny_tuning": (None, [C.c_int] * 3),
    "atoma_debug_skinny_mode": (None, [C.c_int] * 2),
    "atoma_debug_skinny_budget": (None, [C.c_int]),
    "atoma_debug_step_prefetch": (None, [C.c_int] * 2),
    "atoma_debug_step_fusion": (None, [C.c_int]),
    "atoma_llama_abi_sizeof": (C.c_int64, [C.c_int]),
    "atoma_llama_step_status": (C.c_int, [C.c_void_p]),
    "atoma_mlp_gate_up_down": (None, [C.c_void_p] * 5 + [C.c_int64] * 8 + [C.c_uint32, C.c_void_p]),
    "atoma_sample_greedy_tp": (None, [C.c_void_p] + [C.c_int64] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p,
                                                                        C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "atoma_gemm_tn_add_rmsnorm": (None, [C.c_void_p] * 3 + [C.c_int64] * 6 + [C.c_void_p] * 4 + [C.c_int64] * 3 +
                                  [C.c_float, C.c_uint32, C.c_void_p]),
    # atoma_ops.h
    "atoma_flash_attn_plan": (C.c_char_p, [C.POINTER(Tensor)] * 4 + [C.c_float, C.c_int64, C.c_int64, C.c_float,
                                                                    C.c_int, C.POINTER(MhaPlan)]),
    "atoma_flash_attn_varlen_plan": (C.c_char_p, [C.POINTER(Tensor)] * 8 + [C.c_int64, C.c_int64, C.c_float,
                                                                           C.c_int64, C.c_int64, C.c_float, C.c_int,
                                                                           C.POINTER(MhaPlan)]),
    "atoma_flash_attn_kv_cache_plan": (C.c_char_p, [C.POINTER(Tensor)] * 6 + [C.c_float, C.c_int64, C.c_int64,
                                                                             C.c_int, C.POINTER(MhaPlan)]),
    "atoma_mha_execute": (C.c_char_p, [C.POINTER(MhaPlan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p]),
    "atoma_compute_num_splits": (C.c_uint32, [C.c_int64] * 5 + [C.c_int]),
    "atoma_reshape_and_cache_flash": (C.c_char_p, [C.POINTER(Tensor)] * 5 + [C.c_void_p]),
    "atoma_copy_blocks": (C.c_char_p, [C.POINTER(Tensor), C.c_int64, C.POINTER(Tensor), C.c_int64,
                                       C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "atoma_swap_blocks": (C.c_char_p, [C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.c_int64, C.c_void_p]),
    "atoma_prepare_input": (C.c_char_p, [C.POINTER(Sequence), C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                         C.POINTER(ModelInputLayout), C.c_void_p, C.c_int64]),
    "atoma_flash_attention_forward": (C.c_char_p, [C.POINTER(AttentionConfig)] + [C.POINTER(Tensor)] * 5 +
                                      [C.POINTER(AttentionMetadata), C.POINTER(Tensor), C.c_void_p]),
}


def source_hash() -> str:
    """What csrc/Makefile compiles into atoma_source_hash(): SHA-256 (16 hex digits) of the library's sources,
    csrc/ then include/, each in sorted order.  Equal to the loaded library's hash <=> the binary was built
    from this tree."""
    import hashlib
    src = _PKG / "csrc"
    files = sorted(str(p) for pat in ("*.cu", "*.cuh", "*.h", "*.cpp") for p in src.glob(pat))
    files += sorted(str(p) for p in (_PKG.parent / "include").glob("*.h"))
    h = hashlib.sha256()
    for f in files:
        h.update(Path(f).read_bytes())
    return h.hexdigest()[:16]


def load(path: os.PathLike | str | None = None) -> C.CDLL:
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise ImportError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C atoma_infer_b200/csrc`). There is no CPU or PyTorch fallback for this library."
        )
    lib = C.CDLL(str(p), mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB: C.CDLL | None = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB
