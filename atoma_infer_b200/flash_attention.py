"""Operator layer: mirror of `models::FlashAttention` (reference models/src/flash_attention.rs).

Metadata structs keep the reference's field names (:11-146); `FlashAttention.forward` (:322-469)
routes one flattened batch — cache write, prefill attention on tokens [0, P), decode attention on
tokens [P, P+Dn) — through the C++ host layer (`atoma_flash_attention_forward`), which owns the
routing logic; this file only marshals tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from ._lib import AttentionConfig, AttentionMetadata
from .ops import AtomaError, desc


@dataclass
class FlashAttentionDecodingMetadata:
    """flash_attention.rs:11-29"""
    block_tables: Optional[torch.Tensor]
    max_decoding_sequence_length: int
    sequence_lengths: Optional[torch.Tensor]


@dataclass
class FlashAttentionPrefillMetadata:
    """flash_attention.rs:34-62"""
    block_tables: Optional[torch.Tensor]
    max_query_length: Optional[int]
    max_prefill_sequence_length: int
    query_start_locations: Optional[torch.Tensor]
    sequence_start_locations: Optional[torch.Tensor]
    sequence_lengths: Optional[torch.Tensor]


@dataclass
class FlashAttentionMetadata:
    """flash_attention.rs:67-146"""
    context_lengths: Optional[torch.Tensor]
    slot_mapping: torch.Tensor
    decoding_metadata: Optional[FlashAttentionDecodingMetadata]
    prefill_metadata: Optional[FlashAttentionPrefillMetadata]
    num_prefill_tokens: int
    num_decoding_tokens: int

    @classmethod
    def new(cls, context_lengths, slot_mapping, query_start_locations, num_prefill_tokens,
            num_decoding_tokens, max_query_length, max_decoding_sequence_length,
            max_prefill_sequence_length, num_prefill_sequences, sequence_start_locations,
            sequence_lengths, block_table, prefix_caching: bool) -> "FlashAttentionMetadata":
        """flash_attention.rs:87-146"""
        prefill_bt = block_table[:num_prefill_tokens] if prefix_caching and num_prefill_tokens > 0 else None
        decode_bt = block_table[num_prefill_tokens:] if prefix_caching and num_decoding_tokens > 0 else block_table
        prefill = None
        if num_prefill_tokens > 0:
            prefill = FlashAttentionPrefillMetadata(
                block_tables=prefill_bt, max_query_length=max_query_length,
                max_prefill_sequence_length=max_prefill_sequence_length,
                query_start_locations=query_start_locations[:num_prefill_sequences],
                sequence_start_locations=sequence_start_locations[:num_prefill_sequences + 1],
                sequence_lengths=sequence_lengths[:num_prefill_sequences])
        decode = None
        if num_decoding_tokens > 0:
            decode = FlashAttentionDecodingMetadata(
                block_tables=decode_bt, max_decoding_sequence_length=max_decoding_sequence_length,
                sequence_lengths=sequence_lengths[num_prefill_sequences:])
        return cls(context_lengths, slot_mapping, decode, prefill, num_prefill_tokens, num_decoding_tokens)


class FlashAttention:
    """flash_attention.rs:175-469"""

    def __init__(self, num_heads: int, num_kv_heads: int, head_dim: int, softmax_scale: float,
                 alibi_slopes: Optional[torch.Tensor] = None, sliding_window: Optional[int] = None,
                 kv_cache_dtype: torch.dtype = torch.bfloat16, device=None, check_head_dim: bool = True):
        if num_heads % num_kv_heads != 0:
            raise AtomaError(f"number of heads {num_heads} must divide number of kv heads {num_kv_heads}")
        if check_head_dim and head_dim not in self.supported_head_sizes():
            raise AtomaError(f"head_dim {head_dim} is not supported")
        self.num_heads, self.num_kv_heads, self.head_dim = num_heads, num_kv_heads, head_dim
        self.num_queries_per_kv = num_heads // num_kv_heads
        self.softmax_scale = softmax_scale
        self.alibi_slopes = alibi_slopes
        self.sliding_window = sliding_window
        self.kv_cache_dtype = kv_cache_dtype
        self.device = device

    @staticmethod
    def supported_head_sizes() -> List[int]:
        return [64, 80, 96, 112, 128, 192, 256]  # flash_attention.rs:234-236

    @staticmethod
    def get_kv_cache_shape(num_blocks, block_size, num_kv_heads, head_dim) -> List[int]:
        return [2, num_blocks, block_size, num_kv_heads, head_dim]  # flash_attention.rs:239-246

    def split_kv_cache(self, kv_cache: torch.Tensor):
        """flash_attention.rs:249-278"""
        if kv_cache.dim() != 5:
            raise AtomaError(f"KV cache must have rank 5 (got {kv_cache.dim()})")
        if kv_cache.shape[0] != 2:
            raise AtomaError(f"KV cache must have cache_size 2 (got {kv_cache.shape[0]})")
        if kv_cache.shape[3] != self.num_kv_heads:
            raise AtomaError(f"KV cache must have num_heads {self.num_kv_heads} (got {kv_cache.shape[3]})")
        if kv_cache.shape[4] != self.head_dim:
            raise AtomaError(f"KV cache must have head_dim {self.head_dim} (got {kv_cache.shape[4]})")
        return kv_cache[0], kv_cache[1]

    def swap_blocks(self, src: torch.Tensor, dst: torch.Tensor, block_mapping: Dict[int, int]) -> None:
        """flash_attention.rs:281-291"""
        sk, sv = self.split_kv_cache(src)
        dk, dv = self.split_kv_cache(dst)
        ops.swap_blocks(sk, dk, block_mapping)
        ops.swap_blocks(sv, dv, block_mapping)

    @staticmethod
    def copy_blocks(kv_caches: List[torch.Tensor], block_mapping: torch.Tensor) -> None:
        """flash_attention.rs:294-306"""
        ops.copy_blocks([c[0] for c in kv_caches], [c[1] for c in kv_caches], block_mapping)

    def forward(self, q, k, v, kv_cache, attention_metadata: FlashAttentionMetadata,
                max_prefill_sequence_length_k: Optional[int] = None) -> torch.Tensor:
        """q [T, h, d]; k, v [T, hk, d]; kv_cache [2, NB, bs, hk, d] -> [T, h*d]   (:322-469)"""
        md = attention_metadata
        cfg = AttentionConfig(self.num_heads, self.num_kv_heads, self.head_dim, self.softmax_scale,
                              -1 if self.sliding_window is None else self.sliding_window)
        keep = []

        def ref(t):
            if t is None:
                return None
            d = desc(t)
            keep.append(d)
            return C.pointer(d)

        m = AttentionMetadata()
        m.slot_mapping = ref(md.slot_mapping)
        m.num_prefill_tokens = md.num_prefill_tokens
        m.num_decoding_tokens = md.num_decoding_tokens
        pm, dm = md.prefill_metadata, md.decoding_metadata
        m.has_prefill = 1 if pm is not None else 0
        if pm is not None:
            if pm.block_tables is not None and pm.block_tables.numel() > 0:
                m.prefill_block_tables = ref(pm.block_tables)
                if max_prefill_sequence_length_k is None:
                    # the reference does `sequence_lengths.max(0)` here (a device->host sync)
                    max_prefill_sequence_length_k = int(pm.sequence_lengths.max().item())
                m.max_prefill_sequence_length_k = max_prefill_sequence_length_k
            m.max_prefill_sequence_length = pm.max_prefill_sequence_length
            m.query_start_locations = ref(pm.query_start_locations)
            m.sequence_start_locations = ref(pm.sequence_start_locations)
        m.has_decode = 1 if dm is not None else 0
        if dm is not None:
            m.decode_block_tables = ref(dm.block_tables)
            m.decode_sequence_lengths = ref(dm.sequence_lengths)
        out = torch.empty((q.shape[0], self.num_heads * self.head_dim), dtype=q.dtype, device=q.device)
        qd, kd, vd, cd, od = desc(q), desc(k), desc(v), desc(kv_cache), desc(out)
        ad = desc(self.alibi_slopes)
        msg = _lib.lib().atoma_flash_attention_forward(
            C.byref(cfg), C.byref(qd), C.byref(kd), C.byref(vd), C.byref(cd),
            C.byref(ad) if ad is not None else None, C.byref(m), C.byref(od), ops._stream())
        if msg:
            raise AtomaError(msg.decode())
        return out
