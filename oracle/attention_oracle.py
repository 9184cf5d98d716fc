"""CPU oracle for the attention + KV-cache hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product (atoma_infer_b200/, libflashattention) never does.

It restates, in numpy, what the reference's C entry points compute for given ARGUMENTS AND MEMORY —
i.e. it interprets the same pointer/stride/flag tuple the Rust wrappers hand to `run_mha`
(reference csrc/src/ffi.rs:4-66) over flat element buffers:

  * sequence bookkeeping ........ csrc/kernels/block_info.h:14-36
  * paged K/V addressing ........ csrc/kernels/utils.h:296-314
  * GQA head mapping ............ csrc/kernels/flash_fwd_kernel.h:593, 598
  * masks (tail, causal, ALiBi) . csrc/kernels/mask.h:120-209
  * softmax / output ............ csrc/kernels/softmax.h:136-185, flash_fwd_kernel.h:949-954, 1018
  * empty key range -> zeros .... csrc/kernels/flash_fwd_kernel.h:543-582
  * cache ops ................... csrc/kernels/cache_manager.cu:16-37, 140-170;
                                  csrc/src/cache_manager.rs:18-128 (swap)

Pinned by: the reference's golden tables (tests/golden/attention_tables.json, extracted from
csrc/tests/flash_attn_tests.rs:54-70 and models/src/flash_attention.rs:688-704) in
tests/test_oracle_golden.py, and on the GPU box against the reference's own kernels compiled from
/root/reference (oracle/_ref/libatoma_ref.so) in tests/test_parity_reference_gpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

LOG2E = 1.4426950408889634


# ------------------------------------------------------------------------------------------------
# 16-bit float helpers (numpy has no bfloat16)
# ------------------------------------------------------------------------------------------------
def bf16_bits_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even, NaN preserved (matches __float2bfloat16_rn)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> 16) & np.uint32(1)))
    out = (rounded >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def to_f32(bits: np.ndarray, is_bf16: bool) -> np.ndarray:
    return bf16_bits_to_f32(bits) if is_bf16 else bits.view(np.float16).astype(np.float32)


def from_f32(x: np.ndarray, is_bf16: bool) -> np.ndarray:
    """fp32 -> 16-bit storage bits (uint16)."""
    if is_bf16:
        return f32_to_bf16_bits(x)
    return np.asarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


def round_to_io(x: np.ndarray, is_bf16: bool) -> np.ndarray:
    return to_f32(from_f32(x, is_bf16), is_bf16)


# ------------------------------------------------------------------------------------------------
# run_mha
# ------------------------------------------------------------------------------------------------
@dataclass
class MhaArgs:
    """The live subset of the run_mha argument tuple.  Buffers are flat numpy arrays of uint16
    storage bits (q, k, v) / int32 (cu_seqlens, block_table, seqused_k) / float32 (alibi);
    strides are in elements, exactly as passed over the FFI."""
    q: np.ndarray
    k: np.ndarray
    v: np.ndarray
    b: int
    h: int
    h_k: int
    d: int
    seqlen_q: int
    seqlen_k: int
    q_batch_stride: int
    k_batch_stride: int
    v_batch_stride: int
    o_batch_stride: int
    q_row_stride: int
    k_row_stride: int
    v_row_stride: int
    o_row_stride: int
    q_head_stride: int
    k_head_stride: int
    v_head_stride: int
    o_head_stride: int
    softmax_scale: float
    is_bf16: bool
    is_causal: bool
    cu_seqlens_q: Optional[np.ndarray] = None
    cu_seqlens_k: Optional[np.ndarray] = None
    is_seqlens_k_cumulative: bool = True
    seqused_k: Optional[np.ndarray] = None
    block_table: Optional[np.ndarray] = None
    block_table_batch_stride: int = 0
    page_block_size: int = 1
    alibi_slopes: Optional[np.ndarray] = None
    alibi_slopes_batch_stride: int = 0
    o_size: int = 0  # elements of the flat output buffer


def _seq_info(a: MhaArgs, bidb: int):
    """block_info.h:14-36"""
    sum_s_q = -1 if a.cu_seqlens_q is None else int(a.cu_seqlens_q[bidb])
    sum_s_k = -1 if (a.cu_seqlens_k is None or not a.is_seqlens_k_cumulative) else int(a.cu_seqlens_k[bidb])
    len_q = a.seqlen_q if a.cu_seqlens_q is None else int(a.cu_seqlens_q[bidb + 1]) - sum_s_q
    if a.cu_seqlens_k is None:
        len_k_cache = a.seqlen_k
    elif a.is_seqlens_k_cumulative:
        len_k_cache = int(a.cu_seqlens_k[bidb + 1]) - sum_s_k
    else:
        len_k_cache = int(a.cu_seqlens_k[bidb])
    len_k = int(a.seqused_k[bidb]) if a.seqused_k is not None else len_k_cache
    q_off = bidb * a.q_batch_stride if sum_s_q == -1 else sum_s_q * a.q_row_stride
    o_off = bidb * a.o_batch_stride if sum_s_q == -1 else sum_s_q * a.o_row_stride
    k_off = bidb * a.k_batch_stride if sum_s_k == -1 else sum_s_k * a.k_row_stride
    v_off = bidb * a.v_batch_stride if sum_s_k == -1 else sum_s_k * a.v_row_stride
    return len_q, len_k, q_off, o_off, k_off, v_off


def _kv_row_offsets(a: MhaArgs, bidb: int, len_k: int, seq_off: int, batch_stride: int, row_stride: int):
    """Element offset of every key row t < len_k (utils.h:296-314 when paged)."""
    t = np.arange(len_k, dtype=np.int64)
    if a.block_table is not None:
        page = t // a.page_block_size
        phys = a.block_table[bidb * a.block_table_batch_stride + page].astype(np.int64)
        return phys * batch_stride + (t - page * a.page_block_size) * row_stride
    return seq_off + t * row_stride


def run_mha(a: MhaArgs, precision: str = "fp64", return_exact: bool = False):
    """Returns (o_bits uint16 flat [o_size], lse float32 [b, h, seqlen_q]) and, with
    return_exact, also the un-rounded output as float64 flat [o_size] (what the 16-bit result is a
    rounding of — the tests bound |kernel - exact| by the output format's half-ulp plus 1e-3).

    precision="fp64": exact math in float64, one rounding at the output (the specification).
    precision="reference": fp32 scores/softmax, P rounded to the io type before P.V, fp32
    accumulation — the rounding points of the reference kernel (flash_fwd_kernel.h:949-954).
    Output elements the kernel never writes (rows >= len_q) stay 0 here.
    """
    assert precision in ("fp64", "reference")
    ft = np.float64 if precision == "fp64" else np.float32
    o = np.zeros(a.o_size, dtype=np.uint16)
    o_exact = np.zeros(a.o_size, dtype=np.float64)
    lse = np.full((a.b, a.h, a.seqlen_q), np.inf, dtype=np.float32)
    ratio = a.h // a.h_k
    dcol = np.arange(a.d, dtype=np.int64)
    for bidb in range(a.b):
        len_q, len_k, q_off, o_off, k_off, v_off = _seq_info(a, bidb)
        len_k = max(len_k, 0)
        if len_q <= 0:
            continue
        k_rows = _kv_row_offsets(a, bidb, len_k, k_off, a.k_batch_stride, a.k_row_stride)
        v_rows = _kv_row_offsets(a, bidb, len_k, v_off, a.v_batch_stride, a.v_row_stride)
        rows = np.arange(len_q, dtype=np.int64)
        cols = np.arange(len_k, dtype=np.int64)
        for head in range(a.h):
            kvh = head // ratio  # flash_fwd_kernel.h:593
            qi = q_off + rows[:, None] * a.q_row_stride + head * a.q_head_stride + dcol[None, :]
            oi = o_off + rows[:, None] * a.o_row_stride + head * a.o_head_stride + dcol[None, :]
            qm = to_f32(a.q[qi], a.is_bf16).astype(ft)
            if len_k == 0:
                o[oi] = 0
                continue
            km = to_f32(a.k[k_rows[:, None] + kvh * a.k_head_stride + dcol[None, :]], a.is_bf16).astype(ft)
            vm = to_f32(a.v[v_rows[:, None] + kvh * a.v_head_stride + dcol[None, :]], a.is_bf16).astype(ft)
            s = (qm @ km.T) * ft(a.softmax_scale)
            shift = len_k - len_q  # mask.h:171-172: bottom-right aligned
            if a.alibi_slopes is not None:
                slope = ft(a.alibi_slopes[bidb * a.alibi_slopes_batch_stride + head])
                if a.is_causal:
                    s = s + slope * cols[None, :].astype(ft)  # mask.h:179-182
                else:
                    s = s - slope * np.abs(rows[:, None] + shift - cols[None, :]).astype(ft)  # :183-186
            if a.is_causal:
                s = np.where(cols[None, :] > rows[:, None] + shift, -np.inf, s)
            m = s.max(axis=1, keepdims=True)
            m_safe = np.where(np.isneginf(m), 0.0, m)
            p = np.exp(s - m_safe)
            l = p.sum(axis=1, keepdims=True)
            if precision == "reference":
                p = round_to_io(p.astype(np.float32), a.is_bf16).astype(ft)
            acc = p @ vm
            inv = np.where((l == 0) | np.isnan(l), 1.0, 1.0 / np.where(l == 0, 1.0, l))
            o[oi] = from_f32((acc * inv).astype(np.float32), a.is_bf16)
            o_exact[oi] = acc * inv
            with np.errstate(divide="ignore"):
                row_lse = np.where((l == 0) | np.isnan(l), np.inf, m_safe + np.log(np.where(l == 0, 1.0, l)))
            lse[bidb, head, :len_q] = row_lse[:, 0].astype(np.float32)
    if return_exact:
        return o, lse, o_exact
    return o, lse


# ------------------------------------------------------------------------------------------------
# Cache ops (pure data movement, exact)
# ------------------------------------------------------------------------------------------------
def reshape_and_cache_flash(key: np.ndarray, value: np.ndarray, key_cache: np.ndarray,
                            value_cache: np.ndarray, slot_mapping: np.ndarray, block_stride: int,
                            num_tokens: int, num_heads: int, head_size: int, block_size: int,
                            key_stride: int, value_stride: int) -> None:
    """cache_manager.cu:140-170 over flat uint16 buffers (in place)."""
    n = num_heads * head_size
    i = np.arange(n, dtype=np.int64)
    for t in range(num_tokens):
        slot = int(slot_mapping[t])
        if slot < 0:  # padded token
            continue
        blk, off = slot // block_size, slot % block_size
        tgt = blk * block_stride + off * n + i  # == head_idx*head_size + head_offset
        key_cache[tgt] = key[t * key_stride + i]
        value_cache[tgt] = value[t * value_stride + i]


def copy_blocks(key_caches, value_caches, block_mapping: np.ndarray, numel_per_block: int) -> None:
    """cache_manager.cu:16-37: per layer, per pair, K and V (in place; flat uint16 per layer)."""
    for kc, vc in zip(key_caches, value_caches):
        for src, dst in np.asarray(block_mapping).reshape(-1, 2):
            s, d = int(src) * numel_per_block, int(dst) * numel_per_block
            kc[d:d + numel_per_block] = kc[s:s + numel_per_block].copy()
            vc[d:d + numel_per_block] = vc[s:s + numel_per_block].copy()


def swap_blocks(src: np.ndarray, dst: np.ndarray, block_mapping: dict, block_numel: int) -> None:
    """cache_manager.rs:18-128: dst[block d] = src[block s] (flat buffers, in place)."""
    for s, d in block_mapping.items():
        dst[d * block_numel:(d + 1) * block_numel] = src[s * block_numel:(s + 1) * block_numel]


# ------------------------------------------------------------------------------------------------
# Rotary embedding (adjacent op of the attention path)
# ------------------------------------------------------------------------------------------------
def rope_tables(head_dim: int, max_pos: int, theta: float = 500000.0, is_bf16: bool = True):
    """cos/sin [max_pos, head_dim/2] as the reference builds them (models/src/llama.rs:145-199,
    default rope type): inv_freq_i = theta^(-2i/d) in f32, idx_theta = pos * inv_freq, tables cast
    to the model dtype.  Returned as float32 values representable in that dtype."""
    inv = (1.0 / theta ** (np.arange(0, head_dim, 2, dtype=np.float32) / np.float32(head_dim))).astype(np.float32)
    ang = np.arange(max_pos, dtype=np.float32)[:, None] * inv[None, :]
    return round_to_io(np.cos(ang).astype(np.float32), is_bf16), round_to_io(np.sin(ang).astype(np.float32), is_bf16)


def rope(x: np.ndarray, positions: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """Non-interleaved rotary embedding of x [T, heads, d] (float, un-rounded result in float64):
    models/src/llama.rs:218-251 selects cos/sin rows by position and calls
    candle_nn::rotary_emb::rope (candle-nn 0.9.2-alpha.1, not vendored in /root/reference), whose
    published contract is out[:d/2] = x1*cos - x2*sin, out[d/2:] = x1*sin + x2*cos with
    x1 = x[..., :d/2], x2 = x[..., d/2:].  Parity for this op is unpinned in the reference (no test
    asserts rotated values)."""
    x = np.asarray(x, dtype=np.float64)
    d = x.shape[-1]
    c = np.asarray(cos, dtype=np.float64)[np.asarray(positions)][:, None, :]
    s_ = np.asarray(sin, dtype=np.float64)[np.asarray(positions)][:, None, :]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return np.concatenate([x1 * c - x2 * s_, x1 * s_ + x2 * c], axis=-1)


# ------------------------------------------------------------------------------------------------
# Shape-level conveniences used by the tests
# ------------------------------------------------------------------------------------------------
def attention_dense(q: np.ndarray, k: np.ndarray, v: np.ndarray, scale: float, causal: bool,
                    is_bf16: bool, precision: str = "fp64", alibi=None, exact: bool = False):
    """q [b,sq,h,d], k/v [b,sk,hk,d] float32 (already representable in the io type)."""
    b, sq, h, d = q.shape
    _, sk, hk, _ = k.shape
    a = MhaArgs(q=from_f32(q, is_bf16).ravel(), k=from_f32(k, is_bf16).ravel(), v=from_f32(v, is_bf16).ravel(),
                b=b, h=h, h_k=hk, d=d, seqlen_q=sq, seqlen_k=sk,
                q_batch_stride=sq * h * d, k_batch_stride=sk * hk * d, v_batch_stride=sk * hk * d,
                o_batch_stride=sq * h * d, q_row_stride=h * d, k_row_stride=hk * d, v_row_stride=hk * d,
                o_row_stride=h * d, q_head_stride=d, k_head_stride=d, v_head_stride=d, o_head_stride=d,
                softmax_scale=scale, is_bf16=is_bf16, is_causal=causal, alibi_slopes=alibi,
                o_size=b * sq * h * d)
    o, lse, ox = run_mha(a, precision, return_exact=True)
    if exact:
        return ox.reshape(b, sq, h, d), lse
    return to_f32(o, is_bf16).reshape(b, sq, h, d), lse


def attention_paged_decode(q: np.ndarray, k_cache: np.ndarray, v_cache: np.ndarray, block_table: np.ndarray,
                           seqlens_k: np.ndarray, scale: float, is_bf16: bool, precision: str = "fp64",
                           exact: bool = False):
    """q [b,1,h,d]; caches [nb,page,hk,d]; block_table [b,mb]; seqlens_k [b] (per-sequence)."""
    b, sq, h, d = q.shape
    nb, page, hk, _ = k_cache.shape
    mb = block_table.shape[1]
    a = MhaArgs(q=from_f32(q, is_bf16).ravel(), k=from_f32(k_cache, is_bf16).ravel(),
                v=from_f32(v_cache, is_bf16).ravel(), b=b, h=h, h_k=hk, d=d, seqlen_q=sq, seqlen_k=mb * page,
                q_batch_stride=sq * h * d, k_batch_stride=page * hk * d, v_batch_stride=page * hk * d,
                o_batch_stride=sq * h * d, q_row_stride=h * d, k_row_stride=hk * d, v_row_stride=hk * d,
                o_row_stride=h * d, q_head_stride=d, k_head_stride=d, v_head_stride=d, o_head_stride=d,
                softmax_scale=scale, is_bf16=is_bf16, is_causal=False,
                cu_seqlens_k=seqlens_k.astype(np.int32), is_seqlens_k_cumulative=False,
                block_table=block_table.astype(np.int32).ravel(), block_table_batch_stride=mb,
                page_block_size=page, o_size=b * sq * h * d)
    o, lse, ox = run_mha(a, precision, return_exact=True)
    if exact:
        return ox.reshape(b, sq, h, d), lse
    return to_f32(o, is_bf16).reshape(b, sq, h, d), lse


# ---------------------------------------------------------------------------------------------------
# Caller-side split heuristic (csrc/src/lib.rs:2122-2199), restated with f32 arithmetic where the
# Rust uses f32 (the 0.8 and 0.85 thresholds are compared in single precision).
# ---------------------------------------------------------------------------------------------------
def num_splits_heuristic(batch_nheads_mblocks: int, num_sms: int, num_n_blocks: int, max_splits: int) -> int:
    """lib.rs:2122-2166"""
    f = np.float32
    if f(batch_nheads_mblocks) >= f(0.8) * f(num_sms):
        return 1
    max_splits = min(max_splits, num_sms, num_n_blocks)
    ceil_div = lambda a, b: (a + b - 1) // b
    eligible = lambda s: s == 1 or ceil_div(num_n_blocks, s) != ceil_div(num_n_blocks, s - 1)
    eff, best = [], f(0.0)
    for s in range(1, max_splits + 1):
        if not eligible(s):
            eff.append(f(0.0))
            continue
        n_waves = f(batch_nheads_mblocks * s) / f(num_sms)
        e = n_waves / np.ceil(n_waves)
        best = max(best, e)
        eff.append(e)
    for s in range(1, max_splits + 1):
        if eligible(s) and eff[s - 1] >= f(0.85) * best:
            return s
    return 1


def compute_num_splits(batch_size: int, num_heads: int, head_size: int, max_seqlen_k: int, max_seqlen_q: int,
                       num_sms: int) -> int:
    """lib.rs:2168-2199"""
    block_n = 256 if head_size <= 64 else 128 if head_size <= 128 else 64
    num_n_blocks = (max_seqlen_k + block_n - 1) // block_n
    num_m_blocks = (max_seqlen_q + 63) // 64
    return num_splits_heuristic(batch_size * num_heads * num_m_blocks, num_sms * 2, num_n_blocks, 128)
