"""CPU oracle for the Llama decode step around the attention path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module; the
product (atoma_infer_b200/, libflashattention) never does.

Numpy restatement of what the reference computes for one decode step, op by op, with a rounding to
the model dtype wherever the reference materialises a tensor in that dtype:

  * Linear (no bias) ............ models/src/llama.rs:269-271, 311, 364-365, 476 (candle `Linear` ->
                                  cuBLAS: bf16 operands, fp32 accumulation, bf16 result)
  * RoPE ........................ models/src/llama.rs:218-251 (candle_nn::rotary_emb::rope), tables :145-199
  * cache write + paged decode .. oracle/attention_oracle.py (pinned there)
  * SwiGLU ...................... models/src/llama.rs:364: silu(gate) * up
  * RMSNorm, residual adds ...... models/src/llama.rs:398-410, 474
  * greedy sampling ............. backends/vllm/src/model_executor.rs:227-262 + candle_transformers
                                  LogitsProcessor::sample_argmax (max_by(total_cmp): LAST maximum)
  * tensor parallelism .......... models/src/llama_nccl.rs:97-139, multi_gpu.rs:141-179 (column / row
                                  parallel linears, all-reduce(sum) in the model dtype)

PARITY for the candle-side ops is UNPINNED BY THE REFERENCE (candle-core / candle-nn / candle-kernels /
candle-transformers 0.9.2-alpha.1 are registry dependencies pinned in Cargo.lock, absent from /root/reference,
and the reference holds no test that asserts values of RoPE, RMSNorm, SiLU, Linear or the logits).  Substitute
anchor: an independent implementation of the same architecture — Hugging Face transformers' Llama in float64,
of which candle-transformers' llama (what models/src/llama.rs follows) is a port: tests/golden/make_llama_golden.py
commits its KV caches and logits for four sequences and tests/test_llama_oracle.py reproduces the logits with this
module's `exact` path to 1e-6 (non-interleaved RoPE with Llama-3 scaling, RMSNorm, GQA attention over a paged
cache, SwiGLU, residual wiring, final norm, lm_head).  What is restated is the ops' published algorithm —
  rms_norm:  y = T( x * rsqrt(mean(x^2) + eps) * w ), statistics and products in fp32;
  silu:      x / (1 + exp(-x));   rope: see attention_oracle.rope;   Linear: x @ W^T.
The attention and cache pieces reuse the pinned oracle.  `exact=True` variants skip the intermediate
roundings (fp64 throughout) and serve as the tolerance anchor.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import attention_oracle as ao


def _r(x, is_bf16: bool, exact: bool):
    return np.asarray(x, dtype=np.float64) if exact else ao.round_to_io(np.asarray(x, dtype=np.float32), is_bf16).astype(np.float64)


def linear(x: np.ndarray, w: np.ndarray, is_bf16: bool = True, exact: bool = False) -> np.ndarray:
    """x [T, K] @ w[N, K]^T, fp64 accumulation, rounded to the model dtype (llama.rs:3, 322-325)."""
    return _r(np.asarray(x, np.float64) @ np.asarray(w, np.float64).T, is_bf16, exact)


def rms_norm(x: np.ndarray, weight: np.ndarray, eps: float, is_bf16: bool = True, exact: bool = False) -> np.ndarray:
    x = np.asarray(x, np.float64)
    scale = 1.0 / np.sqrt((x * x).mean(axis=-1, keepdims=True) + eps)
    return _r(x * scale * np.asarray(weight, np.float64), is_bf16, exact)


def silu(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float64)
    return x / (1.0 + np.exp(-x))


def silu_mul(gate: np.ndarray, up: np.ndarray, is_bf16: bool = True, exact: bool = False) -> np.ndarray:
    """(candle_nn::ops::silu(gate) * up): the reference rounds silu(gate) to the model dtype before the
    multiply (two element-wise kernels); the product path rounds once (fused) — both lie within one
    rounding of this value."""
    return _r(_r(silu(gate), is_bf16, exact) * np.asarray(up, np.float64), is_bf16, exact)


def add(a: np.ndarray, b: np.ndarray, is_bf16: bool = True, exact: bool = False) -> np.ndarray:
    return _r(np.asarray(a, np.float64) + np.asarray(b, np.float64), is_bf16, exact)


def greedy(logits: np.ndarray):
    """(token ids, their logits): LAST index among equal maxima under f32 total order."""
    x = np.asarray(logits, np.float32)
    key = x.view(np.int32).astype(np.int64)
    key = np.where(key < 0, key ^ 0x7FFFFFFF, key)           # sign-magnitude -> total order
    best = key.max(axis=-1, keepdims=True)
    idx = np.where(key == best, np.arange(x.shape[-1])[None, :], -1).max(axis=-1)
    return idx.astype(np.int64), x[np.arange(x.shape[0]), idx]


def _paged_decode_fp64(q, k_cache, v_cache, block_tables, seq_lens, scale):
    """softmax(scale q K^T) V per sequence and head in float64; K/V rows gathered through the block table
    (utils.h:305-313), kv head = head // (H / Hkv) (flash_fwd_kernel.h:593), empty sequence -> 0."""
    q = np.asarray(q, np.float64)
    T, h, d = q.shape
    page, hk = k_cache.shape[1], k_cache.shape[2]
    out = np.zeros((T, h, d), np.float64)
    for t in range(T):
        n = int(seq_lens[t])
        if n <= 0:
            continue
        pos = np.arange(n)
        blk = np.asarray(block_tables[t])[pos // page]
        k = np.asarray(k_cache, np.float64)[blk, pos % page]      # [n, hk, d]
        v = np.asarray(v_cache, np.float64)[blk, pos % page]
        for head in range(h):
            kvh = head // (h // hk)
            s_ = (k[:, kvh] @ q[t, head]) * scale
            p = np.exp(s_ - s_.max())
            out[t, head] = (p @ v[:, kvh]) / p.sum()
    return out


@dataclass
class LayerWeights:
    wq: np.ndarray
    wk: np.ndarray
    wv: np.ndarray
    wo: np.ndarray
    w_gate: np.ndarray
    w_up: np.ndarray
    w_down: np.ndarray
    rms_attn: np.ndarray
    rms_mlp: np.ndarray


def split_fused(wqkv: np.ndarray, w_gate_up: np.ndarray, num_heads: int, num_kv_heads: int, head_dim: int):
    nq, nkv = num_heads * head_dim, num_kv_heads * head_dim
    inter = w_gate_up.shape[0] // 2
    return wqkv[:nq], wqkv[nq:nq + nkv], wqkv[nq + nkv:], w_gate_up[:inter], w_gate_up[inter:]


def decode_step(hidden: np.ndarray, layers: List[LayerWeights], kv_caches: List[np.ndarray], norm_f: np.ndarray,
                lm_head: Optional[np.ndarray], cos: np.ndarray, sin: np.ndarray, positions: np.ndarray,
                slot_mapping: np.ndarray, block_tables: np.ndarray, seq_lens: np.ndarray, num_heads: int,
                num_kv_heads: int, head_dim: int, eps: float, is_bf16: bool = True, exact: bool = False,
                tp_shards: Optional[List[List[LayerWeights]]] = None):
    """One decode step for T sequences (llama.rs:456-478).  hidden [T, hidden] = wte(tokens); kv_caches[l]
    float32 [2, NB, page, Hkv, D] (values representable in the model dtype), UPDATED IN PLACE with the
    new token's k/v.  Returns (final normed hidden [T, hidden], logits [T, vocab] or None).

    tp_shards: optional [rank][layer] LayerWeights of a tensor-parallel split (column-parallel q/k/v/gate/up
    shards, row-parallel o/down shards, each rank's kv heads = its slice of the cache's head axis): the
    row-parallel outputs are summed over ranks in fp64 and rounded once (the product's all-reduce; NCCL's
    ring would round per hop)."""
    T = hidden.shape[0]
    D, H, HK = head_dim, num_heads, num_kv_heads
    scale = 1.0 / np.sqrt(D)
    x = _r(hidden, is_bf16, exact)
    page = kv_caches[0].shape[2]

    def attention(xn, lw: LayerWeights, kc, vc, h, hk):
        q = linear(xn, lw.wq, is_bf16, exact).reshape(T, h, D)
        k = linear(xn, lw.wk, is_bf16, exact).reshape(T, hk, D)
        v = linear(xn, lw.wv, is_bf16, exact).reshape(T, hk, D)
        q = _r(ao.rope(q, positions, cos, sin), is_bf16, exact)
        k = _r(ao.rope(k, positions, cos, sin), is_bf16, exact)
        for t in range(T):                       # reshape_and_cache_flash (cache_manager.cu:140-170)
            s = int(slot_mapping[t])
            if s < 0:
                continue
            kc[s // page, s % page] = k[t]
            vc[s // page, s % page] = v[t]
        if exact:
            # no model-dtype rounding anywhere, the attention boundary included (ao.attention_paged_decode takes 16-bit
            # buffers like the FFI does): the same gather + softmax in float64 — pinned against Hugging Face's Llama by
            # tests/test_llama_oracle.py::test_decode_step_oracle_reproduces_hf_llama_logits
            return _paged_decode_fp64(q, kc, vc, block_tables, seq_lens, scale).reshape(T, h * D)
        o, _ = ao.attention_paged_decode(q.reshape(T, 1, h, D).astype(np.float32), kc.astype(np.float32),
                                         vc.astype(np.float32), block_tables, seq_lens, scale, is_bf16,
                                         precision="fp64", exact=False)
        return np.asarray(o, np.float64).reshape(T, h * D)

    for l, lw in enumerate(layers):
        kv = kv_caches[l]
        xn = rms_norm(x, lw.rms_attn, eps, is_bf16, exact)
        if tp_shards is None:
            a = attention(xn, lw, kv[0], kv[1], H, HK)
            attn_out = linear(a, lw.wo, is_bf16, exact)
        else:
            n = len(tp_shards)
            parts = []
            for r in range(n):
                sw = tp_shards[r][l]
                hs = slice(r * HK // n, (r + 1) * HK // n)
                kc, vc = kv[0][:, :, hs].copy(), kv[1][:, :, hs].copy()
                a = attention(xn, sw, kc, vc, H // n, HK // n)
                kv[0][:, :, hs], kv[1][:, :, hs] = kc, vc
                parts.append(np.asarray(a, np.float64) @ np.asarray(sw.wo, np.float64).T if exact
                             else linear(a, sw.wo, is_bf16, False))
            attn_out = _r(np.sum(parts, axis=0), is_bf16, exact)
        x = add(attn_out, x, is_bf16, exact)
        xn = rms_norm(x, lw.rms_mlp, eps, is_bf16, exact)
        if tp_shards is None:
            act = silu_mul(linear(xn, lw.w_gate, is_bf16, exact), linear(xn, lw.w_up, is_bf16, exact), is_bf16, exact)
            mlp = linear(act, lw.w_down, is_bf16, exact)
        else:
            parts = []
            for r in range(len(tp_shards)):
                sw = tp_shards[r][l]
                act = silu_mul(linear(xn, sw.w_gate, is_bf16, exact), linear(xn, sw.w_up, is_bf16, exact), is_bf16, exact)
                parts.append(np.asarray(act, np.float64) @ np.asarray(sw.w_down, np.float64).T if exact
                             else linear(act, sw.w_down, is_bf16, False))
            mlp = _r(np.sum(parts, axis=0), is_bf16, exact)
        x = add(mlp, x, is_bf16, exact)
    xf = rms_norm(x, norm_f, eps, is_bf16, exact)
    logits = linear(xf, lm_head, is_bf16, exact) if lm_head is not None else None
    return xf, logits
