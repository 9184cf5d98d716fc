"""Python face of the host operator layer — same names, argument meaning and error behaviour as the
reference's Rust crate `csrc` (csrc/src/lib.rs pub fns :392-2105, csrc/src/cache_manager.rs).

Everything here is a thin shim: tensors are described to the C ABI (include/atoma_ops.h), the C++
host layer validates / derives the `run_mha` arguments, and the CUDA kernels do the work.  torch is
used for device memory and streams only.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import MhaPlan, Tensor


class AtomaError(RuntimeError):
    """The analogue of `candle_core::Error` raised by `bail!` in the reference wrappers."""


_DTYPES = {
    torch.float16: _lib.F16,
    torch.bfloat16: _lib.BF16,
    torch.float32: _lib.F32,
    torch.int32: _lib.U32,  # candle has no i32; u32 tensors travel as int32 bit patterns
    torch.int64: _lib.I64,
    torch.uint8: _lib.U8,
    torch.float64: _lib.F64,
}
if hasattr(torch, "uint32"):
    _DTYPES[torch.uint32] = _lib.U32


def desc(t: Optional[torch.Tensor], device: Optional[int] = None) -> Optional[Tensor]:
    """Describe a torch tensor (pointer already includes the storage offset)."""
    if t is None:
        return None
    if t.dim() > _lib.MAX_RANK:
        raise AtomaError(f"rank {t.dim()} tensors are not supported")
    d = Tensor()
    d.data = t.data_ptr()
    d.dtype = _DTYPES[t.dtype]
    if device is None:
        device = t.device.index if t.is_cuda else _lib.DEVICE_CPU
        if t.is_cuda and device is None:
            device = torch.cuda.current_device()
    d.device = device
    d.rank = t.dim()
    for i in range(t.dim()):
        d.shape[i] = t.shape[i]
        d.stride[i] = t.stride(i)
    return d


def spec(shape: Sequence[int], dtype: int, device: int = 0, stride: Optional[Sequence[int]] = None,
         data: int = 0x1000) -> Tensor:
    """A descriptor without memory behind it (plan-only calls: argument-derivation tests)."""
    d = Tensor()
    d.data = data
    d.dtype = dtype
    d.device = device
    d.rank = len(shape)
    if stride is None:
        stride, acc = [0] * len(shape), 1
        for i in reversed(range(len(shape))):
            stride[i] = acc
            acc *= shape[i]
    for i, (s, st) in enumerate(zip(shape, stride)):
        d.shape[i] = s
        d.stride[i] = st
    return d


def _as_desc(t) -> Optional[Tensor]:
    if t is None or isinstance(t, Tensor):
        return t
    return desc(t)


def _ref(d: Optional[Tensor]):
    return C.byref(d) if d is not None else None


def _check(msg: Optional[bytes]) -> None:
    if msg:
        raise AtomaError(msg.decode())


def _opt(v: Optional[int]) -> int:
    return -1 if v is None else int(v)


def _stream(t: Optional[torch.Tensor] = None) -> int:
    if torch.cuda.is_available():
        return torch.cuda.current_stream().cuda_stream
    return 0


# ------------------------------------------------------------------------------------------------
# plans (argument derivation; run anywhere, no GPU needed)
# ------------------------------------------------------------------------------------------------
def flash_attn_plan(q, k, v, softmax_scale: float, window_size_left=None, window_size_right=None,
                    alibi_slopes=None, softcap: Optional[float] = None, num_sms: int = 0) -> MhaPlan:
    plan = MhaPlan()
    qd, kd, vd, ad = map(_as_desc, (q, k, v, alibi_slopes))
    _check(_lib.lib().atoma_flash_attn_plan(_ref(qd), _ref(kd), _ref(vd), _ref(ad), softmax_scale,
                                            _opt(window_size_left), _opt(window_size_right),
                                            softcap or 0.0, num_sms, C.byref(plan)))
    return plan


def flash_attn_varlen_plan(q, k, v, seqlens_q, seqlens_k, max_seqlen_q: int, max_seqlen_k: int,
                           softmax_scale: float, window_size_left=None, window_size_right=None,
                           alibi_slopes=None, seqused_k=None, block_table=None,
                           softcap: Optional[float] = None, num_sms: int = 0) -> MhaPlan:
    plan = MhaPlan()
    ds = list(map(_as_desc, (q, k, v, alibi_slopes, seqlens_q, seqlens_k, seqused_k, block_table)))
    _check(_lib.lib().atoma_flash_attn_varlen_plan(*[_ref(d) for d in ds], max_seqlen_q, max_seqlen_k,
                                                   softmax_scale, _opt(window_size_left),
                                                   _opt(window_size_right), softcap or 0.0, num_sms,
                                                   C.byref(plan)))
    return plan


def flash_attn_kv_cache_plan(q, k_cache, v_cache, softmax_scale: float, alibi_slopes=None,
                             block_table=None, seqlens_k=None, window_size_left=None,
                             window_size_right=None, num_sms: int = 0) -> MhaPlan:
    plan = MhaPlan()
    ds = list(map(_as_desc, (q, k_cache, v_cache, alibi_slopes, block_table, seqlens_k)))
    _check(_lib.lib().atoma_flash_attn_kv_cache_plan(*[_ref(d) for d in ds], softmax_scale,
                                                     _opt(window_size_left), _opt(window_size_right),
                                                     num_sms, C.byref(plan)))
    return plan


def compute_num_splits(batch_size, num_heads, head_size, max_seqlen_k, max_seqlen_q, num_sms) -> int:
    return _lib.lib().atoma_compute_num_splits(batch_size, num_heads, head_size, max_seqlen_k,
                                               max_seqlen_q, num_sms)


def _execute(plan: MhaPlan, like: torch.Tensor, *, legacy_stream: bool = False,
             with_lse: bool = False):
    shape = [plan.out_shape[i] for i in range(plan.out_rank)]
    out = torch.empty(shape, dtype=like.dtype, device=like.device)  # uninitialised, like the reference
    lse = None
    if with_lse:
        lse = torch.zeros(plan.lse_elems, dtype=torch.float32, device=like.device)
    if legacy_stream:
        torch.cuda.current_stream().synchronize()
    _check(_lib.lib().atoma_mha_execute(C.byref(plan), out.data_ptr(), lse.data_ptr() if with_lse else None,
                                        None, None, 1 if legacy_stream else 0, _stream()))
    if legacy_stream:
        torch.cuda.synchronize()
    return (out, lse) if with_lse else out


# ------------------------------------------------------------------------------------------------
# csrc::flash_attn*  (csrc/src/lib.rs:392-571)
# ------------------------------------------------------------------------------------------------
def flash_attn(q, k, v, softmax_scale: float, causal: bool, **kw):
    """q [b, sq, h, d]; k, v [b, sk, hk, d] -> [b, sq, h, d]   (lib.rs:392-410)"""
    plan = flash_attn_plan(q, k, v, softmax_scale, None, 0 if causal else None)
    return _execute(plan, q, **kw)


def flash_attn_windowed(q, k, v, softmax_scale, window_size_left, window_size_right, **kw):
    plan = flash_attn_plan(q, k, v, softmax_scale, window_size_left, window_size_right)
    return _execute(plan, q, **kw)


def flash_attn_alibi(q, k, v, alibi_slopes, softmax_scale, causal: bool, **kw):
    plan = flash_attn_plan(q, k, v, softmax_scale, None, 0 if causal else None, alibi_slopes=alibi_slopes)
    return _execute(plan, q, **kw)


# ------------------------------------------------------------------------------------------------
# csrc::flash_attn_varlen*  (csrc/src/lib.rs:1160-1501)
# ------------------------------------------------------------------------------------------------
def flash_attn_varlen(q, k, v, seqlens_q, seqlens_k, max_seqlen_q: int, max_seqlen_k: int,
                      softmax_scale: float, causal: bool, **kw):
    """q [total_q, h, d]; k, v [total_k, hk, d]; cumulative u32 seqlens [b+1]   (lib.rs:1160-1188)"""
    plan = flash_attn_varlen_plan(q, k, v, seqlens_q, seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale,
                                  None, 0 if causal else None)
    return _execute(plan, q, **kw)


def flash_attn_varlen_with_block_table(q, k, v, alibi_slopes, seqlens_q, seqlens_k, max_seqlen_q: int,
                                       max_seqlen_k: int, softmax_scale: float, window_size_left,
                                       window_size_right, block_table, **kw):
    """k, v [num_blocks, page, hk, d] gathered through block_table [b, max_blocks]  (lib.rs:1392-1420)"""
    plan = flash_attn_varlen_plan(q, k, v, seqlens_q, seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale,
                                  window_size_left, window_size_right, alibi_slopes=alibi_slopes,
                                  block_table=block_table)
    return _execute(plan, q, **kw)


# ------------------------------------------------------------------------------------------------
# csrc::flash_attn_kv_cache*  (csrc/src/lib.rs:1907-2105)
# ------------------------------------------------------------------------------------------------
def flash_attn_kv_cache(q, k, v, softmax_scale: float, causal: bool, **kw):
    plan = flash_attn_kv_cache_plan(q, k, v, softmax_scale, window_size_left=None,
                                    window_size_right=0 if causal else None)
    return _execute(plan, q, **kw)


def flash_attn_kv_cache_full(q, k, v, alibi_slopes, softmax_scale: float, block_table, seqlens_k,
                             causal: bool, **kw):
    """q [b, sq, h, d]; caches [num_blocks, page, hk, d] (+block_table) or [b, sk, hk, d];
    seqlens_k u32 [b] per-sequence lengths   (lib.rs:2083-2105)"""
    plan = flash_attn_kv_cache_plan(q, k, v, softmax_scale, alibi_slopes=alibi_slopes, block_table=block_table,
                                    seqlens_k=seqlens_k, window_size_left=None,
                                    window_size_right=0 if causal else None)
    return _execute(plan, q, **kw)


# ------------------------------------------------------------------------------------------------
# cache ops  (csrc/src/cache_manager.rs)
# ------------------------------------------------------------------------------------------------
def reshape_and_cache_flash(key, value, key_cache, value_cache, slot_mapping) -> None:
    """cache_manager.rs:319-535"""
    ds = [desc(t) for t in (key, value, key_cache, value_cache, slot_mapping)]
    _check(_lib.lib().atoma_reshape_and_cache_flash(*[_ref(d) for d in ds], _stream()))


def rope_reshape_and_cache(q, k, v, key_cache, value_cache, positions, cos, sin, slot_mapping=None) -> None:
    """Fused rotary embedding (in place on q and k) + KV-cache write: what
    CausalSelfAttention::forward does between the projections and the attention call
    (models/src/llama.rs:269-304 + flash_attention.rs:361).  q [T, H, D], k/v [T, Hkv, D] with a
    contiguous [heads, D] tail, positions i64 [T], cos/sin [max_pos, D/2] in the q dtype."""
    T, H, D = q.shape
    HK = k.shape[1]
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise AtomaError("Only support f16/bf16 dtypes and q, k, v must have same dtype")
    if cos.dtype != q.dtype or sin.dtype != q.dtype or cos.shape[-1] != D // 2 or sin.shape != cos.shape:
        raise AtomaError(f"cos/sin must be [max_pos, {D // 2}] tensors of the query dtype")
    if positions.dtype != torch.int64 or positions.numel() != T:
        raise AtomaError(f"input_positions must be of dtype i64 and hold {T} positions")
    for name, t, heads in (("q", q, H), ("k", k, HK), ("v", v, HK)):
        if t.stride(2) != 1 or t.stride(1) != D or t.shape[2] != D or t.shape[0] != T:
            raise AtomaError(f"{name} must be [num_tokens, heads, head_size] with contiguous heads")
    if v.shape[1] != HK:
        raise AtomaError("key and value must have the same number of heads")
    block_stride = block_size = 0
    kc = vc = sm = None
    if slot_mapping is not None:
        if key_cache.shape[2:] != (HK, D) or value_cache.shape != key_cache.shape:
            raise AtomaError(f"caches must be [num_blocks, block_size, {HK}, {D}]")
        if key_cache.stride(0) != value_cache.stride(0):
            raise AtomaError("Only support block_stride == value_cache.stride[0]")
        block_stride, block_size = key_cache.stride(0), key_cache.shape[1]
        kc, vc, sm = key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr()
    L = _lib.lib()
    L.atoma_rope_reshape_and_cache(q.data_ptr(), k.data_ptr(), v.data_ptr(), kc, vc, positions.data_ptr(),
                                   cos.data_ptr(), sin.data_ptr(), sm, T, H, HK, D, q.stride(0), k.stride(0),
                                   v.stride(0), cos.stride(0), block_stride, block_size,
                                   1 if q.dtype == torch.bfloat16 else 0, _stream())
    if L.atoma_last_status() != 0:
        raise AtomaError(L.atoma_last_error().decode())


def copy_blocks(key_caches: Sequence[torch.Tensor], value_caches: Sequence[torch.Tensor],
                block_mapping: torch.Tensor) -> None:
    """cache_manager.rs:148-307 — block_mapping: device i64 [num_pairs, 2]"""
    nk, nv = len(key_caches), len(value_caches)
    karr = (Tensor * max(nk, 1))(*[desc(t) for t in key_caches])
    varr = (Tensor * max(nv, 1))(*[desc(t) for t in value_caches])
    scratch = None
    if nk and key_caches[0].is_cuda:
        scratch = torch.empty(2 * max(nk, nv), dtype=torch.int64, device=key_caches[0].device)
    bm = desc(block_mapping)
    _check(_lib.lib().atoma_copy_blocks(karr, nk, varr, nv, _ref(bm),
                                        scratch.data_ptr() if scratch is not None else None, _stream()))
    if scratch is not None:
        scratch.record_stream(torch.cuda.current_stream())


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, block_mapping: Dict[int, int]) -> None:
    """cache_manager.rs:18-128 — dst[block d] = src[block s] for (s, d) in block_mapping"""
    n = len(block_mapping)
    sb = (C.c_int64 * max(n, 1))(*block_mapping.keys())
    db = (C.c_int64 * max(n, 1))(*block_mapping.values())
    sd, dd = desc(src), desc(dst)
    _check(_lib.lib().atoma_swap_blocks(_ref(sd), _ref(dd), sb, db, n, _stream()))
    if src.is_cuda != dst.is_cuda:
        torch.cuda.current_stream().synchronize()  # host buffers are usable on return, like the reference


def gemm_bf16_tn(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[T, N] = x[T, K] @ w[N, K]^T (bf16, fp32 accumulate): models/src/llama.rs:269-271, 311"""
    _mat("x", x, dtype=torch.bfloat16)
    T, K = x.shape
    _mat("w", w, cols=K, dtype=torch.bfloat16)
    N = w.shape[0]
    if out is None:
        out = torch.empty((T, N), dtype=x.dtype, device=x.device)
    _mat("out", out, T, N, x.dtype)
    _cuda(x=x, w=w, out=out)
    L = _lib.lib()
    L.atoma_gemm_bf16_tn(x.data_ptr(), w.data_ptr(), out.data_ptr(), T, N, K, x.stride(0), w.stride(0),
                         out.stride(0), _stream())
    if L.atoma_last_status() != 0:
        raise AtomaError(L.atoma_last_error().decode())
    return out


LAUNCH_PDL, LAUNCH_CONST_WEIGHTS = 1, 2
CALL_STABLE_CACHE = 1


def set_launch_mode(pdl: bool = False, const_weights: bool = False) -> None:
    """include/atoma_flash.h: atoma_set_launch_mode (process-wide: how kernels are launched, nothing about
    what any particular call may assume)."""
    _lib.lib().atoma_set_launch_mode((LAUNCH_PDL if pdl else 0) | (LAUNCH_CONST_WEIGHTS if const_weights else 0))


def promise_stable_cache() -> None:
    """include/atoma_flash.h: atoma_mha_call_flags(ATOMA_CALL_STABLE_CACHE) — for the NEXT attention call of
    this thread only: whatever is still in flight on the stream writes nothing but final-page rows of the
    sequences that call decodes."""
    _lib.lib().atoma_mha_call_flags(CALL_STABLE_CACHE)


def last_decode_mode() -> int:
    """bit 0: the last paged decode streamed KV early; bit 1: it ran behind a cache-maintenance op."""
    return _lib.lib().atoma_last_decode_mode()


def last_mha_path() -> str:
    return _lib.lib().atoma_last_mha_path().decode()


def launch_count() -> int:
    return _lib.lib().atoma_launch_count()


def reset_launch_count() -> None:
    _lib.lib().atoma_reset_launch_count()


# ------------------------------------------------------------------------------------------------
# decode-step operators (include/atoma_llama.h): what models/src/llama.rs issues as candle calls
# ------------------------------------------------------------------------------------------------
def _dt(t: torch.Tensor) -> int:
    if t.dtype not in (torch.float16, torch.bfloat16):
        raise AtomaError("Only support f16/bf16 dtypes")
    return 1 if t.dtype == torch.bfloat16 else 0


def _mat(name: str, t: torch.Tensor, rows: Optional[int] = None, cols: Optional[int] = None,
         dtype: Optional[torch.dtype] = None) -> None:
    """What every matrix argument of the decode-step operators must be — the C entries take a pointer and a ROW
    stride only, so a transposed or otherwise non-row-major view or a mismatched dtype would be read as something
    else without any error: refuse them here (candle's `Linear::forward` makes its operands contiguous
    instead; a silent copy is not this library's to make)."""
    if t.dim() != 2:
        raise AtomaError(f"{name} must be a 2-D tensor, got shape {tuple(t.shape)}")
    if (rows is not None and t.shape[0] != rows) or (cols is not None and t.shape[1] != cols):
        want = f"[{rows if rows is not None else '*'}, {cols if cols is not None else '*'}]"
        raise AtomaError(f"{name} must be {want}, got {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise AtomaError(f"{name} must be row-major (element stride 1 in its last dimension, got {t.stride(1)}): "
                         "pass .contiguous() of a transposed view")
    if dtype is not None and t.dtype != dtype:
        raise AtomaError(f"{name} must have dtype {dtype}, got {t.dtype}")


def _cuda(**named) -> None:
    """Last check of a wrapper (after shapes, layouts and dtypes): every tensor handed to a kernel lives on a GPU."""
    for name, t in named.items():
        if t is not None and not t.is_cuda:
            raise AtomaError(f"{name} must be a CUDA tensor (there is no CPU path)")


def _raise_on_status() -> None:
    L = _lib.lib()
    if L.atoma_last_status() != 0:
        raise AtomaError(L.atoma_last_error().decode())


def gemm_tn(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[T, N] = x[T, K] @ w[N, K]^T for f16/bf16 (`Linear::forward`, llama.rs:269-271, 311, 365, 476)."""
    _mat("x", x)
    T, K = x.shape
    _mat("w", w, cols=K, dtype=x.dtype)
    N = w.shape[0]
    if out is None:
        out = torch.empty((T, N), dtype=x.dtype, device=x.device)
    _mat("out", out, T, N, x.dtype)
    _cuda(x=x, w=w, out=out)
    _lib.lib().atoma_gemm_tn(x.data_ptr(), w.data_ptr(), out.data_ptr(), T, N, K, x.stride(0), w.stride(0),
                             out.stride(0), _dt(x), _stream())
    _raise_on_status()
    return out


def qkv_rope_cache(x, wqkv, num_heads: int, num_kv_heads: int, head_size: int, positions, cos, sin,
                   key_cache=None, value_cache=None, slot_mapping=None, out=None) -> torch.Tensor:
    """Fused q/k/v projection + RoPE + KV-cache write (llama.rs:269-304 + flash_attention.rs:361) ->
    qkv [T, (H + 2 Hkv) D] with q and k rotated."""
    _mat("x", x)
    T, K = x.shape
    N = (num_heads + 2 * num_kv_heads) * head_size
    if wqkv.dim() != 2 or wqkv.shape[0] != N or wqkv.shape[1] != K:
        raise AtomaError(f"wqkv must be [{N}, {K}], got {tuple(wqkv.shape)}")
    _mat("wqkv", wqkv, N, K, x.dtype)
    if positions.dtype != torch.int64 or positions.numel() != T:
        raise AtomaError(f"input_positions must be of dtype i64 and hold {T} positions")
    if cos.dtype != x.dtype or sin.dtype != x.dtype or cos.shape[-1] != head_size // 2 or sin.shape != cos.shape:
        raise AtomaError(f"cos/sin must be [max_pos, {head_size // 2}] tensors of the activation dtype")
    if out is None:
        out = torch.empty((T, N), dtype=x.dtype, device=x.device)
    _mat("out", out, T, N, x.dtype)
    _cuda(x=x, wqkv=wqkv, out=out, input_positions=positions, cos=cos, sin=sin, key_cache=key_cache, value_cache=value_cache,
          slot_mapping=slot_mapping)
    kc = vc = sm = None
    block_stride = block_size = 0
    if slot_mapping is not None and key_cache is not None:
        if tuple(key_cache.shape[2:]) != (num_kv_heads, head_size) or value_cache.shape != key_cache.shape:
            raise AtomaError(f"caches must be [num_blocks, block_size, {num_kv_heads}, {head_size}]")
        if key_cache.stride(0) != value_cache.stride(0):
            raise AtomaError("Only support block_stride == value_cache.stride[0]")
        if slot_mapping.dtype != torch.int64 or slot_mapping.numel() != T:
            raise AtomaError(f"slot_mapping must be of dtype i64 and hold {T} slots")
        kc, vc, sm = key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr()
        block_stride, block_size = key_cache.stride(0), key_cache.shape[1]
    _lib.lib().atoma_qkv_rope_cache(x.data_ptr(), wqkv.data_ptr(), out.data_ptr(), T, K, x.stride(0), wqkv.stride(0),
                                    out.stride(0), num_heads, num_kv_heads, head_size, positions.data_ptr(),
                                    cos.data_ptr(), sin.data_ptr(), cos.stride(0), sm, kc, vc, block_stride,
                                    block_size, _dt(x), _stream())
    _raise_on_status()
    return out


def gate_up_silu(x, w_gate_up, out=None) -> torch.Tensor:
    """silu(x Wg^T) * (x Wu^T) with w_gate_up = cat([Wg, Wu]) (llama.rs:364)."""
    _mat("x", x)
    T, K = x.shape
    _mat("w_gate_up", w_gate_up, cols=K, dtype=x.dtype)
    if w_gate_up.shape[0] % 2:
        raise AtomaError(f"w_gate_up must stack W_gate on W_up: an even number of rows, got {w_gate_up.shape[0]}")
    inter = w_gate_up.shape[0] // 2
    if out is None:
        out = torch.empty((T, inter), dtype=x.dtype, device=x.device)
    _mat("out", out, T, inter, x.dtype)
    _cuda(x=x, w_gate_up=w_gate_up, out=out)
    _lib.lib().atoma_gate_up_silu(x.data_ptr(), w_gate_up.data_ptr(), out.data_ptr(), T, inter, K, x.stride(0),
                                  w_gate_up.stride(0), out.stride(0), _dt(x), _stream())
    _raise_on_status()
    return out


def add_rmsnorm(x, residual, weight, eps: float, residual_out=None, normed=None):
    """h = x + residual (either may be None); returns (h, rmsnorm(h) * weight) — llama.rs:398-410."""
    like = x if x is not None else residual
    T, hidden = like.shape
    if residual_out is None:
        residual_out = torch.empty_like(like)
    if normed is None and weight is not None:
        normed = torch.empty_like(like)
    p = lambda t: t.data_ptr() if t is not None else None
    ld = lambda t: t.stride(0) if t is not None else hidden
    _lib.lib().atoma_add_rmsnorm(p(x), p(residual), p(residual_out), p(weight), p(normed), T, hidden, ld(x),
                                 ld(residual), ld(residual_out), ld(normed), eps, _dt(like), _stream())
    _raise_on_status()
    return residual_out, normed


def mlp_gate_up_down(x, w_gate_up, w_down, act=None, out=None):
    """`Mlp::forward` (llama.rs:362-366): down(silu(gate(x)) * up(x)) -> (y, act); one kernel when it qualifies."""
    _mat("x", x)
    T, hidden = x.shape
    _mat("w_gate_up", w_gate_up, cols=hidden, dtype=x.dtype)
    if w_gate_up.shape[0] % 2:
        raise AtomaError(f"w_gate_up must stack W_gate on W_up: an even number of rows, got {w_gate_up.shape[0]}")
    inter = w_gate_up.shape[0] // 2
    _mat("w_down", w_down, hidden, inter, x.dtype)
    if act is None:
        act = torch.empty((T, inter), dtype=x.dtype, device=x.device)
    if out is None:
        out = torch.empty((T, w_down.shape[0]), dtype=x.dtype, device=x.device)
    _mat("act", act, T, inter, x.dtype)
    _mat("out", out, T, hidden, x.dtype)
    _cuda(x=x, w_gate_up=w_gate_up, w_down=w_down, act=act, out=out)
    _lib.lib().atoma_mlp_gate_up_down(x.data_ptr(), w_gate_up.data_ptr(), act.data_ptr(), w_down.data_ptr(), out.data_ptr(), T, inter,
                                      hidden, x.stride(0), w_gate_up.stride(0), act.stride(0), w_down.stride(0), out.stride(0),
                                      _dt(x), _stream())
    _raise_on_status()
    return out, act


def gemm_tn_add_rmsnorm(x, w, residual, weight, eps: float, out=None, residual_out=None, normed=None):
    """y = x w^T; h = y + residual; normed = rmsnorm(h) * weight — one kernel (llama.rs:311 + 401-407).  Returns
    (y, h, normed); bit-identical to gemm_tn followed by add_rmsnorm."""
    _mat("x", x)
    T, K = x.shape
    _mat("w", w, cols=K, dtype=x.dtype)
    N = w.shape[0]
    if out is None:
        out = torch.empty((T, N), dtype=x.dtype, device=x.device)
    if residual_out is None:
        residual_out = torch.empty((T, N), dtype=x.dtype, device=x.device)
    for name, t in (("out", out), ("residual", residual), ("residual_out", residual_out), ("normed", normed)):
        if t is not None:
            _mat(name, t, T, N, x.dtype)
    _cuda(x=x, w=w, out=out, residual=residual, residual_out=residual_out, weight=weight, normed=normed)
    if normed is None and weight is not None:
        normed = torch.empty((T, N), dtype=x.dtype, device=x.device)
    p = lambda t: t.data_ptr() if t is not None else None
    ld = lambda t: t.stride(0) if t is not None else N
    _lib.lib().atoma_gemm_tn_add_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), T, N, K, x.stride(0), w.stride(0),
                                         out.stride(0), p(residual), p(residual_out), p(weight), p(normed), ld(residual),
                                         ld(residual_out), ld(normed), eps, _dt(x), _stream())
    _raise_on_status()
    return out, residual_out, normed


def embedding(tokens, table, out=None) -> torch.Tensor:
    """table[tokens] (llama.rs:470); tokens: int32/uint32 bit patterns [T]."""
    T = tokens.numel()
    if tokens.dtype not in (torch.int32, getattr(torch, "uint32", torch.int32)):
        raise AtomaError("token ids must be u32")
    _mat("table", table)
    vocab, hidden = table.shape
    if not tokens.is_contiguous():
        raise AtomaError("token ids must be a contiguous tensor")
    if out is None:
        out = torch.empty((T, hidden), dtype=table.dtype, device=table.device)
    _mat("out", out, T, hidden, table.dtype)
    _cuda(token_ids=tokens, table=table, out=out)
    _lib.lib().atoma_embedding(tokens.data_ptr(), table.data_ptr(), out.data_ptr(), T, hidden, vocab, table.stride(0),
                               out.stride(0), _dt(table), _stream())
    _raise_on_status()
    return out


def sample_greedy(logits, want_f32: bool = False):
    """(token ids int32 [T], their logits f32 [T], optional f32 copy of the logits) — model_executor.rs:227-262."""
    _mat("logits", logits)
    _cuda(logits=logits)
    T, vocab = logits.shape
    tok = torch.empty((T,), dtype=torch.int32, device=logits.device)
    val = torch.empty((T,), dtype=torch.float32, device=logits.device)
    f32 = torch.empty((T, vocab), dtype=torch.float32, device=logits.device) if want_f32 else None
    _lib.lib().atoma_sample_greedy(logits.data_ptr(), T, vocab, logits.stride(0), _dt(logits), tok.data_ptr(),
                                   val.data_ptr(), f32.data_ptr() if f32 is not None else None, vocab, _stream())
    _raise_on_status()
    return tok, val, f32
