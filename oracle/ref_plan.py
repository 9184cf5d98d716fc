"""Pure-Python derivation of the `run_mha` argument tuple the reference's Rust wrapper
`flash_attn_kv_cache_full` builds (csrc/src/lib.rs:1523-1855), and the direct
call into the REFERENCE's kernels (oracle/_ref/libatoma_ref.so).  TEST INFRASTRUCTURE ONLY.

Exists so that `bench.py --impl reference` runs without loading anything of this repo's product: no
libflashattention.so, no atoma_infer_b200 operator — only the reference's own CUDA code, cuBLAS and
element-wise torch.  `tests/test_ref_plan.py` pins it field by field against the C++ restatement
(`atoma_flash_attn_kv_cache_plan`), which itself is pinned by SURVEY appendix A.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from pathlib import Path

from .attention_oracle import compute_num_splits

REF_SO = Path(__file__).resolve().parent / "_ref" / "libatoma_ref.so"

_RUN_MHA_ARGS = (
    [C.c_void_p] * 8 + [C.c_bool] + [C.c_uint32] * 5 + [C.c_uint32] * 4 + [C.c_uint32] * 4 + [C.c_uint32] +
    [C.c_uint32] * 5 + [C.c_float, C.c_float] + [C.c_void_p, C.c_uint32, C.c_int] + [C.c_void_p] +
    [C.c_uint32] * 4 + [C.c_int] * 4 + [C.c_float, C.c_bool, C.c_bool] + [C.c_void_p, C.c_void_p])


def round_multiple(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class KvCachePlan:
    """The live fields of the tuple (lib.rs:1801-1849) + what the wrapper allocates (:1726-1729, :1785-1800)."""
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
    num_splits: int
    b: int
    h: int
    h_k: int
    d: int
    d_rounded: int
    softmax_scale: float
    scale_softmax_log2: float
    block_table_batch_stride: int
    page_block_size: int
    seqlen_q: int
    seqlen_k: int
    seqlen_q_rounded: int
    seqlen_k_rounded: int
    is_bf16: int
    is_causal: int
    window_size_left: int
    window_size_right: int
    lse_elems: int
    lseaccum_elems: int
    oaccum_elems: int


def kv_cache_plan(q_shape, q_stride, kc_shape, kc_stride, vc_stride, bt_shape, softmax_scale: float, is_bf16: bool,
                  num_sms: int, causal: bool = True) -> KvCachePlan:
    """q [B, 1, H, D]; k_cache / v_cache [NB, page, Hkv, D]; block_table [B, MB]; seqlens_k [B]."""
    b, sq, h, d = q_shape
    nb, page, hk, _ = kc_shape
    mb = bt_shape[1]
    seqlen_k = mb * page                                         # :1591-1598 (an upper bound, not the true length)
    d_rounded = round_multiple(d, 32)                            # :1720-1721
    sq_r, sk_r = round_multiple(sq, 128), round_multiple(seqlen_k, 128)   # :1722-1723
    is_causal = 1 if (causal and sq != 1) else 0                 # :1627-1631 (no alibi)
    wl, wr = (seqlen_k, 0) if causal else (-1, -1)               # :1610-1641 (ignored by the kernels)
    splits = compute_num_splits(b, h, d, seqlen_k, sq, num_sms)  # :1772-1779
    return KvCachePlan(
        q_batch_stride=q_stride[0], k_batch_stride=kc_stride[0], v_batch_stride=vc_stride[0], o_batch_stride=sq * h * d,
        q_row_stride=q_stride[1], k_row_stride=kc_stride[1], v_row_stride=vc_stride[1], o_row_stride=h * d,
        q_head_stride=q_stride[2], k_head_stride=kc_stride[2], v_head_stride=vc_stride[2], o_head_stride=d,
        num_splits=splits, b=b, h=h, h_k=hk, d=d, d_rounded=d_rounded, softmax_scale=softmax_scale,
        scale_softmax_log2=softmax_scale * math.log2(math.e), block_table_batch_stride=mb, page_block_size=page,
        seqlen_q=sq, seqlen_k=seqlen_k, seqlen_q_rounded=sq_r, seqlen_k_rounded=sk_r, is_bf16=1 if is_bf16 else 0,
        is_causal=is_causal, window_size_left=wl, window_size_right=wr, lse_elems=b * h * sq,
        lseaccum_elems=splits * b * h * sq if splits > 1 else 0,
        oaccum_elems=splits * b * h * sq * d_rounded if splits > 1 else 0)


class RefKernels:
    """The reference's C entry points, bound with their own signatures (csrc/src/ffi.rs:4-101)."""

    def __init__(self):
        self.lib = C.CDLL(str(REF_SO), mode=C.RTLD_LOCAL)
        self.lib.run_mha.restype = None
        self.lib.run_mha.argtypes = _RUN_MHA_ARGS
        self.lib.reshape_and_cache_flash.restype = None
        # cache_manager.cu:10 re-typedefs uint32_t as `unsigned long`: ITS dtype parameter is 64 bits wide
        self.lib.reshape_and_cache_flash.argtypes = [C.c_void_p] * 5 + [C.c_int64] * 7 + [C.c_uint64, C.c_void_p]

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def run_kv_cache(self, p: KvCachePlan, q_ptr, k_ptr, v_ptr, o_ptr, lse_ptr, block_table_ptr, seqlens_k_ptr,
                     lseaccum_ptr=None, oaccum_ptr=None) -> None:
        """run_mha on the legacy default stream (flash_api.cu:157), arguments in ffi.rs order."""
        self.lib.run_mha(
            q_ptr, k_ptr, v_ptr, o_ptr, lse_ptr, None, None, seqlens_k_ptr, False,
            p.q_batch_stride, p.k_batch_stride, p.v_batch_stride, p.o_batch_stride, 0,
            p.q_row_stride, p.k_row_stride, p.v_row_stride, p.o_row_stride,
            p.q_head_stride, p.k_head_stride, p.v_head_stride, p.o_head_stride,
            p.num_splits, p.b, p.h, p.h_k, p.d, p.d_rounded, p.softmax_scale, p.scale_softmax_log2,
            block_table_ptr, p.block_table_batch_stride, p.page_block_size, None,
            p.seqlen_q, p.seqlen_k, p.seqlen_q_rounded, p.seqlen_k_rounded, p.is_bf16, p.is_causal,
            p.window_size_left, p.window_size_right, 0.0, False, True, lseaccum_ptr, oaccum_ptr)
