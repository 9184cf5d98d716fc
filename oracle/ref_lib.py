"""ctypes harness for the REFERENCE's own CUDA kernels (oracle/_ref/libatoma_ref.so, built from
/root/reference/csrc/kernels by oracle/ref_build/Makefile).  TEST INFRASTRUCTURE ONLY: imported by
tests/ (parity against the reference itself) and by bench.py's `--impl reference` arm; the product
never loads it."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

REF_SO = Path(__file__).resolve().parent / "_ref" / "libatoma_ref.so"


class RefLib:
    def __init__(self):
        from atoma_infer_b200 import _lib
        self.lib = C.CDLL(str(REF_SO), mode=C.RTLD_LOCAL)
        self.lib.run_mha.restype = None
        self.lib.run_mha.argtypes = _lib.SIGNATURES["run_mha"][1]
        for n in ("copy_blocks_f16", "copy_blocks_bf16", "reshape_and_cache_flash"):
            args = list(_lib.SIGNATURES[n][1])
            if n == "reshape_and_cache_flash":
                # The reference's cache_manager.cu:10 re-typedefs uint32_t as `unsigned long`, so ITS
                # `dtype` parameter is 64 bits wide; pass it as such or the upper half is whatever
                # the stack slot held and neither dtype branch runs.
                args[12] = C.c_uint64
            getattr(self.lib, n).restype = None
            getattr(self.lib, n).argtypes = args

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def run_mha_plan(self, plan, out: torch.Tensor, lse: torch.Tensor, lseaccum=None, oaccum=None, sync=True):
        """Call the reference run_mha with the argument tuple of a plan (legacy stream 0).  sync=False
        is for callers that already run everything on the legacy default stream (bench.py)."""
        a = plan.args
        if sync:
            torch.cuda.synchronize()
        self.lib.run_mha(
            a.q_ptr, a.k_ptr, a.v_ptr, out.data_ptr(), lse.data_ptr(), a.alibi_slopes_ptr,
            a.cu_seqlens_q_ptr, a.cu_seqlens_k_ptr, a.is_seqlens_k_cumulative,
            a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.o_batch_stride, a.alibi_slopes_batch_stride,
            a.q_row_stride, a.k_row_stride, a.v_row_stride, a.o_row_stride,
            a.q_head_stride, a.k_head_stride, a.v_head_stride, a.o_head_stride,
            a.num_splits, a.b, a.h, a.h_k, a.d, a.d_rounded, a.softmax_scale, a.scale_softmax_log2,
            a.block_table, a.block_table_batch_stride, a.page_block_size, a.seqused_k,
            a.seqlen_q, a.seqlen_k, a.seqlen_q_rounded, a.seqlen_k_rounded, a.is_bf16, a.is_causal,
            a.window_size_left, a.window_size_right, a.softcap, a.unpadded_lse, a.force_split_kernel,
            lseaccum.data_ptr() if lseaccum is not None else None,
            oaccum.data_ptr() if oaccum is not None else None)
        if sync:
            torch.cuda.synchronize()

    def execute(self, plan, like: torch.Tensor, sync=True) -> torch.Tensor:
        """Allocate what the Rust wrapper allocates (csrc/src/lib.rs:1726-1729, 1785-1800) and run."""
        shape = [plan.out_shape[i] for i in range(plan.out_rank)]
        out = torch.empty(shape, dtype=like.dtype, device=like.device)
        lse = torch.zeros(max(int(plan.lse_elems), 1), dtype=torch.float32, device=like.device)
        lacc = oacc = None
        if plan.args.num_splits > 1:
            lacc = torch.zeros(int(plan.lseaccum_elems), dtype=torch.float32, device=like.device)
            oacc = torch.zeros(int(plan.oaccum_elems), dtype=torch.float32, device=like.device)
        self.run_mha_plan(plan, out, lse, lacc, oacc, sync=sync)
        return out
