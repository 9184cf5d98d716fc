"""Shared test utilities: seeded inputs, the reference's golden-input recipe, the harness that drives
the reference's own kernels (oracle/_ref/libatoma_ref.so) and comparison helpers."""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = json.loads((ROOT / "tests" / "golden" / "attention_tables.json").read_text())
REF_SO = ROOT / "oracle" / "_ref" / "libatoma_ref.so"


def golden_inputs_np():
    """q, k, v as float32 arrays [heads=3, tokens=2, dim=8] built the way candle evaluates
    `arange(48).to_dtype(f16) / c`: an affine op whose multiplier is rounded to f16."""
    g = GOLDEN["inputs"]
    a = np.arange(g["numel"], dtype=np.float32).astype(np.float16)

    def div(c):
        return (a * np.float16(1.0 / c)).astype(np.float16).astype(np.float32).reshape(g["heads"], g["tokens"], g["dim"])

    return div(g["q_div"]), div(g["k_div"]), div(g["v_div"])


def arange_inputs_np(n, shape, dtype=np.float16):
    a = np.arange(n, dtype=np.float32).astype(dtype)

    def div(c):
        return (a * dtype(1.0 / c)).astype(dtype).astype(np.float32).reshape(shape)

    return div(30.0), div(40.0), div(50.0)


def round4(x):
    """Rust `f32::round(t * 1e4) / 1e4` (half away from zero) on float32."""
    x = np.asarray(x, dtype=np.float32)
    b = np.float32(1e4)
    y = x * b
    r = np.where(y >= 0, np.floor(y + np.float32(0.5)), np.ceil(y - np.float32(0.5))).astype(np.float32)
    return r / b


def rng_for(seed):
    return np.random.default_rng(seed)


def randn_t(shape, seed, dtype, device, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float32) * scale).to(dtype).to(device)


def t2f32(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def bits(t: torch.Tensor) -> np.ndarray:
    """Storage bits of a 16-bit tensor as uint16 numpy (contiguous copy)."""
    return t.detach().contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def ulp16(x: np.ndarray, is_bf16: bool) -> np.ndarray:
    """Spacing of the 16-bit type at |x| (float64)."""
    x = np.abs(np.asarray(x, dtype=np.float64))
    mant = 7 if is_bf16 else 10
    tiny = 2.0 ** (-126 if is_bf16 else -14)
    e = np.floor(np.log2(np.maximum(x, tiny)))
    return 2.0 ** (e - mant)


def assert_close_16(actual: np.ndarray, expected64: np.ndarray, is_bf16: bool, what: str,
                    norm_tol: float = 1e-3, ulps: float = 1.0, atol: float = 2e-3):
    """The parity bar: norm-wise relative error <= 1e-3 (the north star's
    tolerance) AND every element within `ulps` units in the last place of the 16-bit type (+ a small
    absolute floor for outputs near zero, where cancellation dominates)."""
    actual = np.asarray(actual, dtype=np.float64)
    expected64 = np.asarray(expected64, dtype=np.float64)
    assert actual.shape == expected64.shape, (what, actual.shape, expected64.shape)
    assert np.isfinite(actual).all(), f"{what}: non-finite output"
    r = rel_err(actual, expected64)
    assert r <= norm_tol, f"{what}: norm-wise rel err {r:.3e} > {norm_tol}"
    bound = ulps * ulp16(expected64, is_bf16) + atol * np.maximum(1e-3, np.abs(expected64).max()) * 1e-1
    bad = np.abs(actual - expected64) > bound
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} elements beyond {ulps} ulp; "
                           f"max abs diff {np.abs(actual - expected64).max():.3e}")


# ------------------------------------------------------------------------------------------------
# The reference's own kernels (compiled from /root/reference into oracle/_ref by oracle/ref_build)
# ------------------------------------------------------------------------------------------------
class RefLib:
    def __init__(self):
        from atoma_infer_b200 import _lib
        self.lib = C.CDLL(str(REF_SO), mode=C.RTLD_LOCAL)
        self.lib.run_mha.restype = None
        self.lib.run_mha.argtypes = _lib.SIGNATURES["run_mha"][1]
        for n in ("copy_blocks_f16", "copy_blocks_bf16", "reshape_and_cache_flash"):
            getattr(self.lib, n).restype = None
            getattr(self.lib, n).argtypes = _lib.SIGNATURES[n][1]

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def run_mha_plan(self, plan, out: torch.Tensor, lse: torch.Tensor, lseaccum=None, oaccum=None):
        """Call the reference run_mha with the argument tuple of a plan (legacy stream 0)."""
        a = plan.args
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
        torch.cuda.synchronize()

    def execute(self, plan, like: torch.Tensor) -> torch.Tensor:
        """Allocate what the Rust wrapper allocates (csrc/src/lib.rs:1726-1729, 1785-1800) and run."""
        shape = [plan.out_shape[i] for i in range(plan.out_rank)]
        out = torch.empty(shape, dtype=like.dtype, device=like.device)
        lse = torch.zeros(max(int(plan.lse_elems), 1), dtype=torch.float32, device=like.device)
        lacc = oacc = None
        if plan.args.num_splits > 1:
            lacc = torch.zeros(int(plan.lseaccum_elems), dtype=torch.float32, device=like.device)
            oacc = torch.zeros(int(plan.oaccum_elems), dtype=torch.float32, device=like.device)
        self.run_mha_plan(plan, out, lse, lacc, oacc)
        return out


def make_paged_cache(num_blocks, page, hk, d, dtype, device, seed, poison_unused=False):
    """One [2, NB, page, hk, d] allocation, like the reference's per-layer cache tensor
    (models/src/flash_attention.rs:239-246)."""
    kv = randn_t((2, num_blocks, page, hk, d), seed, dtype, device)
    return kv


def random_block_table(batch, max_blocks, num_blocks, seed, device):
    """Distinct physical blocks in random order (no sharing), as SURVEY §8d prescribes."""
    rng = rng_for(seed)
    perm = rng.permutation(num_blocks)[: batch * max_blocks].astype(np.int32)
    assert perm.size == batch * max_blocks, "not enough blocks"
    return torch.from_numpy(perm.reshape(batch, max_blocks)).to(device)
