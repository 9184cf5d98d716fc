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


def quant_floor(expected64: np.ndarray, is_bf16: bool) -> float:
    """Norm-wise relative error of the CORRECTLY ROUNDED 16-bit image of `expected64`: the error any
    implementation with a 16-bit output carries (about 1.6e-3 for bf16, 2e-4 for f16)."""
    from oracle import attention_oracle as orc
    x = np.asarray(expected64, dtype=np.float64)
    r = orc.round_to_io(x.astype(np.float32), is_bf16).astype(np.float64)
    return rel_err(r, x)


# Half-ulp of the 16-bit io type: the relative error of every P element, which the reference
# rounds to the io type before P.V (flash_fwd_kernel.h:949-954) and so do we.
EPS_P = {True: 2.0 ** -9, False: 2.0 ** -12}


def p_noise(exact64: np.ndarray, is_bf16: bool, sigmas: float = 8.0) -> np.ndarray:
    """Per-element allowance for the rounding of P before P.V.  The error of one output element is
    sum_t p_t e_t v_t / sum_t p_t with independent |e_t| <= EPS_P: standard deviation
    EPS_P / sqrt(3) * sqrt(sum p_t^2 v_t^2) / sum p_t, i.e. EPS_P / sqrt(3) times the typical
    magnitude of an output element of that row.  `sigmas` of it, with the row's rms as the typical
    magnitude (rows = last axis); 8 sigma because the tests compare up to 5e6 elements and the row rms
    is itself an estimate from D samples."""
    x = np.asarray(exact64, dtype=np.float64)
    rms = np.sqrt(np.mean(x ** 2, axis=-1, keepdims=True)) if x.size else x
    return sigmas * EPS_P[is_bf16] / np.sqrt(3.0) * np.broadcast_to(rms, x.shape)


def _worst(diff, bound, ref):
    i = np.unravel_index(int(np.argmax(diff - bound)), diff.shape)
    return f"worst excess {(diff - bound)[i]:.3e} at index {i} where |exact|={abs(ref[i]):.3e}, diff={diff[i]:.3e}"


def assert_close_16(actual: np.ndarray, exact64: np.ndarray, is_bf16: bool, what: str,
                    tol: float = 1e-3, atol: float = 1e-5):
    """The parity bar against the fp64 oracle (DESIGN.md "Parity"): the north star's 1e-3 relative
    tolerance on top of the two roundings that the reference's algorithm itself performs and that no
    16-bit implementation of it avoids — the output format's and P's:

      norm-wise     ||out - exact|| / ||exact|| <= tol + quant_floor(exact)
      element-wise  |out - exact| <= tol * |exact| + 0.5 ulp_16(exact) + p_noise(exact) + atol

    `exact64` is the un-rounded fp64 oracle output."""
    actual = np.asarray(actual, dtype=np.float64)
    exact64 = np.asarray(exact64, dtype=np.float64)
    assert actual.shape == exact64.shape, (what, actual.shape, exact64.shape)
    assert np.isfinite(actual).all(), f"{what}: non-finite output"
    r, floor = rel_err(actual, exact64), quant_floor(exact64, is_bf16)
    assert r <= tol + floor, f"{what}: norm-wise rel err {r:.3e} > {tol} + quantisation floor {floor:.3e}"
    bound = tol * np.abs(exact64) + 0.5 * ulp16(exact64, is_bf16) + p_noise(exact64, is_bf16) + atol
    diff = np.abs(actual - exact64)
    bad = diff > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} elements beyond tol; {_worst(diff, bound, exact64)}"
    return r, floor


def assert_same_up_to_rounding(a: np.ndarray, b: np.ndarray, exact64: np.ndarray, is_bf16: bool, what: str,
                               tol: float = 1e-3, slack: float = 1.10):
    """Two 16-bit implementations of the same fp32 algorithm (ours = a, the reference's kernels = b)
    that differ in tile order and split policy, hence in every rounding:

      norm-wise     ||a - b|| / ||b|| <= tol + sqrt(2) * quant_floor        (two output roundings)
      element-wise  |a - b| <= tol * |b| + 1 ulp_16 + sqrt(2) * p_noise     (two roundings of P each)
      accuracy      ours is no further from the exact result than the reference is (within `slack`)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    exact64 = np.asarray(exact64, dtype=np.float64)
    assert np.isfinite(a).all(), f"{what}: non-finite output"
    r, floor = rel_err(a, b), quant_floor(exact64, is_bf16)
    assert r <= tol + np.sqrt(2.0) * floor, f"{what}: norm-wise rel diff {r:.3e} vs the reference kernels"
    bound = (tol * np.abs(b) + ulp16(np.maximum(np.abs(a), np.abs(b)), is_bf16) +
             np.sqrt(2.0) * p_noise(exact64, is_bf16) + 1e-5)
    diff = np.abs(a - b)
    bad = diff > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} elements apart; {_worst(diff, bound, exact64)}"
    ra, rb = rel_err(a, exact64), rel_err(b, exact64)
    assert ra <= slack * rb + 1e-4, f"{what}: ours {ra:.3e} vs reference {rb:.3e} from the exact result"
    return ra, rb


# The reference's own kernels (compiled from /root/reference into oracle/_ref by oracle/ref_build)
from oracle.ref_lib import REF_SO, RefLib  # noqa: E402,F401


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
