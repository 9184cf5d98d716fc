"""GPU parity of the tcgen05 projection GEMM (atoma_gemm_bf16_tn, row a11 of SURVEY §8) through the
C ABI.  The reference sends these GEMMs to cuBLAS via candle and pins no values (parity
unpinned), so the oracle is fp64 numpy on the same bf16 inputs, with cuBLAS (torch.matmul) timed and
compared beside it.

Tolerance: bf16 in, fp32 accumulate, bf16 out — element-wise
|y - exact| <= 0.5 ulp_bf16(exact) + 1e-3 |exact| + 4e-6 * sum_k |x_k w_k|   (fp32 accumulation),
norm-wise relative error <= 1e-3 + the bf16 quantisation floor."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from helpers import quant_floor, randn_t, rel_err, t2f32, ulp16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(x, w, y, what):
    xe, we = t2f32(x).astype(np.float64), t2f32(w).astype(np.float64)
    exact = xe @ we.T
    mag = np.abs(xe) @ np.abs(we).T
    got = t2f32(y).astype(np.float64)
    assert np.isfinite(got).all(), what
    bound = 0.5 * ulp16(exact, True) + 1e-3 * np.abs(exact) + 4e-6 * mag
    diff = np.abs(got - exact)
    bad = diff > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} beyond tolerance, worst excess {(diff - bound).max():.3e}"
    r = rel_err(got, exact)
    assert r <= 1e-3 + quant_floor(exact, True), f"{what}: norm-wise {r:.3e}"
    return r


@pytest.mark.parametrize("T,N,K", [
    (32, 6144, 4096),    # Llama-3.1-8B fused qkv at decode batch 32
    (32, 4096, 4096),    # o_proj
    (1, 3072, 2048),     # Llama-3.2-1B single-stream decode (fused qkv)
    (16, 1280, 8192),    # Llama-3.1-70B TP=8 per-rank fused qkv
    (16, 8192, 1024),    # ... and its o_proj (row parallel)
    (128, 1024, 512),
    (5, 1000, 520),      # N not a multiple of 128, K not a multiple of 64
    (200, 256, 192),
    (600, 512, 1024),    # three activation chunks (256 + 256 + 88) in one launch
    (1000, 384, 320),    # ragged everywhere: T, N % 128, K % 64
])
def test_gemm_vs_fp64(T, N, K):
    x = randn_t((T, K), 1, torch.bfloat16, DEV)
    w = randn_t((N, K), 2, torch.bfloat16, DEV, scale=K ** -0.5)
    y = ops.gemm_bf16_tn(x, w)
    r = _check(x, w, y, f"gemm T={T} N={N} K={K}")
    rc = rel_err(t2f32(torch.matmul(x, w.t())), t2f32(x).astype(np.float64) @ t2f32(w).astype(np.float64).T)
    assert r <= 1.1 * rc + 1e-4, f"ours {r:.3e} vs cuBLAS {rc:.3e} from fp64"


def test_gemm_strided_views_and_determinism():
    """x and y as column slices of wider buffers (the fused qkv output feeds attention as strided
    views), and bit-identical results run to run (fixed-order split-K reduction)."""
    T, N, K = 32, 4096, 4096
    xw = randn_t((T, K + 64), 3, torch.bfloat16, DEV)
    x = xw[:, 32:32 + K]                       # 64-byte offset, ldx = K + 64
    w = randn_t((N, K), 4, torch.bfloat16, DEV, scale=K ** -0.5)
    yw = torch.full((T, N + 256), 7.0, dtype=torch.bfloat16, device=DEV)
    y = yw[:, 128:128 + N]
    ops.gemm_bf16_tn(x, w, out=y)
    _check(x, w, y, "strided")
    assert bool((yw[:, :128] == 7).all()) and bool((yw[:, 128 + N:] == 7).all()), "wrote outside the output view"
    y2 = ops.gemm_bf16_tn(x, w)
    y3 = ops.gemm_bf16_tn(x, w)
    assert torch.equal(y2, y3) and torch.equal(y2, y.contiguous())


def test_gemm_exact_small_integers():
    """Integer-valued inputs with exactly representable sums: the result must be exact, whatever the
    split-K partition."""
    T, N, K = 32, 512, 4096
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randint(-2, 3, (T, K), generator=g).to(torch.bfloat16).to(DEV)
    w = torch.randint(-2, 3, (N, K), generator=g).to(torch.bfloat16).to(DEV)
    y = ops.gemm_bf16_tn(x, w)
    exact = x.double().cpu() @ w.double().cpu().T            # |.| <= 4 * 4096, not all bf16-exact
    want = exact.to(torch.float32).to(torch.bfloat16)
    assert torch.equal(y.cpu(), want)
