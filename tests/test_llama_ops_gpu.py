"""GPU parity of the decode-step operators and the step executor (include/atoma_llama.h; SURVEY §8 rows
a11, f1, f2, f3) through the C ABI, against oracle/llama_oracle.py (numpy fp64 restatement of
models/src/llama.rs; parity for the candle-side ops is unpinned in the reference — see that module's
header) and against each other (fused == unfused, bit for bit).

Tolerances (written where they are used):
  GEMM          |y - exact| <= 0.5 ulp_16(exact) + 1e-3 |exact| + 4e-6 sum_k |x_k w_k|; norm-wise 1e-3 + quantisation floor
  element-wise  one output rounding: |y - exact| <= 0.5 ulp_16 + 1e-3 |exact| (+ one input rounding where the op
                re-rounds a 16-bit intermediate)
  whole step    norm-wise relative error of the logits against the fp64 oracle <= 2e-2 (bf16 activations through
                2 layers: every materialised tensor carries 2^-9 relative rounding noise), and no further from
                it than the oracle's own bf16-rounded evaluation by more than 1.5x
"""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import _lib, ops
from atoma_infer_b200.llama import DecodeStep, LlamaConfig, LlamaShard
from helpers import quant_floor, randn_t, rel_err, t2f32, ulp16
from oracle import attention_oracle as ao
from oracle import llama_oracle as lo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF, F16 = torch.bfloat16, torch.float16


def _gemm_check(x, w, y, what, is_bf16=True):
    xe, we = t2f32(x).astype(np.float64), t2f32(w).astype(np.float64)
    exact = xe @ we.T
    mag = np.abs(xe) @ np.abs(we).T
    got = t2f32(y).astype(np.float64)
    assert np.isfinite(got).all(), what
    bound = 0.5 * ulp16(exact, is_bf16) + 1e-3 * np.abs(exact) + 4e-6 * mag
    diff = np.abs(got - exact)
    bad = diff > bound
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} beyond tolerance, worst excess {(diff - bound).max():.3e}"
    r = rel_err(got, exact)
    assert r <= 1e-3 + quant_floor(exact, is_bf16), f"{what}: norm-wise {r:.3e}"


# ------------------------------------------------------------------------------------------------
# decode-time GEMM (gemm_skinny.cu)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,N,K,dtype", [
    (32, 6144, 4096, BF),     # Llama-3.1-8B fused qkv, decode batch 32
    (32, 4096, 14336, BF),    # down_proj
    (32, 28672, 4096, BF),    # gate_up (plain epilogue)
    (1, 3072, 2048, BF),      # Llama-3.2-1B single stream
    (16, 8192, 1024, BF),     # 70B TP=8 o_proj
    (7, 1000, 512, BF),       # N % 8 != 0 at the tile edge is impossible (N % 8 == 0), ragged last tile
    (3, 1004, 192, BF),       # N % 8 == 4: scalar tail stores
    (64, 520, 64, BF),        # one k block, M = 64 fully used
    (65, 256, 128, BF),       # M = 128 path, 65 rows
    (128, 2048, 2048, BF),
    (100, 4096, 1024, F16),
    (32, 4096, 4096, F16),
])
def test_gemm_tn_vs_fp64(T, N, K, dtype):
    x = randn_t((T, K), 1, dtype, DEV)
    w = randn_t((N, K), 2, dtype, DEV, scale=K ** -0.5)
    y = ops.gemm_tn(x, w)
    _gemm_check(x, w, y, f"gemm_tn T={T} N={N} K={K} {dtype}", dtype == BF)


def test_gemm_tn_small_integers_are_exact_and_rows_do_not_leak():
    """Integer operands: every product and partial sum is exact in fp32 and the result is exact in bf16 —
    any indexing error (tile edges, swizzle, operand rows >= T read from stale shared memory) shows up as
    a wrong integer."""
    T, N, K = 37, 1544, 320
    g = torch.Generator().manual_seed(5)
    x = torch.randint(-2, 3, (T, K), generator=g).to(BF).to(DEV)
    w = torch.randint(-1, 2, (N, K), generator=g).to(BF).to(DEV)
    y = ops.gemm_tn(x, w)
    want = (x.float() @ w.float().t())
    assert torch.equal(y.float(), want)
    # strided x / y views (the fused qkv output feeds attention as column slices)
    xw = torch.zeros((T, K + 64), dtype=BF, device=DEV)
    xw[:, 32:32 + K] = x
    yw = torch.full((T, N + 16), 7.0, dtype=BF, device=DEV)
    ops.gemm_tn(xw[:, 32:32 + K], w, out=yw[:, 8:8 + N])
    assert torch.equal(yw[:, 8:8 + N].float(), want)
    assert (yw[:, :8] == 7).all() and (yw[:, 8 + N:] == 7).all()


def test_gemm_is_deterministic_and_graph_capturable():
    T, N, K = 32, 4096, 4096
    x = randn_t((T, K), 3, BF, DEV)
    w = randn_t((N, K), 4, BF, DEV, scale=K ** -0.5)
    y0 = ops.gemm_tn(x, w).clone()
    for _ in range(3):
        assert torch.equal(ops.gemm_tn(x, w), y0)
    ops.set_launch_mode(pdl=True, const_weights=True)
    try:
        y = torch.empty_like(y0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.gemm_tn(x, w, out=y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            ops.gemm_tn(x, w, out=y)
            ops.gemm_tn(x, w, out=y)
        y.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, y0)
    finally:
        ops.set_launch_mode()


def test_stream_k_gemm_counters_survive_interleaved_shapes():
    """Regression (round-1 advisor, high): the stream-K kernel's arrival counters and fp32 partials shared
    one workspace whose split moved with the shape, so a launch with more than 64 tiles after a small one
    found partials where it expected zeroed counters and silently skipped its split reductions."""
    L = _lib.lib()
    L.atoma_debug_gemm_algo(2)   # keep decode shapes on the stream-K kernel
    try:
        x32 = randn_t((32, 4096), 11, BF, DEV)
        w_small = randn_t((4096, 4096), 12, BF, DEV, scale=4096 ** -0.5)
        w_big = randn_t((28672, 4096), 13, BF, DEV, scale=4096 ** -0.5)
        x600 = randn_t((600, 4096), 14, BF, DEV)
        for _ in range(2):
            y1 = ops.gemm_bf16_tn(x32, w_small)
            y2 = ops.gemm_bf16_tn(x32, w_big)        # 224 tiles
            y3 = ops.gemm_bf16_tn(x600, w_small)     # 3 activation chunks x 32 tiles
            y4 = ops.gemm_bf16_tn(x32, w_small)
            torch.cuda.synchronize()
            _gemm_check(x32, w_small, y1, "stream-K small")
            _gemm_check(x32, w_big, y2, "stream-K N=28672 after a small launch")
            _gemm_check(x600, w_small, y3, "stream-K T=600 after a small launch")
            assert torch.equal(y1, y4)
    finally:
        L.atoma_debug_gemm_algo(0)


# ------------------------------------------------------------------------------------------------
# fused qkv + RoPE + cache write (row f1)
# ------------------------------------------------------------------------------------------------
def _rope_tables(D, n, dtype):
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().to(dtype).to(DEV), ang.sin().to(dtype).to(DEV)


@pytest.mark.parametrize("T,H,HK,D,K,dtype", [
    (32, 32, 8, 128, 4096, BF),    # Llama-3.1-8B
    (1, 32, 8, 64, 2048, BF),      # Llama-3.2-1B, single stream
    (16, 8, 1, 128, 8192, BF),     # 70B TP=8 per-rank
    (5, 4, 2, 64, 256, F16),
    (150, 8, 2, 128, 512, BF),     # > 128 tokens: two passes
])
def test_qkv_rope_cache_fused_is_bit_identical_to_unfused_and_matches_oracle(T, H, HK, D, K, dtype):
    is_bf16 = dtype == BF
    x = randn_t((T, K), 21, dtype, DEV)
    w = randn_t(((H + 2 * HK) * D, K), 22, dtype, DEV, scale=K ** -0.5)
    cos, sin = _rope_tables(D, 600, dtype)
    g = torch.Generator().manual_seed(23)
    pos = torch.randint(0, 600, (T,), generator=g).to(DEV)
    NB, page = (T + 15) // 16 + 3, 16
    slots = torch.randperm(NB * page, generator=g)[:T].to(torch.int64)
    slots[T // 2] = -1                                   # a padded token (cache_manager.cu:151-155)
    slots = slots.to(DEV)
    kv1 = randn_t((2, NB, page, HK, D), 24, dtype, DEV)
    kv2 = kv1.clone()
    fused = ops.qkv_rope_cache(x, w, H, HK, D, pos, cos, sin, kv1[0], kv1[1], slots)
    # unfused: plain projection (in the same 128-row passes: beyond 128 rows gemm_tn switches to the stream-K
    # kernel, whose accumulation order differs), then the RoPE + cache-write kernel of round 1
    qkv = torch.cat([ops.gemm_tn(x[t0:t0 + 128], w) for t0 in range(0, T, 128)], dim=0)
    plain = qkv.clone()
    nq, nkv = H * D, HK * D
    ops.rope_reshape_and_cache(qkv[:, :nq].unflatten(1, (H, D)), qkv[:, nq:nq + nkv].unflatten(1, (HK, D)),
                               qkv[:, nq + nkv:].unflatten(1, (HK, D)), kv2[0], kv2[1], pos, cos, sin, slots)
    torch.cuda.synchronize()
    assert torch.equal(fused, qkv), "fused qkv+rope differs from projection followed by rope"
    assert torch.equal(kv1, kv2), "fused cache write differs from reshape_and_cache_flash"
    # oracle: rope of the (rounded) projection in fp64, one rounding
    p = t2f32(plain).astype(np.float64)
    want_q = ao.rope(p[:, :nq].reshape(T, H, D), pos.cpu().numpy(), t2f32(cos), t2f32(sin)).reshape(T, nq)
    want_k = ao.rope(p[:, nq:nq + nkv].reshape(T, HK, D), pos.cpu().numpy(), t2f32(cos), t2f32(sin)).reshape(T, nkv)
    got = t2f32(fused).astype(np.float64)
    for name, a, b in (("q", got[:, :nq], want_q), ("k", got[:, nq:nq + nkv], want_k)):
        # one output rounding + fp32 (not fp64) arithmetic on two products
        bound = 0.5 * ulp16(b, is_bf16) + 1e-6 * (np.abs(b) + 1.0)
        assert (np.abs(a - b) <= bound * 1.0001 + 1e-7).all(), f"rope {name}: max excess {(np.abs(a - b) - bound).max():.3e}"
    assert np.array_equal(got[:, nq + nkv:], p[:, nq + nkv:]), "v must pass through unchanged"
    # the cache rows are where the slot mapping says, untouched elsewhere
    kc, vc = t2f32(kv1[0]).reshape(NB * page, HK * D), t2f32(kv1[1]).reshape(NB * page, HK * D)
    k0, v0 = t2f32(kv2[0]).reshape(NB * page, HK * D), t2f32(kv2[1]).reshape(NB * page, HK * D)
    for t, s in enumerate(slots.cpu().tolist()):
        if s >= 0:
            assert np.array_equal(kc[s], t2f32(fused)[t, nq:nq + nkv]) and np.array_equal(vc[s], t2f32(fused)[t, nq + nkv:])
    assert np.array_equal(kc, k0) and np.array_equal(vc, v0)


def test_qkv_rope_cache_without_cache_and_error_texts():
    T, H, HK, D, K = 4, 4, 2, 64, 128
    x = randn_t((T, K), 1, BF, DEV)
    w = randn_t(((H + 2 * HK) * D, K), 2, BF, DEV)
    cos, sin = _rope_tables(D, 32, BF)
    pos = torch.arange(T, device=DEV)
    out = ops.qkv_rope_cache(x, w, H, HK, D, pos, cos, sin)
    assert out.shape == (T, (H + 2 * HK) * D) and torch.isfinite(out.float()).all()
    with pytest.raises(ops.AtomaError, match="input_positions must be of dtype i64"):
        ops.qkv_rope_cache(x, w, H, HK, D, pos.int(), cos, sin)
    with pytest.raises(ops.AtomaError, match="wqkv must be"):
        ops.qkv_rope_cache(x, w[:-8], H, HK, D, pos, cos, sin)
    with pytest.raises(ops.AtomaError, match="K % 64"):
        ops.qkv_rope_cache(x[:, :96].contiguous(), w[:, :96].contiguous(), H, HK, D, pos, cos, sin)


# ------------------------------------------------------------------------------------------------
# fused gate_up + SwiGLU, add + RMSNorm, embedding, greedy sampling (row f3)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,I,K,dtype", [(32, 14336, 4096, BF), (1, 8192, 2048, BF), (16, 3584, 8192, BF),
                                         (9, 200, 128, F16), (130, 1024, 256, BF)])
def test_gate_up_silu(T, I, K, dtype):
    is_bf16 = dtype == BF
    x = randn_t((T, K), 31, dtype, DEV)
    w = randn_t((2 * I, K), 32, dtype, DEV, scale=K ** -0.5)
    y = ops.gate_up_silu(x, w)
    # the two projections, rounded to the model dtype (same kernel and 128-row passes as the fused op)
    gu = torch.cat([ops.gemm_tn(x[t0:t0 + 128], w) for t0 in range(0, T, 128)], dim=0)
    g, u = t2f32(gu[:, :I]).astype(np.float64), t2f32(gu[:, I:]).astype(np.float64)
    want = lo.silu(g) * u                                   # fp64 on the rounded projections
    got = t2f32(y).astype(np.float64)
    # one output rounding; __expf and fp32 arithmetic: 2e-6 relative
    bound = 0.5 * ulp16(want, is_bf16) + 4e-6 * np.abs(want) + 1e-30
    assert (np.abs(got - want) <= bound * 1.0001).all(), f"silu*up: max excess {(np.abs(got - want) - bound).max():.3e}"
    # and the reference's two-rounding evaluation (silu rounded, then the product) is within one more rounding
    ref2 = lo.silu_mul(g, u, is_bf16)
    assert (np.abs(got - ref2) <= 1.01 * ulp16(ref2, is_bf16) + 2.0 ** -8 * np.abs(ref2)).all()


@pytest.mark.parametrize("T,hidden,dtype", [(32, 4096, BF), (1, 2048, BF), (16, 8192, BF), (5, 264, F16), (3, 16384, BF)])
def test_add_rmsnorm(T, hidden, dtype):
    is_bf16 = dtype == BF
    x = randn_t((T, hidden), 41, dtype, DEV)
    r = randn_t((T, hidden), 42, dtype, DEV, scale=3.0)
    w = (1.0 + 0.1 * randn_t((hidden,), 43, torch.float32, DEV)).to(dtype)
    h, n = ops.add_rmsnorm(x, r, w, 1e-5)
    want_h = lo.add(t2f32(x), t2f32(r), is_bf16)
    assert np.array_equal(t2f32(h).astype(np.float64), want_h), "residual add must be the exact model-dtype add"
    want_n = lo.rms_norm(want_h, t2f32(w), 1e-5, is_bf16, exact=True)
    got = t2f32(n).astype(np.float64)
    bound = 0.5 * ulp16(want_n, is_bf16) + 4e-6 * np.abs(want_n) + 1e-30     # one rounding + fp32 statistics
    assert (np.abs(got - want_n) <= bound * 1.0001).all(), f"rmsnorm: max excess {(np.abs(got - want_n) - bound).max():.3e}"
    # no addend (first layer) and in-place residual
    h2, n2 = ops.add_rmsnorm(None, r, w, 1e-5)
    assert torch.equal(h2, r)
    r2 = r.clone()
    ops.add_rmsnorm(x, r2, w, 1e-5, residual_out=r2, normed=n2)
    assert torch.equal(r2, h) and torch.equal(n2, n)


@pytest.mark.parametrize("T,N,K,dtype", [(32, 4096, 4096, BF), (1, 2048, 512, BF), (7, 1024, 1344, F16), (64, 4096, 1792, BF),
                                          (128, 8192, 1024, BF), (100, 264, 192, BF), (16, 8192, 3584, F16)])
def test_gemm_tn_add_rmsnorm_is_bit_identical_to_the_two_calls(T, N, K, dtype):
    """The fused tail (the CTAs of the GEMM that finish last normalise its rows) against gemm_tn + add_rmsnorm:
    same bits for y, h and normed; its arrival counters reset themselves (repeated launches, graph replay); h may
    alias the residual (what the step does)."""
    x = randn_t((T, K), 51, dtype, DEV)
    w = randn_t((N, K), 52, dtype, DEV, scale=K ** -0.5)
    r = randn_t((T, N), 53, dtype, DEV, scale=2.0)
    nw = (1.0 + 0.1 * randn_t((N,), 54, torch.float32, DEV)).to(dtype)
    y0 = ops.gemm_tn(x, w)
    h0, n0 = ops.add_rmsnorm(y0, r, nw, 1e-5)
    for it in range(3):
        y, h, n = ops.gemm_tn_add_rmsnorm(x, w, r, nw, 1e-5)
        torch.cuda.synchronize()
        assert torch.equal(y, y0) and torch.equal(h, h0) and torch.equal(n, n0), f"launch {it}"
    r2, n2 = r.clone(), torch.empty_like(r)
    ops.gemm_tn_add_rmsnorm(x, w, r2, nw, 1e-5, residual_out=r2, normed=n2)      # in-place residual stream
    assert torch.equal(r2, h0) and torch.equal(n2, n0)
    # no norm weight: only the add;  no residual: h = y
    _, h3, n3 = ops.gemm_tn_add_rmsnorm(x, w, r, None, 1e-5)
    assert n3 is None and torch.equal(h3, h0)
    _, h4, n4 = ops.gemm_tn_add_rmsnorm(x, w, None, nw, 1e-5)
    h5, n5 = ops.add_rmsnorm(y0, None, nw, 1e-5)
    assert torch.equal(h4, h5) and torch.equal(n4, n5)
    # graph replay, programmatic dependent launch on, two fused GEMMs back to back
    ops.set_launch_mode(pdl=True, const_weights=True)
    try:
        ya, ha, na = torch.empty_like(y0), torch.empty_like(y0), torch.empty_like(y0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.gemm_tn_add_rmsnorm(x, w, r, nw, 1e-5, out=ya, residual_out=ha, normed=na)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            ops.gemm_tn_add_rmsnorm(x, w, r, nw, 1e-5, out=ya, residual_out=ha, normed=na)
            ops.gemm_tn_add_rmsnorm(x, w, r, nw, 1e-5, out=ya, residual_out=ha, normed=na)
        for _ in range(3):
            ya.zero_(); ha.zero_(); na.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(ya, y0) and torch.equal(ha, h0) and torch.equal(na, n0)
    finally:
        ops.set_launch_mode()


@pytest.mark.parametrize("T,hidden,inter,dtype", [(32, 4096, 14336, BF), (1, 2048, 8192, BF), (7, 512, 1024, F16), (64, 4096, 1792, BF),
                                                  (128, 1024, 2048, BF), (100, 8192, 3584, BF)])
def test_mlp_chain_is_bit_identical_to_the_two_kernels(T, hidden, inter, dtype):
    """gate_up + SwiGLU and down_proj chained in one kernel (needs programmatic dependent launch) against the two
    separate kernels: same bits for the activations and the output; its grid-wide arrival counter resets itself
    (repeated launches, graph replay with two chains back to back)."""
    x = randn_t((T, hidden), 61, dtype, DEV)
    wgu = randn_t((2 * inter, hidden), 62, dtype, DEV, scale=hidden ** -0.5)
    wd = randn_t((hidden, inter), 63, dtype, DEV, scale=inter ** -0.5)
    ops.reset_launch_count()
    act0 = ops.gate_up_silu(x, wgu)
    y0 = ops.gemm_tn(act0, wd)
    assert ops.launch_count() == 2
    ops.set_launch_mode(pdl=True, const_weights=True)
    try:
        for it in range(3):
            ops.reset_launch_count()
            y, act = ops.mlp_gate_up_down(x, wgu, wd)
            torch.cuda.synchronize()
            assert ops.launch_count() == 1, "the shapes qualify for the chained kernel"
            assert torch.equal(act, act0) and torch.equal(y, y0), f"launch {it}"
        ya, aa = torch.empty_like(y0), torch.empty_like(act0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.mlp_gate_up_down(x, wgu, wd, act=aa, out=ya)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            ops.mlp_gate_up_down(x, wgu, wd, act=aa, out=ya)
            ops.mlp_gate_up_down(x, wgu, wd, act=aa, out=ya)
        for _ in range(3):
            ya.zero_(); aa.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(aa, act0) and torch.equal(ya, y0)
        # the bounded grid barrier never expired on either stream (the sticky words of atoma_llama_step_status)
        for st in (torch.cuda.current_stream(), side):
            assert _lib.lib().atoma_llama_step_status(st.cuda_stream) == 0
    finally:
        ops.set_launch_mode()
    # without programmatic dependent launch the same entry runs the two kernels
    ops.reset_launch_count()
    y2, act2 = ops.mlp_gate_up_down(x, wgu, wd)
    assert ops.launch_count() == 2 and torch.equal(y2, y0) and torch.equal(act2, act0)


def test_gemm_tn_add_rmsnorm_errors():
    x = randn_t((4, 128), 1, BF, DEV)
    w = randn_t((16392, 128), 2, BF, DEV)
    with pytest.raises(ops.AtomaError, match="gemm_tn_add_rmsnorm"):
        ops.gemm_tn_add_rmsnorm(x, w, None, None, 1e-5)            # wider than a GEMM CTA can normalise
    xw = randn_t((130, 128), 3, BF, DEV)
    with pytest.raises(ops.AtomaError, match="T <= 128"):
        ops.gemm_tn_add_rmsnorm(xw, randn_t((256, 128), 4, BF, DEV), None, None, 1e-5)


def test_embedding_and_greedy_sampling():
    vocab, hidden, T = 1000, 264, 9
    table = randn_t((vocab, hidden), 51, BF, DEV)
    tok = torch.tensor([0, 999, 5, 5, 17, 123, 998, 1, 500], dtype=torch.int32, device=DEV)
    out = ops.embedding(tok, table)
    assert torch.equal(out, table[tok.long()])
    # greedy: LAST index among equal maxima (max_by), NaN above +inf under total order, f32 copy exact
    V = 128256
    logits = randn_t((T, V), 52, BF, DEV)
    logits[0, 100] = 50.0
    logits[0, 77777] = 50.0                      # tie -> 77777
    logits[1, :] = -float("inf")
    logits[1, 4242] = -1e30                       # everything else -inf
    logits[2, 9] = float("inf")
    logits[2, 128255] = float("inf")              # tie at +inf -> last
    logits[3, 3] = float("nan")                   # +NaN sorts above +inf
    logits[3, 10] = float("inf")
    logits[4, :] = 0.0                            # all equal -> last index
    ids, vals, f32 = ops.sample_greedy(logits, want_f32=True)
    want_ids, want_vals = lo.greedy(t2f32(logits))
    assert ids.cpu().tolist() == want_ids.tolist()
    assert ids[0].item() == 77777 and ids[1].item() == 4242 and ids[2].item() == 128255 and ids[3].item() == 3
    assert ids[4].item() == V - 1
    np.testing.assert_array_equal(vals.cpu().numpy()[[0, 1, 2, 4]], want_vals[[0, 1, 2, 4]])
    assert np.isnan(vals[3].item())
    assert torch.equal(f32.nan_to_num(1.0, 2.0, 3.0), logits.float().nan_to_num(1.0, 2.0, 3.0))
    # a vocabulary that is not a multiple of 8 (rows stay 16-byte aligned: a view of a padded buffer)
    small = randn_t((3, 1008), 53, BF, DEV)[:, :1003]
    ids2, _, _ = ops.sample_greedy(small)
    assert ids2.cpu().tolist() == lo.greedy(t2f32(small))[0].tolist()


# ------------------------------------------------------------------------------------------------
# the decode-step executor (rows f2 + f3)
# ------------------------------------------------------------------------------------------------
def _small_setup(T, dtype=BF, layers=2, with_head=True, seed=7):
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, vocab_size=2000, num_hidden_layers=layers,
                      num_attention_heads=8, num_key_value_heads=2, rms_norm_eps=1e-5, max_position_embeddings=256)
    shard = LlamaShard.random_init(cfg, dtype=dtype, device=DEV, seed=seed, gen_device="cpu", with_head=with_head)
    page, MB = 16, 6
    NB = T * MB + 2
    rng = np.random.default_rng(seed)
    kv = [randn_t((2, NB, page, 2, 64), 100 + l, dtype, DEV) for l in range(layers)]
    bt = torch.from_numpy(rng.permutation(NB)[: T * MB].astype(np.int32).reshape(T, MB)).to(DEV)
    lens = torch.from_numpy(rng.integers(1, MB * page + 1, size=T).astype(np.int32)).to(DEV)
    pos = (lens.long() - 1)
    slots = bt[torch.arange(T, device=DEV), (pos // page)].long() * page + pos % page     # worker.rs:392-401
    tokens = torch.from_numpy(rng.integers(0, cfg.vocab_size, size=T).astype(np.int32)).to(DEV)
    return cfg, shard, kv, bt, lens, pos, slots, tokens


def _oracle_layers(shard, cfg):
    out = []
    for ly in shard.layers:
        wq, wk, wv, wg, wu = lo.split_fused(t2f32(ly["wqkv"]), t2f32(ly["w_gate_up"]), cfg.num_attention_heads,
                                            cfg.num_key_value_heads, cfg.head_dim)
        out.append(lo.LayerWeights(wq, wk, wv, t2f32(ly["wo"]), wg, wu, t2f32(ly["w_down"]), t2f32(ly["rms_attn"]),
                                   t2f32(ly["rms_mlp"])))
    return out


@pytest.mark.parametrize("T,dtype", [(5, BF), (1, BF), (32, BF), (3, F16)])
def test_decode_step_matches_the_oracle_and_the_composition_of_its_operators(T, dtype):
    is_bf16 = dtype == BF
    cfg, shard, kv, bt, lens, pos, slots, tokens = _small_setup(T, dtype)
    kv_ops = [c.clone() for c in kv]
    kv_np = [t2f32(c).copy() for c in kv]
    kv_np_x = [c.copy() for c in kv_np]
    step = DecodeStep(shard, kv, max_batch=T, want_logits=True, want_logits_f32=True, want_hidden=True)
    step.bind(tokens, pos, slots, bt, lens)
    step.launch()
    torch.cuda.synchronize()
    assert step.launches_per_step() == 1 + 7 * cfg.num_hidden_layers + 3   # embedding, 7 per layer, final norm, lm_head, sampling

    # (1) the same step composed from the individual operators: bit-identical
    x = ops.embedding(tokens, shard.embed)
    resid, proj = x, None
    D, H, HK = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
    for l, ly in enumerate(shard.layers):
        resid, n = ops.add_rmsnorm(proj, resid, ly["rms_attn"], cfg.rms_norm_eps)
        qkv = ops.qkv_rope_cache(n, ly["wqkv"], H, HK, D, pos, shard.cos, shard.sin, kv_ops[l][0], kv_ops[l][1], slots)
        q = qkv[:, :H * D].unflatten(1, (H, D)).unsqueeze(1)
        o = ops.flash_attn_kv_cache_full(q, kv_ops[l][0], kv_ops[l][1], None, D ** -0.5, bt, lens, True)
        proj = ops.gemm_tn(o.reshape(T, H * D), ly["wo"])
        resid, n = ops.add_rmsnorm(proj, resid, ly["rms_mlp"], cfg.rms_norm_eps)
        proj = ops.gemm_tn(ops.gate_up_silu(n, ly["w_gate_up"]), ly["w_down"])
    _, xf = ops.add_rmsnorm(proj, resid, shard.norm_f, cfg.rms_norm_eps)
    logits = ops.gemm_tn(xf, shard.lm_head)
    ids, vals, _ = ops.sample_greedy(logits)
    torch.cuda.synchronize()
    assert torch.equal(step.hidden_out[:T], xf)
    assert torch.equal(step.logits[:T], logits)
    assert torch.equal(step.next_tokens[:T], ids) and torch.equal(step.next_logits[:T], vals)
    assert torch.equal(step.logits_f32[:T], logits.float())
    for l in range(cfg.num_hidden_layers):
        assert torch.equal(kv[l], kv_ops[l]), f"layer {l}: cache written differently"

    # (2) the oracle: fp64 throughout (exact) as the anchor, its model-dtype evaluation as the yardstick
    layers = _oracle_layers(shard, cfg)
    hidden = t2f32(shard.embed)[tokens.cpu().numpy()]
    args = (t2f32(shard.norm_f), t2f32(shard.lm_head), t2f32(shard.cos), t2f32(shard.sin), pos.cpu().numpy(),
            slots.cpu().numpy(), bt.cpu().numpy(), lens.cpu().numpy(), H, HK, D, cfg.rms_norm_eps, is_bf16)
    xf_exact, lg_exact = lo.decode_step(hidden, layers, kv_np_x, *args, exact=True)
    xf_model, lg_model = lo.decode_step(hidden, layers, kv_np, *args, exact=False)
    got = t2f32(step.logits[:T]).astype(np.float64)
    r_ours, r_model = rel_err(got, lg_exact), rel_err(lg_model, lg_exact)
    assert np.isfinite(got).all()
    assert r_ours <= 2e-2, f"logits vs fp64 oracle: {r_ours:.3e}"
    assert r_ours <= 1.5 * r_model + 1e-3, f"ours {r_ours:.3e} vs the oracle's model-dtype evaluation {r_model:.3e}"
    # the cache: untouched rows bit-identical, the rows written by the step (k after RoPE, v) equal to the oracle's
    # up to the rounding noise the layer's input carries (first layer: one ulp; deeper layers: norm-wise 3e-2)
    page = kv[0].shape[2]
    sl = slots.cpu().numpy()
    for l in range(cfg.num_hidden_layers):
        a = t2f32(kv[l]).astype(np.float64).reshape(2, -1, HK * D)
        b = kv_np[l].astype(np.float64).reshape(2, -1, HK * D)
        mask = np.ones(a.shape[1], bool)
        mask[sl[sl >= 0]] = False
        assert np.array_equal(a[:, mask], b[:, mask]), f"layer {l}: rows outside the slot mapping changed"
        wa, wb = a[:, ~mask], b[:, ~mask]
        assert rel_err(wa, wb) <= (4e-3 if l == 0 else 3e-2), f"layer {l}: written rows {rel_err(wa, wb):.3e}"


def test_decode_step_graph_replay_follows_refreshed_inputs():
    """Capture once, then change the step's inputs in place (what InputStager.refresh does between steps):
    the replay must equal a fresh eager step on the new inputs, bit for bit (row f2)."""
    T, page = 8, 16

    def second_step_inputs(bt, lens, first_tokens):
        lens2 = (lens + 1).clamp(max=bt.shape[1] * page)
        pos2 = lens2.long() - 1
        slots2 = bt[torch.arange(T, device=DEV), (pos2 // page)].long() * page + pos2 % page
        return first_tokens.clone(), pos2, slots2, lens2

    results = []
    for use_graph in (True, False):
        cfg, shard, kv, bt, lens, pos, slots, tokens = _small_setup(T, seed=9)
        ops.set_launch_mode(pdl=use_graph, const_weights=use_graph)
        try:
            step = DecodeStep(shard, kv, max_batch=T, want_logits=True)
            step.bind(tokens, pos, slots, bt, lens)
            if use_graph:
                step.capture()
            step.run()
            torch.cuda.synchronize()
            first = step.logits[:T].clone()
            tok2, pos2, slots2, lens2 = second_step_inputs(bt, lens, step.next_tokens[:T])
            tokens.copy_(tok2); pos.copy_(pos2); slots.copy_(slots2); lens.copy_(lens2)   # in place: same addresses
            step.run()
            torch.cuda.synchronize()
            step.check()   # no bounded in-kernel wait expired (atoma_llama_step_status on the capture and current streams)
            results.append((first, step.logits[:T].clone(), step.next_tokens[:T].clone(), [c.clone() for c in kv]))
        finally:
            ops.set_launch_mode()
    (f_g, s_g, t_g, kv_g), (f_e, s_e, t_e, kv_e) = results
    assert torch.equal(f_g, f_e) and torch.equal(s_g, s_e) and torch.equal(t_g, t_e)
    assert not torch.equal(f_g, s_g), "the second step must see the refreshed inputs"
    for a, b in zip(kv_g, kv_e):
        assert torch.equal(a, b)


def test_decode_step_pdl_modes_are_bit_identical():
    """Programmatic dependent launch, weight prefetch and early KV streaming change only when things run."""
    T = 6
    outs = []
    for mode in (dict(), dict(pdl=True), dict(pdl=True, const_weights=True)):
        cfg, shard, kv, bt, lens, pos, slots, tokens = _small_setup(T, seed=13)
        ops.set_launch_mode(**mode)
        try:
            step = DecodeStep(shard, kv, max_batch=T, want_logits=True)
            step.bind(tokens, pos, slots, bt, lens)
            step.capture()
            step.run()
            torch.cuda.synchronize()
            outs.append((step.logits[:T].clone(), [c.clone() for c in kv]))
        finally:
            ops.set_launch_mode()
    for lg, caches in outs[1:]:
        assert torch.equal(lg, outs[0][0])
        for a, b in zip(caches, outs[0][1]):
            assert torch.equal(a, b)


def test_decode_step_norm_fusion_on_and_off_are_bit_identical():
    """atoma_debug_step_fusion: the add+RMSNorm in the tail of o_proj / down_proj (5 launches per layer) against the
    separate kernels (7 per layer)."""
    from atoma_infer_b200 import _lib
    L = _lib.lib()
    T = 9
    outs = []
    for fused in (1, 0):
        cfg, shard, kv, bt, lens, pos, slots, tokens = _small_setup(T, seed=17)
        L.atoma_debug_step_fusion(fused)
        try:
            step = DecodeStep(shard, kv, max_batch=T, want_logits=True, want_hidden=True)
            step.bind(tokens, pos, slots, bt, lens)
            n = step.launches_per_step()
            assert n == (2 + 5 * cfg.num_hidden_layers + 2 if fused else 1 + 7 * cfg.num_hidden_layers + 3)
            step.launch()
            torch.cuda.synchronize()
            outs.append((step.logits[:T].clone(), step.hidden_out[:T].clone(), step.next_tokens[:T].clone(), [c.clone() for c in kv]))
        finally:
            L.atoma_debug_step_fusion(0)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    for a, b in zip(outs[0][3], outs[1][3]):
        assert torch.equal(a, b)


def test_decode_step_argument_errors():
    T = 4
    cfg, shard, kv, bt, lens, pos, slots, tokens = _small_setup(T)
    with pytest.raises(ops.AtomaError, match="1..128 sequences"):
        DecodeStep(shard, kv, max_batch=129)
    with pytest.raises(ops.AtomaError, match="kv caches"):
        DecodeStep(shard, kv[:1], max_batch=T)
    step = DecodeStep(shard, kv, max_batch=T)
    with pytest.raises(ops.AtomaError, match="bind"):
        step.launch()
    with pytest.raises(ops.AtomaError, match="block_tables must be"):
        step.bind(tokens, pos, slots, bt.long(), lens)
    with pytest.raises(ops.AtomaError, match="does not fit max_batch"):
        step.bind(tokens.repeat(2), pos.repeat(2), slots.repeat(2), bt.repeat(2, 1), lens.repeat(2))
