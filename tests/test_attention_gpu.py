"""GPU parity tests for the attention kernels, called through the C ABI (host layer -> run_mha).

Layout follows the reference's own tests (csrc/tests/flash_attn_tests.rs, models/src/
flash_attention.rs:472-719): golden tables, paged == unpaged self-consistency; then seeded parity
against the CPU oracle at sizes it finishes in seconds, and size-independent properties at
BASELINE.json's full configuration.

Tolerance (the north star's "<= 1e-3 rel bf16"), helpers.assert_close_16: against the UN-ROUNDED
fp64 oracle, norm-wise relative error <= 1e-3 + the 16-bit output format's own quantisation floor,
and element-wise |out - exact| <= 1e-3 |exact| + 0.5 ulp_16 + the P-rounding allowance (the
reference rounds P to the io type before P.V, flash_fwd_kernel.h:949-954; so do we).  Indexing
properties (one-hot values, permuted block tables, paged == unpaged) are bit-exact.
"""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from atoma_infer_b200.flash_attention import (FlashAttention, FlashAttentionDecodingMetadata,
                                              FlashAttentionMetadata, FlashAttentionPrefillMetadata)
from helpers import (GOLDEN, arange_inputs_np, assert_close_16, bits, golden_inputs_np, randn_t,
                     random_block_table, rel_err, round4, t2f32)
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def u32(x):
    return torch.tensor(x, dtype=torch.int32, device=DEV)


# ------------------------------------------------------------------------------------------------
# golden tables (reference csrc/tests/flash_attn_tests.rs)
# ------------------------------------------------------------------------------------------------
def _golden_t():
    q, k, v = golden_inputs_np()  # [heads, tokens, dim]
    f = lambda x: torch.from_numpy(x).to(torch.float16).to(DEV)
    return f(q), f(k), f(v)


def test_flash_attn_acausal_golden():
    """flash_attn_tests.rs:31-93"""
    q, k, v = _golden_t()
    ys = ops.flash_attn(q[None].transpose(1, 2), k[None].transpose(1, 2), v[None].transpose(1, 2), 0.5, False)
    got = round4(t2f32(ys.transpose(1, 2)[0]))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["acausal"]["table"], np.float32))
    assert ops.last_mha_path() == "tile"


def test_flash_attn_varlen_golden():
    """flash_attn_tests.rs:95-138: cu_seqlens [0, 2], max_seqlen 32, transposed (strided) inputs"""
    q, k, v = _golden_t()
    cu = u32([0, 2])
    ys = ops.flash_attn_varlen(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), cu, cu, 32, 32, 0.5, False)
    got = round4(t2f32(ys.transpose(0, 1)))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["acausal"]["table"], np.float32))


def test_flash_attn_kv_cache_golden():
    """flash_attn_tests.rs:194-236: kv-cache API without block table, seqlens_k = [2]"""
    q, k, v = _golden_t()
    ys = ops.flash_attn_kv_cache_full(q[None].transpose(1, 2), k[None].transpose(1, 2), v[None].transpose(1, 2),
                                      None, 0.5, None, u32([2]), False)
    got = round4(t2f32(ys.transpose(1, 2)[0]))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["acausal"]["table"], np.float32))


def test_forward_with_varlen_causal_golden():
    """models/src/flash_attention.rs:632-705: operator forward, causal prefill of 2 tokens, 3 heads"""
    q, k, v = _golden_t()
    kv_cache = torch.zeros((2, 128, 16, 3, 8), dtype=torch.float16, device=DEV)
    fa = FlashAttention(3, 3, 8, 0.5, kv_cache_dtype=torch.float16, device=DEV, check_head_dim=False)
    md = FlashAttentionMetadata(
        context_lengths=None, slot_mapping=torch.arange(2, dtype=torch.int64, device=DEV),
        prefill_metadata=FlashAttentionPrefillMetadata(None, None, 32, None, u32([0, 2]),
                                                       torch.tensor([2], dtype=torch.int64, device=DEV)),
        decoding_metadata=None, num_prefill_tokens=2, num_decoding_tokens=0)
    out = fa.forward(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), kv_cache, md)
    got = round4(t2f32(out.reshape(2, 3, 8).transpose(0, 1)))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["causal"]["table"], np.float32))
    # The cache write happened first (flash_attention.rs:360-361).  reshape_and_cache_flash only
    # knows the TOKEN stride (cache_manager.cu:157-159 reads key[token * key_stride + i]), so with
    # these transposed views (token stride 8, head stride 16) the reference scatters the 24 elements
    # that follow each token's first head in memory — not the token's three heads.  Same ABI here.
    flat_k, flat_v = k.reshape(-1), v.reshape(-1)
    for t in range(2):
        assert torch.equal(kv_cache[0, 0, t].reshape(-1), flat_k[t * 8: t * 8 + 24])
        assert torch.equal(kv_cache[1, 0, t].reshape(-1), flat_v[t * 8: t * 8 + 24])


# ------------------------------------------------------------------------------------------------
# paged == unpaged (the reference's only test of block-table indexing)
# ------------------------------------------------------------------------------------------------
def test_varlen_with_block_table_equals_unpaged():
    """flash_attn_tests.rs:140-192 (with the 64 tokens its sequence lengths call for)"""
    qn, kn, vn = arange_inputs_np(1024, (64, 2, 8))
    f = lambda x: torch.from_numpy(x / 10).to(torch.float16).to(DEV)
    q, k, v = f(qn), f(kn), f(vn)
    cu = u32([0, 32, 64])
    bt = torch.arange(4, dtype=torch.int32, device=DEV).reshape(2, 2)
    ys = ops.flash_attn_varlen_with_block_table(q, k.reshape(4, 16, 2, 8), v.reshape(4, 16, 2, 8), None, cu, cu,
                                                32, 32, 0.5, None, None, bt)
    want = ops.flash_attn_varlen(q, k, v, cu, cu, 32, 32, 0.5, False)
    assert ys.shape == (64, 2, 8)
    assert torch.equal(ys, want)
    assert float(want.abs().max()) > 0


def test_kv_cache_with_block_table_equals_paged_varlen():
    """flash_attn_tests.rs:238-303: 32 decode queries, seqlens_k = 1 each, block_table
    arange(64).reshape(32, 2); kv-cache API == paged-varlen API."""
    qn, kn, vn = arange_inputs_np(512, (32, 2, 8))
    f = lambda x: torch.from_numpy(x).to(torch.float16).to(DEV)
    q = f(qn)
    kc = torch.zeros((64, 16, 2, 8), dtype=torch.float16, device=DEV)
    vc = torch.zeros_like(kc)
    kc.view(-1, 2, 8)[:32] = f(kn)
    vc.view(-1, 2, 8)[:32] = f(vn)
    bt = torch.arange(64, dtype=torch.int32, device=DEV).reshape(32, 2)
    ys = ops.flash_attn_kv_cache_full(q.unsqueeze(1), kc, vc, None, 0.5, bt, u32([1] * 32), False).squeeze(1)
    cu = torch.arange(33, dtype=torch.int32, device=DEV)
    want = ops.flash_attn_varlen_with_block_table(q, kc, vc, None, cu, cu, 32, 32, 0.5, None, None, bt)
    assert torch.equal(ys, want)
    # one key per query: the output is exactly that key's value row (block 2*i, offset 0)
    assert torch.equal(ys, vc[bt[:, 0].long(), 0])


# ------------------------------------------------------------------------------------------------
# decode kernel vs CPU oracle
# ------------------------------------------------------------------------------------------------
def _decode_case(B, H, HK, D, page, MB, lens, dtype, seed, scale=None, extra_blocks=3, q_scale=1.0):
    NB = B * MB + extra_blocks
    kv = randn_t((2, NB, page, HK, D), seed, dtype, DEV)
    q = randn_t((B, 1, H, D), seed + 1, dtype, DEV, q_scale)
    bt = random_block_table(B, MB, NB, seed + 2, DEV)
    sl = u32(lens)
    scale = scale or 1.0 / np.sqrt(D)
    out = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, scale, bt, sl, True)
    path = ops.last_mha_path()
    is_bf16 = dtype == torch.bfloat16
    want, _ = orc.attention_paged_decode(t2f32(q), t2f32(kv[0]), t2f32(kv[1]), bt.cpu().numpy(),
                                         np.asarray(lens), scale, is_bf16, exact=True)
    return out, want, path, is_bf16


@pytest.mark.parametrize("B,H,HK,D,page,MB,dtype", [
    (4, 32, 8, 128, 16, 12, torch.bfloat16),   # Llama-3.1-8B head geometry
    (3, 32, 8, 64, 16, 20, torch.bfloat16),    # Llama-3.2-1B head geometry
    (5, 8, 1, 128, 16, 24, torch.bfloat16),    # Llama-3.1-70B TP=8 per-rank geometry
    (2, 16, 4, 128, 32, 6, torch.float16),
    (2, 4, 2, 64, 16, 9, torch.float16),
    (3, 16, 16, 128, 16, 5, torch.bfloat16),   # MHA (G = 1), two head groups
    (2, 24, 12, 64, 48, 4, torch.bfloat16),    # 12 kv heads -> tiles of 4; page not a power of two
    (2, 32, 2, 128, 64, 3, torch.bfloat16),    # G = 16
    # Llama-3.1-8B tensor-parallel shards (Hkv = 8 / tp) at long context: every sequence is split over
    # many CTAs, the token-slice warps of a kv head are merged inside the CTA first
    (8, 16, 4, 128, 16, 96, torch.bfloat16),   # tp2
    (8, 8, 2, 128, 16, 128, torch.bfloat16),   # tp4
    (8, 4, 1, 128, 16, 128, torch.bfloat16),   # tp8
    (6, 8, 1, 64, 16, 64, torch.float16),      # one kv head, D = 64 (4-D tensor map path)
])
def test_decode_paged_vs_oracle(B, H, HK, D, page, MB, dtype):
    rng = np.random.default_rng(B * 1000 + H)
    lens = rng.integers(1, MB * page + 1, size=B).tolist()
    lens[0] = MB * page  # one full-length sequence
    out, want, path, is_bf16 = _decode_case(B, H, HK, D, page, MB, lens, dtype, seed=7)
    assert path == "decode_paged"
    assert_close_16(t2f32(out), want, is_bf16, f"decode B={B} H={H} HK={HK} D={D} page={page}")


@pytest.mark.parametrize("HK,order", [(8, "separate"), (8, "v_first"), (1, "separate"), (4, "no_merge")])
def test_decode_k_and_v_in_separate_allocations(HK, order):
    """The kernel uses ONE tensor map (one op per page for K and V) when K and V are the two halves of
    one allocation, as in the reference's [2, NB, bs, Hkv, D] cache; separate allocations, V below
    K, or the diagnostic switch take the two-map path.  All must give the same bits."""
    from atoma_infer_b200 import _lib
    B, G, D, page, MB = 4, 4, 128, 16, 40
    H = HK * G
    NB = B * MB + 3
    rng = np.random.default_rng(HK)
    lens = rng.integers(1, MB * page + 1, size=B).tolist()
    lens[0] = MB * page
    kv = randn_t((2, NB, page, HK, D), 61, torch.bfloat16, DEV)
    q = randn_t((B, 1, H, D), 62, torch.bfloat16, DEV)
    bt = random_block_table(B, MB, NB, 63, DEV)
    base = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, D ** -0.5, bt, u32(lens), True)
    if order == "separate":
        k, v = kv[0].clone(), kv[1].clone()
    elif order == "v_first":
        both = torch.empty_like(kv)
        both[1].copy_(kv[0]); both[0].copy_(kv[1])
        k, v = both[1], both[0]            # V lies below K in memory
    else:
        k, v = kv[0], kv[1]
        _lib.lib().atoma_debug_decode_merge(0)
    try:
        out = ops.flash_attn_kv_cache_full(q, k, v, None, D ** -0.5, bt, u32(lens), True)
    finally:
        _lib.lib().atoma_debug_decode_merge(1)
    assert ops.last_mha_path() == "decode_paged"
    assert torch.equal(out, base)


def test_decode_ragged_empty_and_single_token():
    """Edge cases of SURVEY appendix B: len 0 -> zeros, len 1, ragged tails in the first / last page,
    a length one short of and exactly at a stage boundary."""
    lens = [0, 1, 15, 16, 17, 127, 128, 129, 160, 0, 5]
    out, want, path, is_bf16 = _decode_case(len(lens), 32, 8, 128, 16, 10, lens, torch.bfloat16, seed=11)
    assert path == "decode_paged"
    assert torch.count_nonzero(out[0]) == 0 and torch.count_nonzero(out[9]) == 0
    assert_close_16(t2f32(out), want, True, "decode ragged")


def test_decode_ignores_garbage_beyond_sequence_and_in_padding():
    """Cache rows past seqlens_k (and padded block-table entries) may hold anything, NaN included:
    the reference masks them (flash_fwd_kernel.h:901-921); so must we."""
    B, H, HK, D, page, MB = 3, 32, 8, 128, 16, 8
    lens = [37, 100, 128]
    NB = B * MB + 2
    kv = randn_t((2, NB, page, HK, D), 3, torch.bfloat16, DEV)
    q = randn_t((B, 1, H, D), 4, torch.bfloat16, DEV)
    bt = random_block_table(B, MB, NB, 5, DEV)
    clean = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.09, bt, u32(lens), True).clone()
    kv2 = kv.clone()
    for b, n in enumerate(lens):
        for t in range(n, MB * page):
            kv2[:, bt[b, t // page].item(), t % page] = float("nan")
    bt2 = bt.clone()
    for b, n in enumerate(lens):
        used = (n + page - 1) // page
        bt2[b, used:] = 0  # the reference pads block tables with 0 (worker.rs:410-412)
    dirty = ops.flash_attn_kv_cache_full(q, kv2[0], kv2[1], None, 0.09, bt2, u32(lens), True)
    assert torch.isfinite(dirty.float()).all()
    assert torch.equal(clean, dirty)


def test_decode_block_table_indexing_is_exact():
    """One-hot values: V[token t, head g, :] = e_(t mod D) and a single dominant key per sequence
    makes the output an exact copy of one value row — any indexing slip changes it."""
    B, H, HK, D, page, MB = 6, 32, 8, 128, 16, 16
    NB = B * MB
    bt = random_block_table(B, MB, NB, 21, DEV)
    lens = [200, 256, 33, 1, 129, 255]
    kv = torch.zeros((2, NB, page, HK, D), dtype=torch.bfloat16, device=DEV)
    q = torch.zeros((B, 1, H, D), dtype=torch.bfloat16, device=DEV)
    hot = [150, 255, 32, 0, 128, 17]
    for b in range(B):
        q[b, 0, :, 0] = 16.0
        t = hot[b]
        blk = bt[b, t // page].item()
        kv[0, blk, t % page, :, 0] = 16.0           # score 256*scale vs 0 elsewhere
        kv[1, blk, t % page, :, (t * 7 + b) % D] = 1.0
    out = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 1.0, bt, u32(lens), True)
    for b in range(B):
        want = torch.zeros(D, device=DEV)
        want[(hot[b] * 7 + b) % D] = 1.0
        assert torch.equal(out[b, 0].float(), want.expand(H, D)), b


def test_decode_is_invariant_to_physical_block_order():
    """models/src/llama.rs:918-1152 runs the model with a random physical block order; here: the
    same logical cache laid out under two different block tables gives bit-identical outputs."""
    B, H, HK, D, page, MB = 4, 32, 8, 128, 16, 20
    NB = B * MB
    lens = [320, 300, 17, 161]
    logical = randn_t((2, B, MB * page, HK, D), 31, torch.bfloat16, DEV)
    q = randn_t((B, 1, H, D), 32, torch.bfloat16, DEV)
    outs = []
    for seed in (1, 2):
        bt = random_block_table(B, MB, NB, seed, DEV)
        kv = torch.empty((2, NB, page, HK, D), dtype=torch.bfloat16, device=DEV)
        idx = bt.long().reshape(-1)
        kv[:, idx] = logical.reshape(2, B * MB, page, HK, D)
        outs.append(ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.088, bt, u32(lens), True))
    assert torch.equal(outs[0], outs[1])


def test_decode_legacy_symbol_matches_stream_variant():
    """`run_mha` (stream 0, the reference's symbol) and `run_mha_stream` give identical bits."""
    B, H, HK, D, page, MB = 2, 32, 8, 128, 16, 8
    kv = randn_t((2, B * MB, page, HK, D), 41, torch.bfloat16, DEV)
    q = randn_t((B, 1, H, D), 42, torch.bfloat16, DEV)
    bt = random_block_table(B, MB, B * MB, 43, DEV)
    a = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.088, bt, u32([100, 128]), True)
    b = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.088, bt, u32([100, 128]), True, legacy_stream=True)
    assert torch.equal(a, b)


def test_decode_lse_output():
    B, H, HK, D, page, MB = 2, 8, 2, 64, 16, 4
    lens = [50, 64]
    kv = randn_t((2, B * MB, page, HK, D), 51, torch.bfloat16, DEV)
    q = randn_t((B, 1, H, D), 52, torch.bfloat16, DEV)
    bt = random_block_table(B, MB, B * MB, 53, DEV)
    out, lse = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.125, bt, u32(lens), True, with_lse=True)
    _, want_lse = orc.attention_paged_decode(t2f32(q), t2f32(kv[0]), t2f32(kv[1]), bt.cpu().numpy(),
                                             np.asarray(lens), 0.125, True)
    np.testing.assert_allclose(lse.cpu().numpy().reshape(B, H, 1), want_lse, rtol=2e-4, atol=2e-4)


# ------------------------------------------------------------------------------------------------
# tile kernel (prefill / dense / paged multi-token) vs CPU oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,HK,D,dtype,causal", [
    (8, 2, 128, torch.bfloat16, True),
    (8, 8, 64, torch.float16, True),
    (4, 2, 64, torch.bfloat16, False),
    (4, 4, 32, torch.float16, True),
    (2, 1, 256, torch.bfloat16, True),
    (6, 3, 40, torch.float16, True),    # head dim that is not a template size
    (3, 3, 8, torch.float16, False),
])
def test_varlen_prefill_vs_oracle(H, HK, D, dtype, causal):
    lens = [1, 77, 130, 64, 200]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T = int(cu[-1])
    q = randn_t((T, H, D), 1, dtype, DEV)
    k = randn_t((T, HK, D), 2, dtype, DEV)
    v = randn_t((T, HK, D), 3, dtype, DEV)
    scale = 1.0 / np.sqrt(D)
    cu_t = torch.from_numpy(cu).to(DEV)
    out = ops.flash_attn_varlen(q, k, v, cu_t, cu_t, max(lens), max(lens), scale, causal)
    assert ops.last_mha_path() == ("prefill_tc" if D in (64, 128) else "tile")
    is_bf16 = dtype == torch.bfloat16
    flat = lambda t: bits(t).ravel()
    a = orc.MhaArgs(q=flat(q), k=flat(k), v=flat(v), b=len(lens), h=H, h_k=HK, d=D, seqlen_q=max(lens),
                    seqlen_k=max(lens), q_batch_stride=0, k_batch_stride=0, v_batch_stride=0, o_batch_stride=0,
                    q_row_stride=H * D, k_row_stride=HK * D, v_row_stride=HK * D, o_row_stride=H * D,
                    q_head_stride=D, k_head_stride=D, v_head_stride=D, o_head_stride=D, softmax_scale=scale,
                    is_bf16=is_bf16, is_causal=causal, cu_seqlens_q=cu, cu_seqlens_k=cu, o_size=T * H * D)
    _, _, want = orc.run_mha(a, return_exact=True)
    assert_close_16(t2f32(out), want.reshape(T, H, D), is_bf16, f"varlen D={D}")


def test_dense_causal_bottom_right_alignment_and_alibi():
    """mask.h:171-172 (seqlen_q < seqlen_k and seqlen_q > seqlen_k) and ALiBi (mask.h:179-186)."""
    for (sq, sk) in [(5, 70), (70, 5), (64, 64)]:
        B, H, HK, D = 2, 4, 2, 64
        q = randn_t((B, sq, H, D), 1, torch.bfloat16, DEV)
        k = randn_t((B, sk, HK, D), 2, torch.bfloat16, DEV)
        v = randn_t((B, sk, HK, D), 3, torch.bfloat16, DEV)
        out = ops.flash_attn(q, k, v, 0.125, True)
        want, _ = orc.attention_dense(t2f32(q), t2f32(k), t2f32(v), 0.125, True, True, exact=True)
        assert_close_16(t2f32(out), want, True, f"dense causal {sq}x{sk}")
        slopes = torch.tensor([0.5, 0.25, 0.125, 0.0625], dtype=torch.float32, device=DEV)
        for causal in (True, False):
            out = ops.flash_attn_alibi(q, k, v, slopes, 0.125, causal)
            want, _ = orc.attention_dense(t2f32(q), t2f32(k), t2f32(v), 0.125, causal, True,
                                          alibi=slopes.cpu().numpy(), exact=True)
            assert_close_16(t2f32(out), want, True, f"alibi causal={causal} {sq}x{sk}")


def test_paged_multi_token_queries_use_the_tile_kernel():
    """kv-cache API with seqlen_q > 1 and a block table (chunked-prefill shape)."""
    B, SQ, H, HK, D, page, MB = 3, 7, 8, 2, 128, 16, 6
    lens = [90, 7, 96]
    kv = randn_t((2, B * MB, page, HK, D), 61, torch.bfloat16, DEV)
    q = randn_t((B, SQ, H, D), 62, torch.bfloat16, DEV)
    bt = random_block_table(B, MB, B * MB, 63, DEV)
    out = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, 0.088, bt, u32(lens), True)
    assert ops.last_mha_path() == "tile"
    a = orc.MhaArgs(q=bits(q).ravel(), k=bits(kv[0]).ravel(), v=bits(kv[1]).ravel(), b=B, h=H, h_k=HK, d=D,
                    seqlen_q=SQ, seqlen_k=MB * page, q_batch_stride=SQ * H * D, k_batch_stride=page * HK * D,
                    v_batch_stride=page * HK * D, o_batch_stride=SQ * H * D, q_row_stride=H * D,
                    k_row_stride=HK * D, v_row_stride=HK * D, o_row_stride=H * D, q_head_stride=D,
                    k_head_stride=D, v_head_stride=D, o_head_stride=D, softmax_scale=0.088, is_bf16=True,
                    is_causal=True, cu_seqlens_k=np.asarray(lens, np.int32), is_seqlens_k_cumulative=False,
                    block_table=bt.cpu().numpy().ravel(), block_table_batch_stride=MB, page_block_size=page,
                    o_size=B * SQ * H * D)
    _, _, want = orc.run_mha(a, return_exact=True)
    assert_close_16(t2f32(out), want.reshape(B, SQ, H, D), True, "paged multi-token")


def test_operator_forward_mixed_prefill_and_decode():
    """models/src/flash_attention.rs:548-630 (`test_forward`): 10 prefill + 5 decode tokens in one
    flattened batch.  The reference asserts only "no zero in the output"; here every row is checked
    against the oracle, and the cache against the scatter."""
    H, HK, D, page, NB = 8, 2, 64, 16, 32
    P, Dn = 10, 5
    fa = FlashAttention(H, HK, D, 0.125, kv_cache_dtype=torch.bfloat16, device=DEV)
    kv_cache = randn_t((2, NB, page, HK, D), 71, torch.bfloat16, DEV)
    q = randn_t((P + Dn, H, D), 72, torch.bfloat16, DEV)
    k = randn_t((P + Dn, HK, D), 73, torch.bfloat16, DEV)
    v = randn_t((P + Dn, HK, D), 74, torch.bfloat16, DEV)
    # prefill: two prompts of 6 and 4 tokens in blocks 0 and 1; decode: 5 sequences of lengths 20..24
    dec_lens = [20, 21, 22, 23, 24]
    bt = torch.zeros((Dn, 2), dtype=torch.int32, device=DEV)
    slots = [0 * page + i for i in range(6)] + [1 * page + i for i in range(4)]
    for i, n in enumerate(dec_lens):
        bt[i, 0], bt[i, 1] = 2 + 2 * i, 3 + 2 * i
        slots.append(int(bt[i, (n - 1) // page]) * page + (n - 1) % page)  # worker.rs:392-401
    md = FlashAttentionMetadata(
        context_lengths=None, slot_mapping=torch.tensor(slots, dtype=torch.int64, device=DEV),
        prefill_metadata=FlashAttentionPrefillMetadata(None, 6, 6, u32([0, 6]), u32([0, 6, 10]), u32([6, 4])),
        decoding_metadata=FlashAttentionDecodingMetadata(bt, 24, u32(dec_lens)),
        num_prefill_tokens=P, num_decoding_tokens=Dn)
    cache_before = kv_cache.clone()
    out = fa.forward(q, k, v, kv_cache, md).reshape(P + Dn, H, D)
    # cache scatter is exact
    want_cache = cache_before.clone()
    for t, s in enumerate(slots):
        want_cache[0, s // page, s % page] = k[t]
        want_cache[1, s // page, s % page] = v[t]
    assert torch.equal(kv_cache, want_cache)
    # prefill rows
    _, _, wp = orc.run_mha(orc.MhaArgs(
        q=bits(q[:P]).ravel(), k=bits(k[:P]).ravel(), v=bits(v[:P]).ravel(), b=2, h=H, h_k=HK, d=D, seqlen_q=6,
        seqlen_k=6, q_batch_stride=0, k_batch_stride=0, v_batch_stride=0, o_batch_stride=0, q_row_stride=H * D,
        k_row_stride=HK * D, v_row_stride=HK * D, o_row_stride=H * D, q_head_stride=D, k_head_stride=D,
        v_head_stride=D, o_head_stride=D, softmax_scale=0.125, is_bf16=True, is_causal=True,
        cu_seqlens_q=np.array([0, 6, 10], np.int32), cu_seqlens_k=np.array([0, 6, 10], np.int32), o_size=P * H * D),
        return_exact=True)
    assert_close_16(t2f32(out[:P]), wp.reshape(P, H, D), True, "forward prefill rows")
    wd, _ = orc.attention_paged_decode(t2f32(q[P:]).reshape(Dn, 1, H, D), t2f32(kv_cache[0]), t2f32(kv_cache[1]),
                                       bt.cpu().numpy(), np.asarray(dec_lens), 0.125, True, exact=True)
    assert_close_16(t2f32(out[P:]), wd.reshape(Dn, H, D), True, "forward decode rows")


# ------------------------------------------------------------------------------------------------
# BASELINE.json configuration 2 at full size: properties that do not need the CPU oracle
# ------------------------------------------------------------------------------------------------
def _config2(seed, lens=None, NB=None):
    B, H, HK, D, page, MB = 32, 32, 8, 128, 16, 256
    NB = NB or (B * MB + 64)
    g = torch.Generator(device=DEV).manual_seed(seed)
    kv = torch.randn((2, NB, page, HK, D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    q = torch.randn((B, 1, H, D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    bt = random_block_table(B, MB, NB, seed, DEV)
    lens = lens or [4096] * B
    return q, kv, bt, u32(lens), lens


def _torch_decode_reference(q, kv, bt, lens, scale):
    """fp32 torch restatement used only at full size (the numpy oracle would take minutes)."""
    B, _, H, D = q.shape
    HK, page = kv.shape[3], kv.shape[2]
    outs = []
    for b in range(B):
        n = lens[b]
        pages = bt[b, : (n + page - 1) // page].long()
        K = kv[0, pages].reshape(-1, HK, D)[:n].float()
        V = kv[1, pages].reshape(-1, HK, D)[:n].float()
        qq = q[b, 0].float().reshape(HK, H // HK, D)
        s = torch.einsum("hgd,thd->hgt", qq, K) * scale
        p = torch.softmax(s, dim=-1)
        outs.append(torch.einsum("hgt,thd->hgd", p, V).reshape(H, D))
    return torch.stack(outs).unsqueeze(1)


def test_config2_full_size_matches_fp32_and_is_deterministic():
    q, kv, bt, sl, lens = _config2(5)
    scale = 1.0 / np.sqrt(128)
    a = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, scale, bt, sl, True)
    assert ops.last_mha_path() == "decode_paged"
    b = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, scale, bt, sl, True)
    assert torch.equal(a, b), "run-to-run determinism"
    want = _torch_decode_reference(q, kv, bt, lens, scale)
    assert_close_16(t2f32(a), t2f32(want).astype(np.float64), True, "config 2 full size")


def test_config2_ragged_full_size():
    rng = np.random.default_rng(9)
    lens = rng.integers(2049, 4097, size=32).tolist()
    q, kv, bt, sl, lens = _config2(6, lens)
    scale = 1.0 / np.sqrt(128)
    a = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, scale, bt, sl, True)
    want = _torch_decode_reference(q, kv, bt, lens, scale)
    assert_close_16(t2f32(a), t2f32(want).astype(np.float64), True, "config 2 ragged")


def test_config2_value_linearity():
    """softmax(QK^T) is independent of V, so O(V1 + V2) = O(V1) + O(V2) up to output rounding."""
    q, kv, bt, sl, lens = _config2(7, [1024] * 32)
    v1 = kv[1]
    v2 = torch.roll(kv[1], 1, dims=0)
    scale = 1.0 / np.sqrt(128)
    o1 = ops.flash_attn_kv_cache_full(q, kv[0], v1, None, scale, bt, sl, True).float()
    o2 = ops.flash_attn_kv_cache_full(q, kv[0], v2, None, scale, bt, sl, True).float()
    o12 = ops.flash_attn_kv_cache_full(q, kv[0], (v1.float() + v2.float()).to(torch.bfloat16), None, scale, bt, sl, True).float()
    assert rel_err(t2f32(o12), t2f32(o1 + o2)) <= 6e-3  # three bf16 roundings of inputs/outputs


def test_peaky_softmax_range():
    """SURVEY §8d "peaky" variant: q * 4 stresses the exp2 range."""
    lens = [300, 511, 64]
    out, want, path, is_bf16 = _decode_case(3, 32, 8, 128, 16, 32, lens, torch.bfloat16, seed=13, q_scale=4.0)
    assert_close_16(t2f32(out), want, True, "peaky decode")


def test_sliding_window_and_softcap_are_refused_not_silently_ignored():
    """The reference compiles LOCAL and SOFTCAP out (static_switch.h:7-11) and computes full attention for
    them; round 1 did the same silently.  A window that would mask something, or a soft cap, now raises."""
    q = randn_t((1, 64, 4, 64), 1, torch.bfloat16, DEV)
    k = randn_t((1, 64, 2, 64), 2, torch.bfloat16, DEV)
    v = randn_t((1, 64, 2, 64), 3, torch.bfloat16, DEV)
    full = ops.flash_attn(q, k, v, 0.125, causal=True)
    # a left window longer than the sequence is folded to "none" by the wrapper (csrc/src/lib.rs:250-255): plain causal
    assert torch.equal(ops.flash_attn_windowed(q, k, v, 0.125, 1000, 0), full)
    with pytest.raises(ops.AtomaError, match="sliding-window"):
        ops.flash_attn_windowed(q, k, v, 0.125, 8, 0)
    with pytest.raises(ops.AtomaError, match="sliding-window"):
        ops.flash_attn_windowed(q, k, v, 0.125, -1, 4)
    plan = ops.flash_attn_plan(q, k, v, 0.125, None, 0, softcap=30.0)
    with pytest.raises(ops.AtomaError, match="soft-capped"):
        ops._execute(plan, q)
