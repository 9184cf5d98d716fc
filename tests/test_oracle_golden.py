"""Pins the CPU oracle against every golden vector the reference's own tests hold for this path
(SURVEY.md §8c): the non-causal table of csrc/tests/flash_attn_tests.rs:54-70 (asserted there through
the dense, varlen and kv-cache APIs) and the causal table of models/src/flash_attention.rs:688-704,
plus the byte-exact expectations of csrc/tests/cache_manager_tests.rs."""
import numpy as np
import pytest

from helpers import GOLDEN, golden_inputs_np, round4
from oracle import attention_oracle as orc


def _dense_from_hts(x):  # [heads, tokens, dim] -> [1, tokens, heads, dim]
    return np.ascontiguousarray(x.transpose(1, 0, 2))[None]


@pytest.mark.parametrize("precision", ["fp64", "reference"])
def test_acausal_table_dense(precision):
    q, k, v = golden_inputs_np()
    o, _ = orc.attention_dense(_dense_from_hts(q), _dense_from_hts(k), _dense_from_hts(v), 0.5, False,
                               is_bf16=False, precision=precision)
    got = round4(o[0].transpose(1, 0, 2))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["acausal"]["table"], dtype=np.float32))


@pytest.mark.parametrize("precision", ["fp64", "reference"])
def test_causal_table_varlen(precision):
    """models/src/flash_attention.rs:632-705: forward() -> flash_attn_varlen(causal) with
    cu_seqlens [0, 2] and q/k/v transposed views (row stride 8, head stride 16)."""
    q, k, v = golden_inputs_np()  # memory layout [head][token][dim]
    H, T, D = q.shape
    flat = lambda x: orc.from_f32(x, False).ravel()
    a = orc.MhaArgs(q=flat(q), k=flat(k), v=flat(v), b=1, h=H, h_k=H, d=D, seqlen_q=32, seqlen_k=32,
                    q_batch_stride=0, k_batch_stride=0, v_batch_stride=0, o_batch_stride=0,
                    q_row_stride=D, k_row_stride=D, v_row_stride=D, o_row_stride=H * D,
                    q_head_stride=T * D, k_head_stride=T * D, v_head_stride=T * D, o_head_stride=D,
                    softmax_scale=0.5, is_bf16=False, is_causal=True,
                    cu_seqlens_q=np.array([0, 2], np.int32), cu_seqlens_k=np.array([0, 2], np.int32),
                    o_size=T * H * D)
    o, _ = orc.run_mha(a, precision)
    got = round4(orc.to_f32(o, False).reshape(T, H, D).transpose(1, 0, 2))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["causal"]["table"], dtype=np.float32))


def test_acausal_table_kv_cache_call_shape():
    """csrc/tests/flash_attn_tests.rs:194-236: kv-cache API, seqlens_k = [2] (non-cumulative)."""
    q, k, v = golden_inputs_np()
    H, T, D = q.shape
    flat = lambda x: orc.from_f32(x, False).ravel()
    a = orc.MhaArgs(q=flat(q), k=flat(k), v=flat(v), b=1, h=H, h_k=H, d=D, seqlen_q=T, seqlen_k=T,
                    q_batch_stride=H * T * D, k_batch_stride=H * T * D, v_batch_stride=H * T * D,
                    o_batch_stride=T * H * D, q_row_stride=D, k_row_stride=D, v_row_stride=D,
                    o_row_stride=H * D, q_head_stride=T * D, k_head_stride=T * D, v_head_stride=T * D,
                    o_head_stride=D, softmax_scale=0.5, is_bf16=False, is_causal=False,
                    cu_seqlens_k=np.array([2], np.int32), is_seqlens_k_cumulative=False, o_size=T * H * D)
    o, _ = orc.run_mha(a)
    got = round4(orc.to_f32(o, False).reshape(T, H, D).transpose(1, 0, 2))
    np.testing.assert_array_equal(got, np.asarray(GOLDEN["acausal"]["table"], dtype=np.float32))


def test_paged_equals_unpaged():
    """csrc/tests/flash_attn_tests.rs:140-192: 2 sequences x 32 tokens, block_size 16,
    block_table = arange(4).reshape(2, 2): the paged result equals the varlen one.  (The reference
    test builds only 32 tokens of q/k/v for cu_seqlens [0, 32, 64] and so reads past its tensors;
    the statement is made here with the 64 tokens the sequence lengths call for.)"""
    a16 = np.arange(1024, dtype=np.float32).astype(np.float16)
    mk = lambda c: orc.from_f32((a16 * np.float16(1 / c)).astype(np.float16).astype(np.float32), False)
    q, k, v = mk(300.0), mk(400.0), mk(500.0)
    cu = np.array([0, 32, 64], np.int32)
    common = dict(q=q, b=2, h=2, h_k=2, d=8, seqlen_q=32, seqlen_k=32, q_batch_stride=0, o_batch_stride=0,
                  q_row_stride=16, o_row_stride=16, q_head_stride=8, o_head_stride=8, k_head_stride=8,
                  v_head_stride=8, k_row_stride=16, v_row_stride=16, softmax_scale=0.5, is_bf16=False,
                  is_causal=False, cu_seqlens_q=cu, cu_seqlens_k=cu, o_size=1024)
    unpaged, _ = orc.run_mha(orc.MhaArgs(k=k, v=v, k_batch_stride=0, v_batch_stride=0, **common))
    paged, _ = orc.run_mha(orc.MhaArgs(k=k, v=v, k_batch_stride=16 * 16, v_batch_stride=16 * 16,
                                       block_table=np.arange(4, dtype=np.int32), block_table_batch_stride=2,
                                       page_block_size=16, **common))
    np.testing.assert_array_equal(paged, unpaged)
    assert np.abs(orc.to_f32(unpaged, False)).max() > 0


def test_empty_and_ragged_sequences():
    rng = np.random.default_rng(0)
    B, H, HK, D, page, MB = 3, 4, 2, 16, 16, 3
    q = rng.standard_normal((B, 1, H, D)).astype(np.float32)
    kc = rng.standard_normal((B * MB, page, HK, D)).astype(np.float32)
    vc = rng.standard_normal((B * MB, page, HK, D)).astype(np.float32)
    bt = rng.permutation(B * MB).reshape(B, MB)
    lens = np.array([0, 17, 48])
    o, lse = orc.attention_paged_decode(q, kc, vc, bt, lens, 0.25, is_bf16=True)
    assert np.all(o[0] == 0) and np.all(np.isinf(lse[0]))  # flash_fwd_kernel.h:543-582
    # brute force for sequence 1, head 3 (kv head 1)
    qq = orc.round_to_io(q, True)[1, 0, 3].astype(np.float64)
    toks = [(bt[1, t // page], t % page) for t in range(17)]
    K = np.stack([orc.round_to_io(kc, True)[b_, o_, 1] for b_, o_ in toks]).astype(np.float64)
    V = np.stack([orc.round_to_io(vc, True)[b_, o_, 1] for b_, o_ in toks]).astype(np.float64)
    s = K @ qq * 0.25
    p = np.exp(s - s.max())
    ref = (p / p.sum()) @ V
    np.testing.assert_allclose(o[1, 0, 3], orc.round_to_io(ref.astype(np.float32), True), rtol=0, atol=1e-6)


# ---- cache ops (csrc/tests/cache_manager_tests.rs) ----
def test_reshape_and_cache_flash_identity_mapping():
    """:552-684: 10 tokens, block_size 8, slot_mapping = arange(10): token t lands in block t/8,
    offset t%8; everything else untouched."""
    rng = np.random.default_rng(1)
    T, H, D, NB, BS = 10, 2, 8, 4, 8
    key = rng.integers(0, 65535, T * H * D, dtype=np.uint16)
    val = rng.integers(0, 65535, T * H * D, dtype=np.uint16)
    kc = np.zeros(NB * BS * H * D, np.uint16)
    vc = np.zeros_like(kc)
    orc.reshape_and_cache_flash(key, val, kc, vc, np.arange(T), BS * H * D, T, H, D, BS, H * D, H * D)
    np.testing.assert_array_equal(kc[: T * H * D], key)
    np.testing.assert_array_equal(vc[: T * H * D], val)
    assert not kc[T * H * D:].any() and not vc[T * H * D:].any()


def test_reshape_and_cache_flash_skips_negative_slots():
    T, H, D, BS = 4, 1, 8, 4
    key = np.arange(T * H * D, dtype=np.uint16) + 1
    kc = np.zeros(2 * BS * H * D, np.uint16)
    vc = np.zeros_like(kc)
    orc.reshape_and_cache_flash(key, key, kc, vc, np.array([5, -1, 0, -1]), BS * H * D, T, H, D, BS, H * D, H * D)
    np.testing.assert_array_equal(kc[5 * 8:6 * 8], key[0:8])
    np.testing.assert_array_equal(kc[0:8], key[16:24])
    assert kc.astype(bool).sum() == 16


def test_copy_blocks_mapping():
    """:241-454: mapping [[0,2],[1,3]], 2 layers, NUM_BLOCKS=4, BLOCK_SIZE=64, H=2, D=8."""
    rng = np.random.default_rng(2)
    n = 64 * 2 * 8
    kcs = [rng.integers(0, 65535, 4 * n, dtype=np.uint16) for _ in range(2)]
    vcs = [rng.integers(0, 65535, 4 * n, dtype=np.uint16) for _ in range(2)]
    k0 = [k.copy() for k in kcs]
    v0 = [v.copy() for v in vcs]
    orc.copy_blocks(kcs, vcs, np.array([[0, 2], [1, 3]]), n)
    for l in range(2):
        np.testing.assert_array_equal(kcs[l][2 * n:3 * n], k0[l][0:n])
        np.testing.assert_array_equal(kcs[l][3 * n:4 * n], k0[l][n:2 * n])
        np.testing.assert_array_equal(kcs[l][:2 * n], k0[l][:2 * n])
        np.testing.assert_array_equal(vcs[l][2 * n:], v0[l][:2 * n])


def test_swap_blocks_mapping():
    """:64-186: mapped blocks equal, unmapped blocks untouched."""
    rng = np.random.default_rng(3)
    n = 32
    src = rng.integers(0, 65535, 8 * n, dtype=np.uint16)
    dst = rng.integers(0, 65535, 8 * n, dtype=np.uint16)
    d0 = dst.copy()
    orc.swap_blocks(src, dst, {0: 3, 5: 1}, n)
    np.testing.assert_array_equal(dst[3 * n:4 * n], src[0:n])
    np.testing.assert_array_equal(dst[n:2 * n], src[5 * n:6 * n])
    keep = np.ones(8 * n, bool)
    keep[3 * n:4 * n] = False
    keep[n:2 * n] = False
    np.testing.assert_array_equal(dst[keep], d0[keep])


def test_bf16_round_trip_helpers():
    import torch
    x = torch.randn(4096, dtype=torch.float32)
    want = x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(orc.f32_to_bf16_bits(x.numpy()), want)
    np.testing.assert_array_equal(orc.bf16_bits_to_f32(want), x.to(torch.bfloat16).float().numpy())
