"""GPU parity of the tcgen05 prefill attention kernel (prefill_attn_kernel; run_mha without a block
table, head dim 64 / 128) through the C ABI: against the fp64 oracle at sizes it finishes in
seconds, against the reference's own kernels, and at BASELINE.json config 3's prompt length against
an fp32 torch restatement.  Tolerance: helpers.assert_close_16 (see tests/test_attention_gpu.py)."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from helpers import RefLib, assert_close_16, assert_same_up_to_rounding, bits, randn_t, t2f32
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def u32(x):
    return torch.tensor(x, dtype=torch.int32, device=DEV)


def _varlen(lens, H, HK, D, dtype, causal, seed=0, q_scale=1.0, kq=None):
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T = int(cu[-1])
    q = randn_t((T, H, D), seed + 1, dtype, DEV, q_scale)
    k = randn_t((T, HK, D), seed + 2, dtype, DEV) if kq is None else kq
    v = randn_t((T, HK, D), seed + 3, dtype, DEV)
    scale = 1.0 / np.sqrt(D)
    cu_t = torch.from_numpy(cu).to(DEV)
    out = ops.flash_attn_varlen(q, k, v, cu_t, cu_t, max(lens), max(lens), scale, causal)
    path = ops.last_mha_path()
    is_bf16 = dtype == torch.bfloat16
    a = orc.MhaArgs(q=bits(q).ravel(), k=bits(k).ravel(), v=bits(v).ravel(), b=len(lens), h=H, h_k=HK, d=D,
                    seqlen_q=max(lens), seqlen_k=max(lens), q_batch_stride=0, k_batch_stride=0, v_batch_stride=0,
                    o_batch_stride=0, q_row_stride=H * D, k_row_stride=HK * D, v_row_stride=HK * D, o_row_stride=H * D,
                    q_head_stride=D, k_head_stride=D, v_head_stride=D, o_head_stride=D, softmax_scale=scale,
                    is_bf16=is_bf16, is_causal=causal, cu_seqlens_q=cu, cu_seqlens_k=cu, o_size=T * H * D)
    _, _, want = orc.run_mha(a, return_exact=True)
    return out, want.reshape(T, H, D), path, is_bf16, (q, k, v, cu_t)


@pytest.mark.parametrize("lens,H,HK,D,dtype,causal", [
    ([128], 4, 4, 128, torch.bfloat16, True),                 # exactly one full tile (pure TMA path)
    ([256, 384], 4, 2, 128, torch.bfloat16, True),            # full tiles only, GQA
    ([1, 77, 130, 64, 200, 127, 129], 8, 2, 128, torch.bfloat16, True),   # ragged tails everywhere
    ([300, 5, 513], 4, 1, 64, torch.bfloat16, True),          # head dim 64, MQA
    ([200, 333], 4, 2, 128, torch.float16, True),
    ([150, 260], 4, 4, 64, torch.float16, False),             # non-causal
    ([1000], 2, 1, 128, torch.bfloat16, True),                # 8 tiles deep
])
def test_prefill_tc_vs_oracle(lens, H, HK, D, dtype, causal):
    out, want, path, is_bf16, _ = _varlen(lens, H, HK, D, dtype, causal)
    assert path == "prefill_tc"
    assert_close_16(t2f32(out), want, is_bf16, f"prefill lens={lens} D={D} causal={causal}")


def test_prefill_tc_growing_maximum_exercises_the_lazy_rescale():
    """Keys whose scores grow along the sequence: the row maximum jumps by far more than 2^8 from
    tile to tile, so O has to be rescaled in tensor memory several times."""
    H, HK, D, S = 2, 2, 128, 640
    q = torch.zeros((S, H, D), dtype=torch.bfloat16, device=DEV)
    k = torch.zeros((S, HK, D), dtype=torch.bfloat16, device=DEV)
    q[:, :, 0] = 4.0
    k[:, :, 0] = (torch.arange(S, device=DEV).float() / 8.0).to(torch.bfloat16)[:, None]   # score = t / 2 * scale...
    k[:, :, 1:] = randn_t((S, HK, D - 1), 5, torch.bfloat16, DEV) * 0.1
    q[:, :, 1:] = randn_t((S, H, D - 1), 6, torch.bfloat16, DEV) * 0.1
    v = randn_t((S, HK, D), 7, torch.bfloat16, DEV)
    cu = u32([0, S])
    scale = 1.0   # scores up to 4 * 80 = 320 -> log2 range ~460: many corrections
    for causal in (True, False):
        out = ops.flash_attn_varlen(q, k, v, cu, cu, S, S, scale, causal)
        assert ops.last_mha_path() == "prefill_tc"
        want, _ = orc.attention_dense(t2f32(q)[None], t2f32(k)[None], t2f32(v)[None], scale, causal, True, exact=True)
        assert_close_16(t2f32(out), want[0], True, f"growing max causal={causal}")


def test_prefill_tc_dense_layout_and_bottom_right_alignment():
    """flash_attn (batched layout) with seqlen_q != seqlen_k: mask.h:171-172."""
    B, H, HK, D = 2, 4, 2, 128
    for (sq, sk) in [(100, 300), (300, 100), (256, 256), (130, 131)]:
        q = randn_t((B, sq, H, D), 1, torch.bfloat16, DEV)
        k = randn_t((B, sk, HK, D), 2, torch.bfloat16, DEV)
        v = randn_t((B, sk, HK, D), 3, torch.bfloat16, DEV)
        for causal in (True, False):
            out = ops.flash_attn(q, k, v, 0.088, causal)
            assert ops.last_mha_path() == "prefill_tc"
            want, _ = orc.attention_dense(t2f32(q), t2f32(k), t2f32(v), 0.088, causal, True, exact=True)
            assert_close_16(t2f32(out), want, True, f"dense {sq}x{sk} causal={causal}")


def test_prefill_tc_isolates_sequences_from_their_neighbours():
    """Tiles of one sequence overlap its neighbours' rows in the packed layout; NaN / inf there must
    not leak (the reference zero-fills what it does not use, flash_fwd_kernel.h:901-906)."""
    lens = [70, 200, 33]
    H, HK, D = 4, 2, 128
    out, want, path, _, (q, k, v, cu) = _varlen(lens, H, HK, D, torch.bfloat16, True, seed=20)
    k2, v2, q2 = k.clone(), v.clone(), q.clone()
    k2[:70], v2[:70], q2[:70] = float("nan"), float("inf"), float("nan")          # poison sequence 0 ...
    k2[270:], v2[270:] = float("nan"), float("-inf")                                # ... and sequence 2
    out2 = ops.flash_attn_varlen(q2, k2, v2, cu, cu, 200, 200, 1.0 / np.sqrt(D), True)
    assert torch.equal(out2[70:270], out[70:270])                                   # sequence 1 untouched


def test_prefill_tc_lse():
    lens = [130, 64]
    cu = u32([0, 130, 194])
    q = randn_t((194, 4, 64), 1, torch.bfloat16, DEV)
    k = randn_t((194, 2, 64), 2, torch.bfloat16, DEV)
    v = randn_t((194, 2, 64), 3, torch.bfloat16, DEV)
    out, lse = ops.flash_attn_varlen(q, k, v, cu, cu, 130, 130, 0.125, True, with_lse=True)
    assert ops.last_mha_path() == "prefill_tc"
    a = orc.MhaArgs(q=bits(q).ravel(), k=bits(k).ravel(), v=bits(v).ravel(), b=2, h=4, h_k=2, d=64, seqlen_q=130,
                    seqlen_k=130, q_batch_stride=0, k_batch_stride=0, v_batch_stride=0, o_batch_stride=0,
                    q_row_stride=256, k_row_stride=128, v_row_stride=128, o_row_stride=256, q_head_stride=64,
                    k_head_stride=64, v_head_stride=64, o_head_stride=64, softmax_scale=0.125, is_bf16=True,
                    is_causal=True, cu_seqlens_q=np.array([0, 130, 194], np.int32),
                    cu_seqlens_k=np.array([0, 130, 194], np.int32), o_size=194 * 256)
    _, want_lse = orc.run_mha(a)
    got = lse.cpu().numpy().reshape(4, 194)   # unpadded layout [h, total_q] (flash_fwd_kernel.h:44-47)
    for b, (s, n) in enumerate([(0, 130), (130, 64)]):
        np.testing.assert_allclose(got[:, s:s + n], want_lse[b, :, :n], rtol=2e-4, atol=2e-4)


@pytest.mark.skipif(not RefLib.available(), reason="oracle/_ref not built")
def test_prefill_tc_matches_reference_kernels_at_2k():
    ref = RefLib()
    lens = [2048, 1500, 548]
    H, HK, D = 32, 8, 128
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    q = randn_t((T, H, D), 1, torch.bfloat16, DEV)
    k = randn_t((T, HK, D), 2, torch.bfloat16, DEV)
    v = randn_t((T, HK, D), 3, torch.bfloat16, DEV)
    scale = 1.0 / np.sqrt(D)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    plan = ops.flash_attn_varlen_plan(q, k, v, cu, cu, max(lens), max(lens), scale, None, 0, num_sms=sms)
    want = ref.execute(plan, q)
    got = ops.flash_attn_varlen(q, k, v, cu, cu, max(lens), max(lens), scale, True)
    assert ops.last_mha_path() == "prefill_tc"
    # fp32 torch restatement as the "exact" side (the numpy oracle would take minutes at this size)
    exact = torch.empty((T, H, D), dtype=torch.float32, device=DEV)
    for b in range(len(lens)):
        s, e = int(cu[b]), int(cu[b + 1])
        qq = q[s:e].float().transpose(0, 1)                                   # [H, n, D]
        kk = k[s:e].float().transpose(0, 1).repeat_interleave(H // HK, dim=0)
        vv = v[s:e].float().transpose(0, 1).repeat_interleave(H // HK, dim=0)
        sc = torch.matmul(qq, kk.transpose(1, 2)) * scale
        n = e - s
        sc = sc.masked_fill(torch.triu(torch.ones(n, n, dtype=torch.bool, device=DEV), 1), float("-inf"))
        exact[s:e] = torch.matmul(torch.softmax(sc, -1), vv).transpose(0, 1)
    ra, rb = assert_same_up_to_rounding(t2f32(got), t2f32(want), t2f32(exact).astype(np.float64), True, "prefill 2k vs reference")
    print(f"rel err vs fp32: ours {ra:.3e}, reference kernels {rb:.3e}")


def test_prefill_tc_config3_full_prompt():
    """BASELINE.json configs[2]: one 8192-token prompt of Llama-3.1-8B (H=32, Hkv=8, D=128), causal."""
    S, H, HK, D = 8192, 32, 8, 128
    q = randn_t((S, H, D), 1, torch.bfloat16, DEV)
    k = randn_t((S, HK, D), 2, torch.bfloat16, DEV)
    v = randn_t((S, HK, D), 3, torch.bfloat16, DEV)
    cu = u32([0, S])
    scale = 1.0 / np.sqrt(D)
    out = ops.flash_attn_varlen(q, k, v, cu, cu, S, S, scale, True)
    assert ops.last_mha_path() == "prefill_tc"
    out2 = ops.flash_attn_varlen(q, k, v, cu, cu, S, S, scale, True)
    assert torch.equal(out, out2), "run-to-run determinism"
    mask = torch.triu(torch.ones(S, S, dtype=torch.bool, device=DEV), 1)
    for h in (0, 13, 31):   # three heads against fp32 torch (a full fp32 pass is 8.6 GB of scores)
        sc = torch.matmul(q[:, h].float(), k[:, h // (H // HK)].float().T) * scale
        want = torch.matmul(torch.softmax(sc.masked_fill(mask, float("-inf")), -1), v[:, h // (H // HK)].float())
        assert_close_16(t2f32(out[:, h]), t2f32(want).astype(np.float64), True, f"config 3 head {h}")
