"""Host operator layer (C++ restatement of csrc/src/lib.rs / cache_manager.rs): derived run_mha
argument tuples against SURVEY.md Appendix A (hand-derived from the Rust), validation error texts the
reference's tests assert, and the split heuristic.  Plan-only: runs without a GPU."""
import math

import pytest

from atoma_infer_b200 import _lib, ops
from atoma_infer_b200.ops import AtomaError, spec

BF16, F16, U32, I64, F32 = _lib.BF16, _lib.F16, _lib.U32, _lib.I64, _lib.F32
SMS = 148


def fields(args):
    return {n: getattr(args, n) for n, _ in args._fields_}


def test_a1_decode_config2():
    """Appendix A.1: Llama-3.1-8B decode, B=32, ctx 4096 (csrc/src/lib.rs:1523-1855)."""
    NB = 8192 + 64
    q = spec([32, 1, 32, 128], BF16, stride=[4096, 4096, 128, 1], data=0x10000)
    kc = spec([NB, 16, 8, 128], BF16, data=0x20000000)
    vc = spec([NB, 16, 8, 128], BF16, data=0x20000000 + NB * 16384 * 2)
    bt = spec([32, 256], U32, data=0x30000)
    sl = spec([32], U32, data=0x40000)
    scale = 1.0 / math.sqrt(128)
    p = ops.flash_attn_kv_cache_plan(q, kc, vc, scale, block_table=bt, seqlens_k=sl,
                                     window_size_left=None, window_size_right=0, num_sms=SMS)
    a = fields(p.args)
    assert a["cu_seqlens_q_ptr"] is None and a["cu_seqlens_k_ptr"] == 0x40000
    assert a["is_seqlens_k_cumulative"] is False
    assert (a["q_batch_stride"], a["q_head_stride"]) == (4096, 128)
    assert (a["k_batch_stride"], a["v_batch_stride"]) == (16384, 16384)
    assert (a["k_row_stride"], a["v_row_stride"]) == (1024, 1024)
    assert (a["k_head_stride"], a["v_head_stride"]) == (128, 128)
    assert (a["o_batch_stride"], a["o_row_stride"], a["o_head_stride"]) == (4096, 4096, 128)
    assert a["num_splits"] == 1
    assert (a["b"], a["h"], a["h_k"], a["d"], a["d_rounded"]) == (32, 32, 8, 128, 128)
    assert a["softmax_scale"] == pytest.approx(0.088388346, rel=1e-6)
    assert a["scale_softmax_log2"] == pytest.approx(0.12751743, rel=1e-6)
    assert (a["block_table"], a["block_table_batch_stride"], a["page_block_size"]) == (0x30000, 256, 16)
    assert a["seqused_k"] is None
    assert (a["seqlen_q"], a["seqlen_k"], a["seqlen_q_rounded"], a["seqlen_k_rounded"]) == (1, 4096, 128, 4096)
    assert (a["is_bf16"], a["is_causal"]) == (1, 0)
    assert (a["window_size_left"], a["window_size_right"]) == (4096, 0)
    assert a["softcap"] == 0.0
    assert (a["unpadded_lse"], a["force_split_kernel"]) == (False, True)
    assert a["softmax_lseaccum_ptr"] is None and a["oaccum_ptr"] is None
    assert p.lse_elems == 32 * 32 and p.lseaccum_elems == 0 and p.oaccum_elems == 0
    assert [p.out_shape[i] for i in range(4)] == [32, 1, 32, 128]


def test_a2_decode_config4_per_rank():
    """Appendix A.2: 70B TP=8 per rank: b=16, h=8, h_k=1 -> num_splits = 2 and accumulators."""
    q = spec([16, 1, 8, 128], BF16)
    kc = spec([4160, 16, 1, 128], BF16)
    vc = spec([4160, 16, 1, 128], BF16)
    p = ops.flash_attn_kv_cache_plan(q, kc, vc, 0.0884, block_table=spec([16, 256], U32), seqlens_k=spec([16], U32),
                                     window_size_right=0, num_sms=SMS)
    a = fields(p.args)
    assert (a["q_batch_stride"], a["q_head_stride"], a["o_batch_stride"]) == (1024, 128, 1024)
    assert (a["k_batch_stride"], a["k_row_stride"]) == (2048, 128)
    assert a["num_splits"] == 2
    assert p.lseaccum_elems == 2 * 16 * 8 * 1 and p.oaccum_elems == 2 * 16 * 8 * 1 * 128


def test_a3_prefill_config3():
    """Appendix A.3: flash_attn_varlen, 8192-token causal prompt (csrc/src/lib.rs:606-1103)."""
    q = spec([8192, 32, 128], BF16)
    k = spec([8192, 8, 128], BF16)
    v = spec([8192, 8, 128], BF16)
    cu = spec([2], U32, data=0x5000)
    p = ops.flash_attn_varlen_plan(q, k, v, cu, cu, 8192, 8192, 0.0884, None, 0, num_sms=SMS)
    a = fields(p.args)
    assert a["cu_seqlens_q_ptr"] == 0x5000 and a["cu_seqlens_k_ptr"] == 0x5000
    assert a["is_seqlens_k_cumulative"] is True
    assert (a["q_batch_stride"], a["k_batch_stride"], a["v_batch_stride"], a["o_batch_stride"]) == (0, 0, 0, 0)
    assert (a["q_row_stride"], a["q_head_stride"], a["k_row_stride"], a["k_head_stride"]) == (4096, 128, 1024, 128)
    assert (a["num_splits"], a["b"], a["block_table"], a["page_block_size"]) == (0, 1, None, 1)
    assert (a["is_causal"], a["window_size_left"], a["window_size_right"]) == (1, 8192, 0)
    assert (a["unpadded_lse"], a["force_split_kernel"]) == (True, False)
    assert p.lse_elems == 8192 * 32 and p.out_rank == 3


def test_varlen_with_block_table_uses_split_kernel_flag():
    """csrc/src/lib.rs:689-690, 772-776, 1088: paged varlen -> page_block_size = k.dims()[1],
    force_split_kernel = true, non-causal when no window is given (flash_attention.rs:445-446)."""
    q = spec([64, 2, 8], F16)
    k = spec([4, 16, 2, 8], F16)
    cu = spec([3], U32)
    p = ops.flash_attn_varlen_plan(q, k, k, cu, cu, 32, 32, 0.5, None, None, block_table=spec([2, 2], U32))
    a = fields(p.args)
    assert (a["page_block_size"], a["force_split_kernel"], a["is_causal"]) == (16, True, 0)
    assert (a["k_batch_stride"], a["k_row_stride"], a["k_head_stride"]) == (256, 16, 8)
    assert a["block_table_batch_stride"] == 2


def test_compute_num_splits_cases():
    """SURVEY §8a8 (csrc/src/lib.rs:2122-2199)."""
    assert ops.compute_num_splits(32, 32, 128, 4096, 1, SMS) == 1      # config 2
    assert ops.compute_num_splits(16, 8, 128, 4096, 1, SMS) == 2       # config 4 per rank
    assert ops.compute_num_splits(1, 32, 64, 256, 1, SMS) == 1         # config 1, ctx <= 256
    assert ops.compute_num_splits(1, 32, 64, 512, 1, SMS) == 2
    assert ops.compute_num_splits(1, 32, 64, 768, 1, SMS) == 3
    # upstream's documented example: 48 work items on 108 SMs -> 2 splits (efficiency 0.89)
    assert ops.compute_num_splits(1, 48, 128, 128 * 64, 1, 54) == 2


def test_compute_num_splits_sweep_against_the_restated_heuristic():
    """The C++ host layer against the oracle's restatement of lib.rs:2122-2199 on a seeded sweep that
    covers the 0.8-occupancy shortcut, ineligible split counts and the 85 % rule."""
    import numpy as np
    from oracle import attention_oracle as orc
    rng = np.random.default_rng(0)
    seen = set()
    for _ in range(3000):
        b = int(rng.integers(1, 65)); h = int(rng.choice([1, 2, 4, 8, 16, 32, 40, 64]))
        d = int(rng.choice([32, 64, 96, 128, 192, 256])); sk = int(rng.integers(1, 20000))
        sq = int(rng.choice([1, 1, 1, 7, 64, 65, 200])); sms = int(rng.choice([54, 108, 132, 148]))
        want = orc.compute_num_splits(b, h, d, sk, sq, sms)
        assert ops.compute_num_splits(b, h, d, sk, sq, sms) == want, (b, h, d, sk, sq, sms)
        seen.add(want)
    assert len(seen) > 10 and 1 in seen


def test_kv_cache_decode_folds_causal_only_for_single_query():
    q1 = spec([2, 1, 4, 64], BF16)
    q3 = spec([2, 3, 4, 64], BF16)
    kc = spec([2, 128, 4, 64], BF16)
    assert ops.flash_attn_kv_cache_plan(q1, kc, kc, 0.1, window_size_right=0, num_sms=SMS).args.is_causal == 0
    p = ops.flash_attn_kv_cache_plan(q3, kc, kc, 0.1, window_size_right=0, num_sms=SMS)
    assert p.args.is_causal == 1 and p.args.page_block_size == 1 and p.args.block_table is None
    assert p.args.is_seqlens_k_cumulative is True


def test_dense_swapped_layout_is_restated():
    """csrc/src/lib.rs:69-100, 259-266: seqlen_q == 1 with GQA and no window re-reads q as
    (b, ngroups, hk, d) and multiplies the batch strides by ngroups."""
    q = spec([2, 1, 8, 64], F16)
    k = spec([2, 40, 2, 64], F16)
    p = ops.flash_attn_plan(q, k, k, 0.125, None, None, num_sms=SMS)
    a = fields(p.args)
    assert p.seqlenq_ngroups_swapped == 1
    assert (a["seqlen_q"], a["h"], a["h_k"]) == (4, 2, 2)
    assert (a["q_row_stride"], a["q_head_stride"]) == (128, 64)
    assert a["q_batch_stride"] == 4 * 2 * 64 * 4
    assert [p.out_shape[i] for i in range(4)] == [2, 1, 8, 64]


@pytest.mark.parametrize("call,msg", [
    (lambda: ops.flash_attn_kv_cache_plan(spec([2, 1, 4, 64], BF16), spec([4, 24, 4, 64], BF16), spec([4, 24, 4, 64], BF16),
                                          0.1, block_table=spec([2, 2], U32)),
     "page_block_size must be a multiple of 16 when block_table is provided"),
    (lambda: ops.flash_attn_kv_cache_plan(spec([2, 1, 4, 64], BF16), spec([4, 16, 4, 64], F16), spec([4, 16, 4, 64], BF16), 0.1),
     "query and key must have the same dtype"),
    (lambda: ops.flash_attn_kv_cache_plan(spec([2, 1, 4, 64], F32), spec([4, 16, 4, 64], F32), spec([4, 16, 4, 64], F32), 0.1),
     "flash-attn is only supported for f16/bf16"),
    (lambda: ops.flash_attn_varlen_plan(spec([8, 4, 12], F16), spec([8, 4, 12], F16), spec([8, 4, 12], F16),
                                        spec([2], U32), spec([2], U32), 8, 8, 0.1),
     "only supports head sizes that are a multiple of 8 (got 12)"),
    (lambda: ops.flash_attn_varlen_plan(spec([8, 4, 264], F16), spec([8, 4, 264], F16), spec([8, 4, 264], F16),
                                        spec([2], U32), spec([2], U32), 8, 8, 0.1),
     "only supports head dimension at most 256 (got 264)"),
    (lambda: ops.flash_attn_varlen_plan(spec([8, 6, 16], F16), spec([8, 4, 16], F16), spec([8, 4, 16], F16),
                                        spec([2], U32), spec([2], U32), 8, 8, 0.1),
     "number of k/v heads 4 must divide number of heads in query 6"),
    (lambda: ops.flash_attn_kv_cache_plan(spec([2, 1, 4, 64], BF16, stride=[256, 256, 64, 2]), spec([4, 16, 4, 64], BF16),
                                          spec([4, 16, 4, 64], BF16), 0.1),
     "the last dim of q must be contiguous"),
    (lambda: ops.flash_attn_kv_cache_plan(spec([2, 1, 4, 64], BF16), spec([4, 16, 4, 64], BF16), spec([4, 16, 4, 64], BF16),
                                          0.1, block_table=spec([3, 2], U32)),
     "shape mismatch of block_table"),
])
def test_validation_messages(call, msg):
    with pytest.raises(AtomaError) as e:
        call()
    assert msg in str(e.value)


def test_cache_op_error_texts_of_the_reference_tests():
    """csrc/tests/cache_manager_tests.rs:457, 476, 492, 508, 524 assert these exact texts."""
    import torch
    k16 = torch.zeros(4, 64, 2, 8, dtype=torch.float16)
    kbf = torch.zeros(4, 64, 2, 8, dtype=torch.bfloat16)
    k32 = torch.zeros(4, 64, 2, 8, dtype=torch.float32)
    bm = torch.zeros(2, 2, dtype=torch.int64)
    with pytest.raises(AtomaError, match="device must be a cuda device"):
        ops.copy_blocks([k16], [k16], bm)
    with pytest.raises(AtomaError, match="Only support f16/bf16 dtypes and src and dst must have same dtype"):
        ops.copy_blocks([k16], [kbf], bm)
    with pytest.raises(AtomaError, match="Only support f16/bf16 dtypes and src and dst must have same dtype"):
        ops.copy_blocks([k32], [k32], bm)
    # the two remaining texts need cuda-resident descriptors: use plan-less specs through the C ABI
    import ctypes as C
    L = _lib.lib()
    kd = (_lib.Tensor * 2)(spec([4, 64, 2, 8], F16), spec([4, 64, 2, 8], F16))
    vd = (_lib.Tensor * 1)(spec([4, 64, 2, 8], F16))
    bmd = spec([2, 2], I64)
    assert L.atoma_copy_blocks(kd, 2, vd, 1, C.byref(bmd), None, None) == b"key_caches and value_caches must have the same length"
    bad = spec([2, 3], I64)
    assert L.atoma_copy_blocks(kd, 1, vd, 1, C.byref(bad), None, None) == b"block_mapping must have shape [num_pairs, 2]"
    # swap_blocks: both on the CPU (cache_manager_tests.rs:189-191)
    with pytest.raises(AtomaError, match="swap_blocks: Either src and dst are on the same cuda device, or src and dst "
                                         "are on cpu and cuda devices, alternately"):
        ops.swap_blocks(k16, k16.clone(), {0: 1})


def test_reshape_and_cache_validation():
    """csrc/tests/cache_manager_tests.rs:686-769 (dtype / shape / slot-mapping errors)."""
    import ctypes as C
    L = _lib.lib()
    key, val = spec([10, 2, 8], F16), spec([10, 2, 8], F16)
    kc, vc = spec([4, 8, 2, 8], F16), spec([4, 8, 2, 8], F16)
    sm = spec([10], I64)
    call = lambda *t: L.atoma_reshape_and_cache_flash(*[C.byref(x) for x in t], None)
    assert b"must have same dtype" in call(key, spec([10, 2, 8], BF16), kc, vc, sm)
    assert b"Only support key_cache with shape [4, 8, 2, 8]" in call(key, val, spec([4, 8, 2, 16], F16, stride=[128, 16, 8, 1]), vc, sm)
    assert b"Only support slot_mapping with shape [10]" in call(key, val, kc, vc, spec([9], I64))
    assert b"Only support value with shape [10, 2, 8]" in call(key, spec([9, 2, 8], F16), kc, vc, sm)
    assert b"Only support block_stride == value_cache.stride[0]" in call(key, val, kc, spec([4, 8, 2, 8], F16, stride=[256, 16, 8, 1]), sm)
