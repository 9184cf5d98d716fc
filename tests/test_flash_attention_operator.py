"""The CPU-runnable unit tests of the reference's operator (`models/src/flash_attention.rs:477-548`:
test_new, test_new_invalid_heads, test_new_invalid_head_dim, test_supported_head_sizes,
test_get_kv_cache_shape, test_split_kv_cache, test_split_kv_cache_invalid_shape) against
atoma_infer_b200.flash_attention.FlashAttention, plus `FlashAttentionMetadata::new` (:87-146).
`test_forward` / `test_forward_with_varlen` need a GPU: tests/test_attention_gpu.py."""
import pytest
import torch

from atoma_infer_b200.flash_attention import FlashAttention, FlashAttentionMetadata
from atoma_infer_b200.ops import AtomaError


def test_new():
    fa = FlashAttention(8, 4, 64, 1.0, None, None, torch.float32, "cpu")
    assert (fa.num_heads, fa.num_kv_heads, fa.num_queries_per_kv, fa.head_dim) == (8, 4, 2, 64)
    assert fa.softmax_scale == 1.0 and fa.kv_cache_dtype == torch.float32


def test_new_invalid_heads():
    with pytest.raises(AtomaError):
        FlashAttention(7, 4, 64, 1.0, None, None, torch.float32, "cpu")


def test_new_invalid_head_dim():
    with pytest.raises(AtomaError):
        FlashAttention(8, 4, 65, 1.0, None, None, torch.float32, "cpu")


def test_supported_head_sizes():
    assert FlashAttention.supported_head_sizes() == [64, 80, 96, 112, 128, 192, 256]


def test_get_kv_cache_shape():
    assert FlashAttention.get_kv_cache_shape(10, 32, 4, 64) == [2, 10, 32, 4, 64]


def test_split_kv_cache():
    fa = FlashAttention(8, 4, 64, 1.0, None, None, torch.float32, "cpu")
    k, v = fa.split_kv_cache(torch.zeros((2, 10, 32, 4, 64)))
    assert list(k.shape) == [10, 32, 4, 64] and list(v.shape) == [10, 32, 4, 64]
    assert v.data_ptr() - k.data_ptr() == 10 * 32 * 4 * 64 * 4   # two views of ONE allocation (:274-275)


@pytest.mark.parametrize("shape", [(2, 10, 32, 4), (3, 10, 32, 4, 64), (2, 10, 32, 5, 64), (2, 10, 32, 4, 32)])
def test_split_kv_cache_invalid_shape(shape):
    fa = FlashAttention(8, 4, 64, 1.0, None, None, torch.float32, "cpu")
    with pytest.raises(AtomaError):
        fa.split_kv_cache(torch.zeros(shape))


def test_metadata_new_slices_prefill_and_decode_parts():
    """flash_attention.rs:87-146 with prefix_caching = false (what the worker passes, worker.rs:449):
    prefill gets no block tables, decode gets the WHOLE table; start locations / lengths are split at
    the number of prefill sequences."""
    bt = torch.arange(12, dtype=torch.int32).reshape(4, 3)
    seq_lens = torch.tensor([5, 7, 20, 33], dtype=torch.int32)
    seq_start = torch.tensor([0, 5, 12, 32, 65], dtype=torch.int32)
    q_start = torch.tensor([0, 5, 12, 13, 14], dtype=torch.int32)
    md = FlashAttentionMetadata.new(torch.tensor([0, 0, 19, 32]), torch.arange(14), q_start, 12, 2, 7, 33, 7, 2,
                                    seq_start, seq_lens, bt, False)
    assert md.num_prefill_tokens == 12 and md.num_decoding_tokens == 2
    p, d = md.prefill_metadata, md.decoding_metadata
    assert p.block_tables is None and p.max_query_length == 7 and p.max_prefill_sequence_length == 7
    assert p.sequence_start_locations.tolist() == [0, 5, 12] and p.sequence_lengths.tolist() == [5, 7]
    assert p.query_start_locations.tolist() == [0, 5]
    assert d.block_tables is bt and d.sequence_lengths.tolist() == [20, 33] and d.max_decoding_sequence_length == 33
    only_decode = FlashAttentionMetadata.new(None, torch.arange(2), q_start, 0, 2, 1, 33, 0, 0, seq_start, seq_lens, bt, False)
    assert only_decode.prefill_metadata is None and only_decode.decoding_metadata is not None
    only_prefill = FlashAttentionMetadata.new(None, torch.arange(12), q_start, 12, 0, 7, 0, 7, 2, seq_start, seq_lens, bt, False)
    assert only_prefill.decoding_metadata is None
