"""The Python operator face refuses what the C entries would misread (CPU, no kernel runs): the decode-step
operators of include/atoma_llama.h take a pointer and a ROW stride per matrix, so a transposed view, a mismatched
dtype or shape, or a CPU tensor has to be stopped in the host layer — the analogue of the reference's `bail!`s in
csrc/src/lib.rs before it crosses its FFI."""
import pytest
import torch

from atoma_infer_b200 import ops

BF = torch.bfloat16


def _t(*shape, dtype=BF):
    return torch.zeros(shape, dtype=dtype)


@pytest.mark.parametrize("call, msg", [
    (lambda: ops.gemm_tn(_t(4, 64).t(), _t(8, 4)), "row-major"),
    (lambda: ops.gemm_tn(_t(4, 64), _t(64, 8).t()), "row-major"),
    (lambda: ops.gemm_tn(_t(4, 64), _t(8, 32)), r"w must be \[\*, 64\]"),
    (lambda: ops.gemm_tn(_t(4, 64), _t(8, 64, dtype=torch.float16)), "w must have dtype"),
    (lambda: ops.gemm_tn(_t(4, 64), _t(8, 64), out=_t(4, 16)), r"out must be \[4, 8\]"),
    (lambda: ops.gemm_tn(_t(2, 4, 64), _t(8, 64)), "2-D"),
    (lambda: ops.gemm_tn(_t(4, 64), _t(8, 64)), "CUDA tensor"),
    (lambda: ops.gemm_bf16_tn(_t(4, 64, dtype=torch.float16), _t(8, 64, dtype=torch.float16)), "dtype torch.bfloat16"),
    (lambda: ops.gate_up_silu(_t(4, 64), _t(7, 64)), "even number of rows"),
    (lambda: ops.gate_up_silu(_t(4, 64), _t(16, 32)), r"w_gate_up must be \[\*, 64\]"),
    (lambda: ops.mlp_gate_up_down(_t(4, 64), _t(32, 64), _t(64, 8)), r"w_down must be \[64, 16\]"),
    (lambda: ops.mlp_gate_up_down(_t(4, 64), _t(32, 64), _t(64, 16), act=_t(4, 8)), r"act must be \[4, 16\]"),
    (lambda: ops.gemm_tn_add_rmsnorm(_t(4, 64), _t(16, 64), _t(4, 8), _t(16), 1e-5), r"residual must be \[4, 16\]"),
    (lambda: ops.qkv_rope_cache(_t(4, 64), _t(100, 64), 2, 1, 16, torch.zeros(4, dtype=torch.int64), _t(8, 8), _t(8, 8)),
     "wqkv must be"),
    (lambda: ops.embedding(torch.zeros(3, dtype=torch.int32), _t(10, 16).t()), "row-major"),
    (lambda: ops.embedding(torch.zeros(3, dtype=torch.int64), _t(10, 16)), "u32"),
    (lambda: ops.sample_greedy(_t(4, 10, 2)), "2-D"),
    (lambda: ops.sample_greedy(_t(10, 4).t()), "row-major"),
])
def test_matrix_arguments_are_validated_before_the_c_call(call, msg):
    with pytest.raises(ops.AtomaError, match=msg):
        call()


def test_column_slices_of_wider_buffers_are_accepted_up_to_the_device_check():
    """Row-strided views (a slice of the fused qkv buffer, a padded logits row) are what the ABI's `ld*` arguments
    are for: they pass the layout checks and stop only at the device check on this CPU box."""
    x = _t(4, 128)[:, 32:96]
    with pytest.raises(ops.AtomaError, match="x must be a CUDA tensor"):
        ops.gemm_tn(x, _t(8, 64))
    with pytest.raises(ops.AtomaError, match="logits must be a CUDA tensor"):
        ops.sample_greedy(_t(3, 1008)[:, :1003])
