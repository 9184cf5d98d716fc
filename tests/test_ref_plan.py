"""oracle/ref_plan.py (the pure-Python derivation of the reference's kv-cache run_mha tuple that
`bench.py --impl reference` uses so that it loads nothing of this repo's product) against the C++
restatement of csrc/src/lib.rs:1523-1855 (`atoma_flash_attn_kv_cache_plan`, itself pinned to SURVEY
appendix A.1 / A.2 by tests/test_host_plan.py) — field by field.  CPU only."""
import pytest

from atoma_infer_b200 import _lib, ops
from oracle import ref_plan

FIELDS = ["q_batch_stride", "k_batch_stride", "v_batch_stride", "o_batch_stride", "q_row_stride", "k_row_stride",
          "v_row_stride", "o_row_stride", "q_head_stride", "k_head_stride", "v_head_stride", "o_head_stride",
          "num_splits", "b", "h", "h_k", "d", "d_rounded", "block_table_batch_stride", "page_block_size",
          "seqlen_q", "seqlen_k", "seqlen_q_rounded", "seqlen_k_rounded", "is_bf16", "is_causal",
          "window_size_left", "window_size_right"]


@pytest.mark.parametrize("B,H,HK,D,MB,NB,sms,qkv_view", [
    (32, 32, 8, 128, 256, 8256, 148, False),   # appendix A.1 (config 2)
    (16, 8, 1, 128, 256, 4200, 148, False),    # appendix A.2 (config 4 per rank): 2 splits
    (1, 32, 8, 64, 34, 100, 148, False),       # config 1
    (32, 32, 8, 128, 256, 8256, 148, True),    # q as a strided view of a fused qkv buffer
    (4, 4, 2, 64, 3, 40, 132, False),
])
def test_kv_cache_plan_matches_the_cpp_restatement(B, H, HK, D, MB, NB, sms, qkv_view):
    page = 16
    scale = D ** -0.5
    qs = [(H + 2 * HK) * D, (H + 2 * HK) * D, D, 1] if qkv_view else [H * D, H * D, D, 1]
    q = ops.spec([B, 1, H, D], _lib.BF16, stride=qs)
    kc = ops.spec([NB, page, HK, D], _lib.BF16, data=0x10000)
    vc = ops.spec([NB, page, HK, D], _lib.BF16, data=0x10000 + NB * page * HK * D * 2)
    bt = ops.spec([B, MB], _lib.U32, data=0x9000000)
    sl = ops.spec([B], _lib.U32, data=0x9100000)
    plan = ops.flash_attn_kv_cache_plan(q, kc, vc, scale, block_table=bt, seqlens_k=sl, window_size_right=0, num_sms=sms)
    mine = ref_plan.kv_cache_plan([B, 1, H, D], qs, [NB, page, HK, D], [page * HK * D, HK * D, D, 1],
                                  [page * HK * D, HK * D, D, 1], [B, MB], scale, True, sms, causal=True)
    for f in FIELDS:
        assert getattr(mine, f) == getattr(plan.args, f), f
    assert abs(mine.softmax_scale - plan.args.softmax_scale) < 1e-7
    assert abs(mine.scale_softmax_log2 - plan.args.scale_softmax_log2) < 1e-7
    assert plan.args.force_split_kernel and not plan.args.unpadded_lse and not plan.args.is_seqlens_k_cumulative
    assert mine.lse_elems == plan.lse_elems
    assert mine.lseaccum_elems == plan.lseaccum_elems and mine.oaccum_elems == plan.oaccum_elems


def test_ref_plan_module_does_not_import_the_product():
    import ast
    import inspect
    tree = ast.parse(inspect.getsource(ref_plan))
    names = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom):
            names.add(node.module or "")
        elif isinstance(node, ast.Import):
            names.update(a.name for a in node.names)
    assert not any("atoma_infer_b200" in n for n in names)
