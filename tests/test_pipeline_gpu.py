"""The decode-step chain (qkv GEMM -> reshape_and_cache_flash -> paged decode attention -> o_proj
GEMM, several layers) with programmatic dependent launch on/off, eagerly and replayed from a CUDA
graph: every variant must give bit-identical results (each kernel waits for its predecessor before
touching dependent memory; split reductions are order-fixed)."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from helpers import randn_t, random_block_table

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _chain(layers, x, wqkv, wo, kv, bt, lens, slots, H, HK, D, stable=False):
    nq, nkv = H * D, HK * D
    B = x.shape[0]
    for l in range(layers):
        qkv = ops.gemm_bf16_tn(x, wqkv[l])
        q = qkv[:, :nq].unflatten(1, (H, D)).unsqueeze(1)
        k = qkv[:, nq:nq + nkv].unflatten(1, (HK, D))
        v = qkv[:, nq + nkv:].unflatten(1, (HK, D))
        ops.reshape_and_cache_flash(k, v, kv[l][0], kv[l][1], slots)
        if stable:
            ops.promise_stable_cache()   # per call: the cache write above touches only final-page slots
        o = ops.flash_attn_kv_cache_full(q, kv[l][0], kv[l][1], None, D ** -0.5, bt, lens, True)
        x = ops.gemm_bf16_tn(o.reshape(B, nq), wo[l])
    return x


@pytest.fixture
def state():
    B, H, HK, D, page, MB, hidden, layers = 16, 32, 8, 128, 16, 40, 2048, 3
    NB = B * MB + 4
    wqkv = [randn_t(((H + 2 * HK) * D, hidden), 10 + l, torch.bfloat16, DEV, hidden ** -0.5) for l in range(layers)]
    wo = [randn_t((hidden, H * D), 20 + l, torch.bfloat16, DEV, (H * D) ** -0.5) for l in range(layers)]
    kv0 = [randn_t((2, NB, page, HK, D), 30 + l, torch.bfloat16, DEV) for l in range(layers)]
    bt = random_block_table(B, MB, NB, 5, DEV)
    rng = np.random.default_rng(3)
    lens_l = rng.integers(1, MB * page + 1, size=B).tolist()
    lens = torch.tensor(lens_l, dtype=torch.int32, device=DEV)
    slots = torch.tensor([int(bt[b, (n - 1) // page]) * page + (n - 1) % page for b, n in enumerate(lens_l)],
                         dtype=torch.int64, device=DEV)
    x = randn_t((B, hidden), 1, torch.bfloat16, DEV)
    yield dict(layers=layers, x=x, wqkv=wqkv, wo=wo, kv0=kv0, bt=bt, lens=lens, slots=slots, H=H, HK=HK, D=D)
    ops.set_launch_mode(False, False)


def _run(st, pdl, const_w, graph, stable_cache=False):
    ops.set_launch_mode(pdl, const_w)
    kv = [c.clone() for c in st["kv0"]]
    args = (st["layers"], st["x"], st["wqkv"], st["wo"], kv, st["bt"], st["lens"], st["slots"], st["H"], st["HK"], st["D"],
            stable_cache)
    if not graph:
        out = _chain(*args)
        torch.cuda.synchronize()
        return out.clone(), [c.clone() for c in kv]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _chain(*args)  # allocates library workspaces on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = _chain(*args)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    return out.clone(), [c.clone() for c in kv]


def test_chain_pdl_and_graph_variants_are_bit_identical(state):
    base, kv_base = _run(state, False, False, False)
    assert torch.isfinite(base.float()).all() and float(base.float().abs().max()) > 0
    for (pdl, cw, graph, sc) in [(True, False, False, False), (True, True, False, False), (False, False, True, False),
                                 (True, True, True, False), (True, True, False, True), (True, True, True, True)]:
        out, kv = _run(state, pdl, cw, graph, sc)
        assert torch.equal(out, base), f"pdl={pdl} const_weights={cw} graph={graph} stable_cache={sc}"
        for a, b in zip(kv, kv_base):
            assert torch.equal(a, b)


def test_stable_cache_promise_is_per_call_and_withdrawn_after_cache_maintenance(state):
    """The early-KV-streaming promise is one-shot and per call (no process-global bit), needs PDL, and the
    library withdraws it by itself when a cache-maintenance op (copy_blocks) is pending on the stream: that
    decode waits for its predecessors before releasing its dependents (bit 1) and reads the copied pages."""
    st = state
    H, HK, D = st["H"], st["HK"], st["D"]
    kv = st["kv0"][0].clone()
    q = randn_t((st["x"].shape[0], 1, H, D), 77, torch.bfloat16, DEV)
    call = lambda: ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, D ** -0.5, st["bt"], st["lens"], True)
    ops.set_launch_mode(True, True)
    base = call()
    assert ops.last_decode_mode() == 0                      # no promise -> no early streaming
    ops.promise_stable_cache()
    assert torch.equal(call(), base) and ops.last_decode_mode() == 1
    assert torch.equal(call(), base) and ops.last_decode_mode() == 0   # one-shot
    ops.set_launch_mode(False, False)
    ops.promise_stable_cache()
    assert torch.equal(call(), base) and ops.last_decode_mode() == 0   # needs programmatic dependent launch
    # copy the first page of sequence 0 somewhere else and point the block table at the copy
    ops.set_launch_mode(True, True)
    bt = st["bt"].clone()
    src = int(bt[0, 0])
    used = set(bt.flatten().tolist())
    dst = next(i for i in range(kv.shape[1]) if i not in used)
    kv[:, dst] = 0
    mapping = torch.tensor([[src, dst]], dtype=torch.int64, device=DEV)
    ops.copy_blocks([kv[0]], [kv[1]], mapping)
    bt[0, 0] = dst
    ops.promise_stable_cache()                              # a promise the library knows to be false here
    out = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, D ** -0.5, bt, st["lens"], True)
    assert ops.last_decode_mode() == 2
    torch.cuda.synchronize()
    assert torch.equal(out, base)
    ops.promise_stable_cache()
    assert torch.equal(call(), base) and ops.last_decode_mode() == 1   # the mark is cleared by the decode that honoured it
