"""The WHOLE tensor-parallel decode step on CPU with world_size 2 over gloo (SURVEY §8 row e, the N > 1 path).

What is under test is the product's host-side sharding — `LlamaShard.random_init(rank, world)` (fused qkv rows of the
rank's heads, o_proj columns, gate/up rows and down columns of the rank's slice of the MLP, replicated norms /
embedding / head: models/src/llama_nccl.rs:153-202, multi_gpu.rs:8-57), `tp.shard_kv_cache`, and the vocab-parallel
head's row ranges and combination rule (llama.py: DecodeStep; llama_ops.cu: argmax_tp_kernel) — driven exactly as
bench.py drives it for N > 1: every rank builds its shard from the same seed, runs its part of the step, the
row-parallel outputs are summed with a real all-reduce (gloo here, the fused push / receive kernels on the GPU) and
the ranks exchange their argmax candidates.  The per-rank arithmetic is the numpy oracle in float64 (no GPU here), so
the sharded step must reproduce the tp = 1 step of `oracle.llama_oracle.decode_step` to summation-order accuracy.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

PAGE, MB, T, SEED = 16, 3, 4, 11


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cfg():
    from atoma_infer_b200.llama import LlamaConfig, Llama3RopeScaling
    return LlamaConfig(hidden_size=64, intermediate_size=160, vocab_size=96, num_hidden_layers=2, num_attention_heads=8,
                       num_key_value_heads=4, max_position_embeddings=64,
                       rope_scaling=Llama3RopeScaling(original_max_position_embeddings=32))


def _batch(cfg):
    g = torch.Generator().manual_seed(5)
    nb = T * MB + 2
    kv = [torch.randn((2, nb, PAGE, cfg.num_key_value_heads, cfg.head_dim), generator=g) for _ in range(cfg.num_hidden_layers)]
    bt = torch.randperm(nb, generator=g)[: T * MB].to(torch.int32).reshape(T, MB).numpy()
    lens = np.array([PAGE * MB, 17, 1, 33])
    pos = lens - 1
    slots = bt[np.arange(T), pos // PAGE].astype(np.int64) * PAGE + pos % PAGE      # worker.rs:392-401
    tokens = torch.randint(0, cfg.vocab_size, (T,), generator=g).numpy()
    return kv, bt, lens, pos, slots, tokens


def _f(t):
    return t.double().numpy()


def _rank_step(shard, kv, bt, lens, pos, slots, tokens, all_reduce, exchange):
    """One rank's part of the step in float64: llama_step.cu's sequence with the oracle's primitives."""
    from oracle import attention_oracle as ao
    from oracle import llama_oracle as lo
    cfg, sh = shard.cfg, shard.shard
    h, hk, D, eps = sh.heads, sh.kv_heads, cfg.head_dim, cfg.rms_norm_eps
    cos, sin = _f(shard.cos), _f(shard.sin)
    x = _f(shard.embed)[tokens]
    for ly, cache in zip(shard.layers, kv):
        wq, wk, wv, wg, wu = lo.split_fused(_f(ly["wqkv"]), _f(ly["w_gate_up"]), h, hk, D)
        xn = lo.rms_norm(x, _f(ly["rms_attn"]), eps, exact=True)
        q = ao.rope((xn @ wq.T).reshape(T, h, D), pos, cos, sin)
        k = ao.rope((xn @ wk.T).reshape(T, hk, D), pos, cos, sin)
        v = (xn @ wv.T).reshape(T, hk, D)
        kc, vc = cache[0], cache[1]
        for t in range(T):
            kc[slots[t] // PAGE, slots[t] % PAGE], vc[slots[t] // PAGE, slots[t] % PAGE] = k[t], v[t]
        a = lo._paged_decode_fp64(q, kc, vc, bt, lens, D ** -0.5).reshape(T, h * D)
        x = x + all_reduce(a @ _f(ly["wo"]).T)                              # row parallel + all-reduce (multi_gpu.rs:141-179)
        xn = lo.rms_norm(x, _f(ly["rms_mlp"]), eps, exact=True)
        act = lo.silu(xn @ wg.T) * (xn @ wu.T)
        x = x + all_reduce(act @ _f(ly["w_down"]).T)
    xf = lo.rms_norm(x, _f(shard.norm_f), eps, exact=True)
    # vocab-parallel head: rows [r V/N, (r+1) V/N) of lm_head, local argmax, candidates exchanged, best kept
    rows = cfg.vocab_size // shard.world
    row0 = rows * shard.rank
    part = (xf @ _f(shard.lm_head)[row0: row0 + rows].T).astype(np.float32)
    idx, val = lo.greedy(part)
    cand_idx, cand_val = exchange(idx + row0, val)                        # [world, T] each, rank order
    tok = np.zeros(T, np.int64)
    for t in range(T):                                                   # argmax_tp_kernel: larger value, ties -> larger index
        best = 0
        for r in range(1, cand_idx.shape[0]):
            if (cand_val[r, t], cand_idx[r, t]) >= (cand_val[best, t], cand_idx[best, t]):
                best = r
        tok[t] = cand_idx[best, t]
    return xf, tok


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atoma_infer_b200 import tp
    from atoma_infer_b200.llama import LlamaShard
    cfg = _cfg()
    shard = LlamaShard.random_init(cfg, rank, world, dtype=torch.float32, device="cpu", seed=SEED, gen_device="cpu", max_pos=64)
    kv_full, bt, lens, pos, slots, tokens = _batch(cfg)
    kv = [_f(tp.shard_kv_cache(c, shard.shard)) for c in kv_full]        # this rank's kv heads (worker.rs:584-591)

    def all_reduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    def exchange(idx, val):
        gi = [torch.zeros(T, dtype=torch.int64) for _ in range(world)]
        gv = [torch.zeros(T, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(gi, torch.from_numpy(idx.astype(np.int64)))
        dist.all_gather(gv, torch.from_numpy(val.astype(np.float32)))
        return torch.stack(gi).numpy(), torch.stack(gv).numpy()

    xf, tok = _rank_step(shard, kv, bt, lens, pos, slots, tokens, all_reduce, exchange)
    np.savez(Path(out_dir) / f"r{rank}.npz", xf=xf, tok=tok, **{f"kv{l}": c for l, c in enumerate(kv)})
    dist.destroy_process_group()


def test_tensor_parallel_step_over_gloo_matches_tp1(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from atoma_infer_b200.llama import LlamaShard
    from oracle import llama_oracle as lo
    cfg = _cfg()
    full = LlamaShard.random_init(cfg, 0, 1, dtype=torch.float32, device="cpu", seed=SEED, gen_device="cpu", max_pos=64)
    kv_full, bt, lens, pos, slots, tokens = _batch(cfg)
    kv1 = [_f(c) for c in kv_full]
    H, HK, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    layers = []
    for ly in full.layers:
        wq, wk, wv, wg, wu = lo.split_fused(_f(ly["wqkv"]), _f(ly["w_gate_up"]), H, HK, D)
        layers.append(lo.LayerWeights(wq, wk, wv, _f(ly["wo"]), wg, wu, _f(ly["w_down"]), _f(ly["rms_attn"]), _f(ly["rms_mlp"])))
    xf1, logits1 = lo.decode_step(_f(full.embed)[tokens], layers, kv1, _f(full.norm_f), _f(full.lm_head), _f(full.cos),
                                  _f(full.sin), pos, slots, bt, lens, H, HK, D, cfg.rms_norm_eps, exact=True)
    tok1, _ = lo.greedy(logits1.astype(np.float32))
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for r, o in enumerate(outs):
        assert np.allclose(o["xf"], xf1, rtol=1e-9, atol=1e-11), f"rank {r}: final hidden state"
        assert o["tok"].tolist() == tok1.tolist(), f"rank {r}: vocab-parallel greedy tokens"
        hs = slice(r * HK // world, (r + 1) * HK // world)
        for l in range(cfg.num_hidden_layers):                           # each rank wrote its heads of the new token's k/v
            assert np.allclose(o[f"kv{l}"], kv1[l][:, :, :, hs], rtol=1e-9, atol=1e-11), f"rank {r} layer {l}: cache"
    assert np.array_equal(outs[0]["xf"], outs[1]["xf"]), "every rank holds the same all-reduced state"


class _FakeComm:
    """What DecodeStep reads of a P2PComm (regions array + slot size) — enough to marshal the model on the host."""

    def __init__(self, world, slot_bytes):
        import ctypes as C
        self._regions = (C.c_void_p * world)(*[0x1000 * (r + 1) for r in range(world)])
        self.slot_bytes = slot_bytes


def test_decode_step_marshals_the_rank_shard_for_the_c_executor():
    """DecodeStep's atoma_llama_model for rank 1 of 2 (host logic only, no launch): per-rank heads / inter, cache
    geometry, and the vocab-parallel head = rows [V/2, V) of lm_head."""
    from atoma_infer_b200.llama import DecodeStep, LlamaShard
    from atoma_infer_b200.ops import AtomaError
    import pytest
    cfg = _cfg()
    shard = LlamaShard.random_init(cfg, 1, 2, dtype=torch.bfloat16, device="cpu", seed=SEED, gen_device="cpu", max_pos=64)
    hk, D = shard.shard.kv_heads, cfg.head_dim
    kv = [torch.zeros((2, 9, PAGE, hk, D), dtype=torch.bfloat16) for _ in range(cfg.num_hidden_layers)]
    step = DecodeStep(shard, kv, max_batch=T, p2p=_FakeComm(2, 1 << 16))
    m = step._model
    assert (m.hidden, m.inter, m.num_layers) == (64, 80, 2) and (m.num_heads, m.num_kv_heads, m.head_dim) == (4, 2, 8)
    assert (m.block_size, m.block_stride, m.vocab, m.dtype) == (PAGE, PAGE * hk * D, 96, 1) and (m.tp_rank, m.tp_world) == (1, 2)
    assert step.vocab_parallel_head and (m.lm_head_rows, m.lm_head_row0) == (48, 48)
    assert m.lm_head == shard.lm_head.data_ptr() + 48 * 64 * 2
    assert step._layers[1].wqkv == shard.layers[1]["wqkv"].data_ptr() and step._layers[0].value_cache == kv[0][1].data_ptr()
    assert shard.layers[0]["wqkv"].shape == ((4 + 2 * 2) * D, 64) and shard.layers[0]["w_gate_up"].shape == (160, 64)
    assert shard.layers[0]["wo"].shape == (64, 4 * D) and shard.layers[0]["w_down"].shape == (64, 80)
    # callers that want logits get the replicated head
    full_head = DecodeStep(shard, kv, max_batch=T, p2p=_FakeComm(2, 1 << 16), want_logits=True)
    assert not full_head.vocab_parallel_head and full_head._model.lm_head_rows == 0
    with pytest.raises(AtomaError):
        DecodeStep(shard, kv, max_batch=T, p2p=_FakeComm(2, 1 << 16), want_logits=True, vocab_parallel_head=True)
    with pytest.raises(AtomaError):
        DecodeStep(shard, kv, max_batch=T)                       # world 2 without a communicator
    with pytest.raises(AtomaError):
        DecodeStep(shard, kv[:1], max_batch=T, p2p=_FakeComm(2, 1 << 16))
    with pytest.raises(AtomaError):
        step.launch()                                            # nothing bound
