"""Tensor-parallel host logic on CPU with world_size 2 over gloo: the head-sharded attention path
(column-parallel qkv, per-rank KV cache slice, row-parallel o_proj + all-reduce) must reproduce the
unsharded result.  The per-rank math is the numpy oracle (no GPU here); what is under test is the
sharding in atoma_infer_b200/tp.py, which bench.py uses for N > 1."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

H, HK, D, HIDDEN, PAGE, MB, B = 8, 4, 16, 64, 16, 3, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    nb = B * MB
    return dict(wq=r(H * D, HIDDEN) * HIDDEN ** -0.5, wk=r(HK * D, HIDDEN) * HIDDEN ** -0.5,
                wv=r(HK * D, HIDDEN) * HIDDEN ** -0.5, wo=r(HIDDEN, H * D) * (H * D) ** -0.5,
                kv=r(2, nb, PAGE, HK, D), x=r(B, HIDDEN),
                bt=torch.randperm(nb, generator=g).to(torch.int32).reshape(B, MB),
                lens=np.array([PAGE * MB, 17, 1]))


def _path(x, wqkv, wo, kv, bt, lens, h, hk):
    """qkv GEMM -> cache write -> paged decode attention -> o_proj, fp64 oracle math."""
    from oracle import attention_oracle as orc
    qkv = x.double() @ wqkv.double().T
    nq, nkv = h * D, hk * D
    q = qkv[:, :nq].reshape(B, 1, h, D).numpy()
    k, v = qkv[:, nq:nq + nkv].reshape(B, hk, D), qkv[:, nq + nkv:].reshape(B, hk, D)
    kv = kv.clone().double()
    for b, n in enumerate(lens):  # slot of the newest token (worker.rs:392-401)
        blk, off = int(bt[b, (n - 1) // PAGE]), (n - 1) % PAGE
        kv[0, blk, off], kv[1, blk, off] = k[b], v[b]
    o, _ = orc.attention_paged_decode(q.astype(np.float32), kv[0].numpy().astype(np.float32),
                                      kv[1].numpy().astype(np.float32), bt.numpy(), lens, D ** -0.5, False, exact=True)
    return torch.from_numpy(o.reshape(B, nq)) @ wo.double().T


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atoma_infer_b200 import tp
    m = _model()
    sh = tp.HeadShard(rank, world, H, HK, D)
    y = _path(m["x"], tp.shard_qkv_weight(m["wq"], m["wk"], m["wv"], sh), tp.shard_o_weight(m["wo"], sh),
              tp.shard_kv_cache(m["kv"], sh), m["bt"], m["lens"], sh.heads, sh.kv_heads)
    tp.all_reduce_sum(y)
    torch.save(y, Path(out_dir) / f"y{rank}.pt")
    dist.destroy_process_group()


def test_head_sharded_path_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from atoma_infer_b200 import tp
    m = _model()
    full = tp.HeadShard(0, 1, H, HK, D)
    want = _path(m["x"], tp.shard_qkv_weight(m["wq"], m["wk"], m["wv"], full), tp.shard_o_weight(m["wo"], full),
                 m["kv"], m["bt"], m["lens"], H, HK)
    for r in range(world):
        got = torch.load(tmp_path / f"y{r}.pt")
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-8), f"rank {r}"  # fp64 math, fp32-cast inputs


def test_shard_shapes_and_errors():
    from atoma_infer_b200 import tp
    sh = tp.HeadShard(1, 2, 32, 8, 128)
    assert (sh.heads, sh.kv_heads, sh.q_range, sh.kv_range) == (16, 4, slice(16, 32), slice(4, 8))
    w = torch.arange(32 * 128 * 4, dtype=torch.float32).reshape(32 * 128, 4)
    wk = torch.arange(8 * 128 * 4, dtype=torch.float32).reshape(8 * 128, 4)
    f = tp.shard_qkv_weight(w, wk, wk + 1, sh)
    assert f.shape == ((16 + 8) * 128, 4)
    assert torch.equal(f[: 16 * 128], w[16 * 128:]) and torch.equal(f[16 * 128: 20 * 128], wk[4 * 128:])
    wo = torch.arange(4 * 32 * 128, dtype=torch.float32).reshape(4, 32 * 128)
    assert torch.equal(tp.shard_o_weight(wo, sh), wo[:, 16 * 128:])
    with pytest.raises(ValueError):
        tp.HeadShard(0, 3, 32, 8, 128)
