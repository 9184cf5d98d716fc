"""Tensor-parallel sharding of the attention path, as the reference does it
(models/src/llama_nccl.rs:153-171, models/src/multi_gpu.rs:8-57, backends/vllm/src/worker.rs:584-591):

  * q/k/v projections are column parallel — rank r keeps the weight rows of its H/N query heads and
    Hkv/N kv heads (`TensorParallelColumnLinear`, shard along dim 0);
  * the KV cache of rank r holds only its Hkv/N kv heads;
  * o_proj is row parallel — rank r keeps the weight COLUMNS of its query heads (shard along dim 1)
    and the partial outputs are summed with one all-reduce (`TensorParallelRowLinear` + `AllReduce`);
  * block tables, slot mappings and sequence lengths are identical on every rank.

Attention, cache write, copy and swap need no communication.  These helpers are pure host logic
(torch tensors in, views out); the collective itself stays NCCL (`torch.distributed.all_reduce`).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class HeadShard:
    rank: int
    world: int
    num_heads: int       # global H
    num_kv_heads: int    # global Hkv
    head_dim: int

    def __post_init__(self):
        if self.num_heads % self.world or self.num_kv_heads % self.world:
            raise ValueError(f"tensor parallel degree {self.world} must divide heads {self.num_heads} "
                             f"and kv heads {self.num_kv_heads} (llama_nccl.rs:164-171)")
        if (self.num_heads // self.num_kv_heads) * self.num_kv_heads != self.num_heads:
            raise ValueError("num_heads must be a multiple of num_kv_heads")

    @property
    def heads(self) -> int:
        return self.num_heads // self.world

    @property
    def kv_heads(self) -> int:
        return self.num_kv_heads // self.world

    @property
    def q_range(self) -> slice:
        """Global query heads of this rank.  Consecutive heads, so each rank's query heads map onto
        exactly its own kv heads (kv_head = q_head // (H / Hkv), flash_fwd_kernel.h:593)."""
        return slice(self.rank * self.heads, (self.rank + 1) * self.heads)

    @property
    def kv_range(self) -> slice:
        return slice(self.rank * self.kv_heads, (self.rank + 1) * self.kv_heads)


def shard_qkv_weight(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, sh: HeadShard) -> torch.Tensor:
    """wq [H*D, hidden], wk/wv [Hkv*D, hidden] -> this rank's fused [(h + 2 hkv) * D, hidden] weight
    (rows = q heads, then k heads, then v heads: the layout bench.py and the fused-N GEMM use)."""
    d = sh.head_dim
    q = wq.unflatten(0, (sh.num_heads, d))[sh.q_range].flatten(0, 1)
    k = wk.unflatten(0, (sh.num_kv_heads, d))[sh.kv_range].flatten(0, 1)
    v = wv.unflatten(0, (sh.num_kv_heads, d))[sh.kv_range].flatten(0, 1)
    return torch.cat([q, k, v], dim=0).contiguous()


def shard_o_weight(wo: torch.Tensor, sh: HeadShard) -> torch.Tensor:
    """wo [hidden, H*D] -> this rank's [hidden, h*D] column block (row-parallel linear)."""
    d = sh.head_dim
    return wo.unflatten(1, (sh.num_heads, d))[:, sh.q_range].flatten(1, 2).contiguous()


def shard_kv_cache(kv_cache: torch.Tensor, sh: HeadShard) -> torch.Tensor:
    """[2, NB, bs, Hkv, D] -> [2, NB, bs, Hkv/N, D] (a copy: each rank allocates its own cache)."""
    return kv_cache[:, :, :, sh.kv_range].contiguous()


def all_reduce_sum(x: torch.Tensor) -> torch.Tensor:
    """models/src/multi_gpu.rs:141-179 — in place, whatever backend the process group was created with."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x)
    return x
