"""Host-side mirror of `models::llama` for the decode step: configuration,
rotary tables, per-rank weight layout and the decode-step executor.

What the reference assembles per step from candle calls (models/src/llama.rs:253-314, 362-366,
398-410, 456-478; tensor parallel: llama_nccl.rs:97-139, multi_gpu.rs:8-57, 141-179) is ONE C call
here, `atoma_llama_decode_step` (csrc/llama_step.cu): 6 kernel launches per layer on one stream, no
host synchronisation, captured once per batch shape in a CUDA graph by `DecodeStep.capture()` and
replayed after each `InputStager.refresh()` (atoma_infer_b200/worker.py).

Weight layout (per tensor-parallel rank; built once at load time, as `load_multi`,
multi_gpu.rs:26-35, concatenates shards in the reference):
  wqkv      [(h + 2 hkv) D, hidden]  = cat(Wq[rank's heads], Wk[rank's kv heads], Wv[...])
  wo        [hidden, h D]            = Wo[:, rank's heads]
  w_gate_up [2 inter/N, hidden]      = cat(Wgate[rank's rows], Wup[rank's rows])
  w_down    [hidden, inter/N]        = Wdown[:, rank's columns]
  embed, lm_head, norms: replicated (llama_nccl.rs:320-321).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib
from ._lib import DecodeBatch, LlamaLayer, LlamaModel
from .ops import AtomaError
from .tp import HeadShard, shard_o_weight, shard_qkv_weight


@dataclass(frozen=True)
class Llama3RopeScaling:
    """models/src/llama.rs:75-90 (`Llama3RopeConfig`)"""
    factor: float = 8.0
    low_freq_factor: float = 1.0
    high_freq_factor: float = 4.0
    original_max_position_embeddings: int = 8192


@dataclass(frozen=True)
class LlamaConfig:
    """models/src/llama.rs:84-98 (`Config`)"""
    hidden_size: int
    intermediate_size: int
    vocab_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_position_embeddings: int = 8192
    rope_scaling: Optional[Llama3RopeScaling] = None
    tie_word_embeddings: bool = False

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @staticmethod
    def llama31_8b(**kw) -> "LlamaConfig":
        return LlamaConfig(4096, 14336, 128256, 32, 32, 8, **kw)

    @staticmethod
    def llama32_1b(**kw) -> "LlamaConfig":
        return LlamaConfig(2048, 8192, 128256, 16, 32, 8, tie_word_embeddings=True, **kw)

    @staticmethod
    def llama31_70b(**kw) -> "LlamaConfig":
        return LlamaConfig(8192, 28672, 128256, 80, 64, 8, **kw)


def rope_tables(cfg: LlamaConfig, dtype: torch.dtype, device, max_pos: Optional[int] = None):
    """cos/sin [max_pos, D/2] exactly as `Cache::new` builds them (models/src/llama.rs:145-199): f32
    inverse frequencies (Llama-3 scaling when configured), f32 outer product with the positions,
    cos/sin in f32, cast to the model dtype."""
    d = cfg.head_dim
    inv = 1.0 / (torch.tensor(cfg.rope_theta, dtype=torch.float32) **
                 (torch.arange(0, d, 2, dtype=torch.float32) / d))
    rs = cfg.rope_scaling
    if rs is not None:
        low_wl = rs.original_max_position_embeddings / rs.low_freq_factor
        high_wl = rs.original_max_position_embeddings / rs.high_freq_factor
        wl = 2.0 * math.pi / inv
        smooth = (rs.original_max_position_embeddings / wl - rs.low_freq_factor) / (rs.high_freq_factor - rs.low_freq_factor)
        mid = (1.0 - smooth) * inv / rs.factor + smooth * inv
        inv = torch.where(wl < high_wl, inv, torch.where(wl > low_wl, inv / rs.factor, mid))
    n = max_pos or cfg.max_position_embeddings
    ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().to(dtype).to(device), ang.sin().to(dtype).to(device)


class LlamaShard:
    """One tensor-parallel rank's weights in the layout the kernels stream (see the module docstring).
    `random_init` draws the FULL matrices from a seeded generator layer by layer and keeps the rank's
    shard, so that N ranks together hold exactly the tp=1 model (used by the parity tests and the
    bench: there is no network for checkpoints)."""

    def __init__(self, cfg: LlamaConfig, rank: int, world: int, dtype: torch.dtype, device):
        self.cfg, self.rank, self.world, self.dtype, self.device = cfg, rank, world, dtype, torch.device(device)
        self.shard = HeadShard(rank, world, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim)
        if cfg.intermediate_size % world:
            raise ValueError(f"tensor parallel degree {world} must divide intermediate_size {cfg.intermediate_size}")
        self.inter = cfg.intermediate_size // world
        self.layers: List[dict] = []
        self.embed = self.lm_head = self.norm_f = None
        self.cos = self.sin = None

    @classmethod
    def random_init(cls, cfg: LlamaConfig, rank: int = 0, world: int = 1, dtype=torch.bfloat16, device="cuda",
                    seed: int = 1234, num_layers: Optional[int] = None, with_head: bool = True,
                    max_pos: Optional[int] = None, gen_device=None) -> "LlamaShard":
        self = cls(cfg, rank, world, dtype, device)
        gdev = torch.device(gen_device) if gen_device is not None else self.device
        g = torch.Generator(device=gdev).manual_seed(seed)
        hid, D = cfg.hidden_size, cfg.head_dim
        sh = self.shard

        def rnd(rows, cols, scale):
            w = torch.randn((rows, cols), generator=g, device=gdev, dtype=torch.float32) * scale
            return w.to(dtype)

        def norm_w():
            return (1.0 + 0.1 * torch.randn((hid,), generator=g, device=gdev, dtype=torch.float32)).to(dtype)

        # the replicated tensors first: a model built with fewer layers is a PREFIX of the full one (same
        # seed -> same embedding, head and first layers), which the bench's tensor-parallel parity check uses
        self.norm_f = norm_w().to(self.device)
        if with_head:
            self.embed = rnd(cfg.vocab_size, hid, 1.0).to(self.device)
            self.lm_head = self.embed if cfg.tie_word_embeddings else rnd(cfg.vocab_size, hid, hid ** -0.5).to(self.device)
        L = num_layers if num_layers is not None else cfg.num_hidden_layers
        isl = slice(rank * self.inter, (rank + 1) * self.inter)
        for _ in range(L):
            wq = rnd(cfg.num_attention_heads * D, hid, hid ** -0.5)
            wk = rnd(cfg.num_key_value_heads * D, hid, hid ** -0.5)
            wv = rnd(cfg.num_key_value_heads * D, hid, hid ** -0.5)
            wo = rnd(hid, cfg.num_attention_heads * D, (cfg.num_attention_heads * D) ** -0.5)
            wg = rnd(cfg.intermediate_size, hid, hid ** -0.5)
            wu = rnd(cfg.intermediate_size, hid, hid ** -0.5)
            wd = rnd(hid, cfg.intermediate_size, cfg.intermediate_size ** -0.5)
            self.layers.append({
                "wqkv": shard_qkv_weight(wq, wk, wv, sh).to(self.device),
                "wo": shard_o_weight(wo, sh).to(self.device),
                "w_gate_up": torch.cat([wg[isl], wu[isl]], dim=0).contiguous().to(self.device),
                "w_down": wd[:, isl].contiguous().to(self.device),
                "rms_attn": norm_w().to(self.device),
                "rms_mlp": norm_w().to(self.device),
            })
            del wq, wk, wv, wo, wg, wu, wd
        self.cos, self.sin = rope_tables(cfg, dtype, self.device, max_pos)
        return self

    def prefix(self, num_layers: int, with_head: bool = True) -> "LlamaShard":
        """The first `num_layers` layers of this shard (tensors shared, nothing copied), with or without the
        embedding / lm_head: a smaller model whose step is a prefix of this one's."""
        sub = LlamaShard(self.cfg, self.rank, self.world, self.dtype, self.device)
        sub.layers = self.layers[:num_layers]
        sub.norm_f, sub.cos, sub.sin = self.norm_f, self.cos, self.sin
        sub.embed = self.embed
        sub.lm_head = self.lm_head if with_head else None
        return sub

    def weight_bytes(self) -> int:
        n = sum(t.numel() * t.element_size() for ly in self.layers for t in ly.values())
        for t in (self.norm_f, self.lm_head):
            if t is not None:
                n += t.numel() * t.element_size()
        return n


class DecodeStep:
    """`Llama::forward` for a decode-only batch (models/src/llama.rs:456-478) + greedy sampling
    (model_executor.rs:227-262) as one stream-ordered call.  Owns the scratch and output buffers;
    the step's inputs are the tensors handed to `bind()` (normally views of the InputStager's device
    image, whose addresses are stable across steps).

        step = DecodeStep(shard, kv_caches, max_batch=32, p2p=comm)
        step.bind(tokens, positions, slot_mapping, block_tables, sequence_lengths)
        step.capture()        # warm-up (allocates library workspaces) + CUDA graph of the step
        step.run()            # replay (or an eager launch when not captured)
        step.next_tokens      # int32 [T] on the device
    """

    def __init__(self, shard: LlamaShard, kv_caches: List[torch.Tensor], max_batch: int, p2p=None,
                 want_logits: bool = False, want_logits_f32: bool = False, want_hidden: bool = False,
                 no_comm: bool = False, vocab_parallel_head: Optional[bool] = None):
        cfg = shard.cfg
        if max_batch < 1 or max_batch > 128:
            raise AtomaError("the decode-step executor takes 1..128 sequences per step")
        if len(kv_caches) != len(shard.layers):
            raise AtomaError(f"{len(shard.layers)} layers need {len(shard.layers)} kv caches, got {len(kv_caches)}")
        if shard.world > 1 and p2p is None and not no_comm:
            # no_comm: diagnostics only — one rank's shard of the step without the all-reduces (timing of the
            # per-rank work; the numbers it produces are partial sums)
            raise AtomaError("tensor parallel degree > 1 needs a P2PComm (atoma_infer_b200.p2p)")
        self.shard, self.cfg, self.p2p = shard, cfg, p2p
        self.device, self.dtype = shard.device, shard.dtype
        self.kv_caches = kv_caches
        hk, D = shard.shard.kv_heads, cfg.head_dim
        self._layers = (LlamaLayer * len(shard.layers))()
        for i, (ly, kv) in enumerate(zip(shard.layers, kv_caches)):
            if kv.dim() != 5 or kv.shape[0] != 2 or tuple(kv.shape[3:]) != (hk, D) or not kv.is_contiguous():
                raise AtomaError(f"kv cache {i} must be a contiguous [2, num_blocks, block_size, {hk}, {D}] tensor")
            for name in ("wqkv", "wo", "w_gate_up", "w_down", "rms_attn", "rms_mlp"):
                setattr(self._layers[i], name, ly[name].data_ptr())
            self._layers[i].key_cache = kv[0].data_ptr()
            self._layers[i].value_cache = kv[1].data_ptr()
        kv0 = kv_caches[0]
        m = LlamaModel()
        m.hidden, m.inter, m.num_layers = cfg.hidden_size, shard.inter, len(shard.layers)
        m.num_heads, m.num_kv_heads, m.head_dim, m.vocab = shard.shard.heads, hk, D, cfg.vocab_size
        m.block_size, m.block_stride = kv0.shape[2], kv0.stride(1)
        m.rot_stride = shard.cos.stride(0)
        m.rms_eps, m.softmax_scale = cfg.rms_norm_eps, 1.0 / math.sqrt(D)
        m.dtype = 1 if shard.dtype == torch.bfloat16 else 0
        m.embed = shard.embed.data_ptr() if shard.embed is not None else None
        m.norm_f = shard.norm_f.data_ptr()
        m.lm_head = shard.lm_head.data_ptr() if shard.lm_head is not None else None
        m.cos, m.sin = shard.cos.data_ptr(), shard.sin.data_ptr()
        m.layers = C.cast(self._layers, C.POINTER(LlamaLayer))
        m.tp_rank, m.tp_world = (shard.rank, shard.world) if p2p is not None else (0, 1)
        if p2p is not None:
            m.p2p_regions = C.cast(p2p._regions, C.POINTER(C.c_void_p))
            m.p2p_slot_bytes = p2p.slot_bytes
        # Vocab-parallel head (tensor parallel, sampled tokens only): this rank multiplies by its [vocab / world, hidden]
        # slice of lm_head and the ranks exchange their argmax candidates (atoma_sample_greedy_tp).  The reference
        # replicates lm_head on every rank (llama_nccl.rs:320-321); the sampled tokens are identical.
        can_vp = (p2p is not None and shard.lm_head is not None and not want_logits and not want_logits_f32 and
                  cfg.vocab_size % (8 * shard.world) == 0 and p2p.slot_bytes >= 32768)
        if vocab_parallel_head and not can_vp:
            raise AtomaError("a vocab-parallel head needs tensor parallelism, vocab % (8 * world) == 0, p2p slots of >= 32 KiB "
                             "and returns sampled tokens only (no logits)")
        self.vocab_parallel_head = can_vp if vocab_parallel_head is None else bool(vocab_parallel_head)
        if self.vocab_parallel_head:
            rows = cfg.vocab_size // shard.world
            m.lm_head_rows, m.lm_head_row0 = rows, rows * shard.rank
            m.lm_head = shard.lm_head.data_ptr() + rows * shard.rank * shard.lm_head.stride(0) * shard.lm_head.element_size()
        self._model = m
        self.max_batch = max_batch
        L = _lib.lib()
        nbytes = L.atoma_llama_decode_scratch_bytes(C.byref(m), max_batch)
        self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.next_tokens = torch.zeros((max_batch,), dtype=torch.int32, device=self.device)
        self.next_logits = torch.zeros((max_batch,), dtype=torch.float32, device=self.device)
        self.logits = (torch.empty((max_batch, cfg.vocab_size), dtype=self.dtype, device=self.device)
                       if want_logits and shard.lm_head is not None else None)
        self.logits_f32 = (torch.empty((max_batch, cfg.vocab_size), dtype=torch.float32, device=self.device)
                           if want_logits_f32 and shard.lm_head is not None else None)
        self.hidden_out = (torch.empty((max_batch, cfg.hidden_size), dtype=self.dtype, device=self.device)
                           if want_hidden or shard.lm_head is None else None)
        self._batch: Optional[DecodeBatch] = None
        self._bound = None
        self._graph = None
        self._graph_stream = None

    # ---- inputs ----
    def bind(self, tokens, positions, slot_mapping, block_tables, sequence_lengths, hidden_in=None) -> None:
        """tokens int32 [T] (u32 bit patterns; ignored when the shard has no embedding and `hidden_in`
        [T, hidden] is given), positions i64 [T], slot_mapping i64 [T], block_tables int32
        [T, max_blocks], sequence_lengths int32 [T].  The tensors are read at every run()."""
        T = positions.numel()
        if T < 1 or T > self.max_batch:
            raise AtomaError(f"batch of {T} sequences does not fit max_batch={self.max_batch}")
        if positions.dtype != torch.int64 or slot_mapping.dtype != torch.int64 or slot_mapping.numel() != T:
            raise AtomaError("input_positions and slot_mapping must be i64 tensors with one entry per token")
        if block_tables.dtype != torch.int32 or block_tables.dim() != 2 or block_tables.shape[0] != T or \
                not block_tables.is_contiguous():
            raise AtomaError("block_tables must be a contiguous u32 [num_sequences, max_blocks] tensor")
        if sequence_lengths.dtype != torch.int32 or sequence_lengths.numel() != T:
            raise AtomaError("sequence_lengths must be a u32 tensor with one entry per sequence")
        if self.shard.embed is None and hidden_in is None:
            raise AtomaError("a shard without an embedding table needs hidden_in")
        b = DecodeBatch()
        b.num_tokens = T
        b.tokens = tokens.data_ptr() if tokens is not None else None
        b.hidden_in = hidden_in.data_ptr() if hidden_in is not None else None
        b.positions, b.slot_mapping = positions.data_ptr(), slot_mapping.data_ptr()
        b.block_tables, b.sequence_lengths = block_tables.data_ptr(), sequence_lengths.data_ptr()
        b.max_blocks = block_tables.shape[1]
        b.scratch, b.scratch_bytes = self._scratch.data_ptr(), self._scratch.numel()
        p = lambda t: t.data_ptr() if t is not None else None
        b.hidden_out, b.logits, b.logits_f32 = p(self.hidden_out), p(self.logits), p(self.logits_f32)
        if self.shard.lm_head is not None:
            b.next_tokens, b.next_logits = self.next_tokens.data_ptr(), self.next_logits.data_ptr()
        self._batch = b
        self._bound = (tokens, positions, slot_mapping, block_tables, sequence_lengths, hidden_in)  # keep alive
        self._graph = None
        self.num_tokens = T

    # ---- execution ----
    def launch(self) -> None:
        """Enqueue the step's kernels on the current stream (eager)."""
        if self._batch is None:
            raise AtomaError("bind() the step's input tensors first")
        L = _lib.lib()
        L.atoma_llama_decode_step(C.byref(self._model), C.byref(self._batch), torch.cuda.current_stream().cuda_stream)
        if L.atoma_last_status() != 0:
            raise AtomaError(L.atoma_last_error().decode())

    def capture(self) -> None:
        """One eager step on the capture stream (the library allocates its per-stream workspaces on
        first use and refuses to allocate during capture), then the CUDA graph of the step."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.launch()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.launch()
        torch.cuda.synchronize()
        self._graph, self._graph_stream = g, side

    def run(self) -> None:
        if self._graph is not None:
            self._graph.replay()
        else:
            self.launch()

    def launches_per_step(self) -> int:
        """Kernels one step launches (counted by the library on an eager launch)."""
        L = _lib.lib()
        L.atoma_reset_launch_count()
        self.launch()
        return int(L.atoma_launch_count())

    def check(self) -> None:
        """Synchronises the stream the step runs on and raises if a bounded in-kernel wait ever expired on it
        (`atoma_llama_step_status`: the chained MLP's grid barrier, the fused norm tail) or, under tensor
        parallelism, if a peer failed to arrive at an all-reduce.  Call it where the serving loop already
        synchronises (after copying the sampled tokens back)."""
        L = _lib.lib()
        cur = torch.cuda.current_stream(self.device)
        cur.synchronize()                                  # replays run on the current stream ...
        streams = {cur.cuda_stream}
        if self._graph_stream is not None:
            streams.add(self._graph_stream.cuda_stream)    # ... over the barrier workspace of the stream they were captured on
        for s in streams:
            rc = L.atoma_llama_step_status(s)
            if rc != 0:
                raise AtomaError(L.atoma_last_error().decode() or f"atoma_llama_step_status = {rc}")
        if self.p2p is not None:
            self.p2p.check()

    def release(self) -> None:
        self._graph = None
