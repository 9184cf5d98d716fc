"""Step input preparation: mirror of `ModelWorker::prepare_input_tensors` and `ModelInput`
(reference backends/vllm/src/worker.rs:224-460, :24-40), SURVEY §8 row f2.

The host logic is C++ (`atoma_prepare_input`, csrc/input_prep.cpp); this module marshals sequences
into it, owns the pinned staging buffer and its device image, and exposes the step's tensors as
VIEWS of that device image — one host-to-device copy per step where the reference issues one per
tensor plus one per sequence for the block-table padding, and two GPU cumsums.  Because the staging
buffers only grow, the device addresses of a fixed-shape decode step stay put and a captured CUDA
graph of the step can be replayed after every `prepare_input_tensors` call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence as Seq_

import numpy as np
import torch

from . import _lib
from ._lib import ModelInputLayout, Sequence
from .flash_attention import FlashAttentionMetadata
from .ops import AtomaError


@dataclass
class SequenceInput:
    """One sequence of a scheduled `SequenceGroupMetadata` (group fields repeated per sequence)."""
    token_ids: np.ndarray                 # uint32: prompt + generated tokens (`get_token_ids()`)
    num_computed_tokens: int = 0          # prompts only
    token_chunk_size: int = 1
    block_table: Optional[np.ndarray] = None   # uint32 physical block numbers; None = no block tables
    is_prompt: bool = False


@dataclass
class ModelInput:
    """worker.rs:24-40"""
    input_tokens_tensor: torch.Tensor     # u32 bits as int32 [1, T]
    input_positions: torch.Tensor         # i64 [1, T]
    num_decode_tokens: int
    num_prefills: int
    cu_query_lengths: torch.Tensor        # u32 bits as int32 [B + 1]
    attention_metadata: FlashAttentionMetadata
    layout: ModelInputLayout


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


class InputStager:
    """Pinned host buffer + device image, grow-only."""

    def __init__(self, device=None, block_size: int = 16, sliding_window: Optional[int] = None,
                 enable_chunked_prefill: bool = False, initial_bytes: int = 1 << 16):
        self.device = torch.device(device) if device is not None else None
        self.block_size, self.sliding_window = block_size, sliding_window
        self.enable_chunked_prefill = enable_chunked_prefill
        self._hosts = self._dev = None
        self._events = [None, None]
        self._turn = 0
        self.generation = 0          # bumped whenever the device image is reallocated (views taken earlier are dead)
        self._reserve(initial_bytes)

    @property
    def _host(self) -> torch.Tensor:
        return self._hosts[self._turn]

    def _reserve(self, nbytes: int) -> None:
        """Two pinned staging buffers (so that step k+1 can be marshalled while step k's copy is still queued)
        and one device image."""
        if self._hosts is not None and self._hosts[0].numel() >= nbytes:
            return
        cap = 1 << max(12, (int(nbytes) - 1).bit_length())
        cuda = self.device is not None and self.device.type == "cuda"
        if self._hosts is not None and cuda:
            torch.cuda.synchronize(self.device)      # queued copies still read the old buffers
        self._hosts = [torch.empty(cap, dtype=torch.uint8, pin_memory=cuda) for _ in range(2 if cuda else 1)]
        self._events = [None] * len(self._hosts)
        self._turn = 0
        self._dev = torch.empty(cap, dtype=torch.uint8, device=self.device) if self.device is not None else self._hosts[0]
        self.generation += 1

    def _marshal(self, seqs: Seq_[SequenceInput]):
        arr = (Sequence * len(seqs))()
        keep = []
        for i, s in enumerate(seqs):
            tok = _u32(s.token_ids)
            keep.append(tok)
            arr[i].token_ids = tok.ctypes.data_as(C.POINTER(C.c_uint32))
            arr[i].length = tok.size
            arr[i].num_computed_tokens = int(s.num_computed_tokens)
            arr[i].token_chunk_size = int(s.token_chunk_size)
            if s.block_table is not None:
                bt = _u32(s.block_table)
                # an empty table is still "allocated" (NULL means "the group has no block tables"):
                # point at a one-element dummy, num_blocks stays 0
                buf = bt if bt.size else np.zeros(1, dtype=np.uint32)
                keep.append(buf)
                arr[i].block_table = buf.ctypes.data_as(C.POINTER(C.c_uint32))
                arr[i].num_blocks = bt.size
            arr[i].is_prompt = 1 if s.is_prompt else 0
        return arr, keep

    def marshal(self, seqs: Seq_[SequenceInput]):
        """The ctypes image of `seqs` (pointers into the callers' arrays, which it keeps alive).  A
        Rust host hands its own vectors to the C entry directly; a Python caller that prepares many
        steps over the same sequence objects can marshal once and call `prepare_marshalled`."""
        return self._marshal(seqs)

    def prepare_input_tensors(self, seqs: Seq_[SequenceInput], prefix_caching: bool = False) -> ModelInput:
        """worker.rs:224-460; the tensors of the result are views of one device buffer."""
        return self.prepare_marshalled(self._marshal(seqs), prefix_caching)

    def prepare_marshalled(self, handle, prefix_caching: bool = False) -> ModelInput:
        return self._views(self._fill(handle, allow_grow=True), prefix_caching)

    def refresh(self, handle) -> ModelInputLayout:
        """Rebuilds the step's tensors in a pinned buffer and ships them with one copy; returns the
        layout only.  For steady-state steps whose tensor views were taken earlier (same shapes ->
        same offsets; the caller may compare `total_bytes` and the counts).  The device image is never
        reallocated here: a step that no longer fits raises (views and captured graphs over the old
        image would silently go stale) — call prepare_input_tensors again and re-capture."""
        return self._fill(handle, allow_grow=False)

    def _fill(self, handle, allow_grow: bool) -> ModelInputLayout:
        L = _lib.lib()
        arr, _keep = handle
        lay = ModelInputLayout()
        args = (arr, len(arr), self.block_size, int(self.sliding_window or 0), 1 if self.enable_chunked_prefill else 0)
        # The staging buffers alternate; before one is rewritten, the copy that last read it must have
        # executed (an asynchronous copy reads pinned memory when it RUNS, not when it is enqueued).
        self._turn = (self._turn + 1) % len(self._hosts)
        if self._events[self._turn] is not None:
            self._events[self._turn].synchronize()
        # one call computes the layout and fills the buffer; grow and retry if it did not fit
        err = L.atoma_prepare_input(*args, C.byref(lay), self._host.data_ptr(), self._host.numel())
        if err and lay.total_bytes > self._host.numel():
            if not allow_grow:
                raise AtomaError(f"step inputs need {int(lay.total_bytes)} bytes but the staging image holds "
                                 f"{self._host.numel()}: views taken earlier (and graphs captured over them) would go "
                                 "stale; call prepare_input_tensors() again and re-capture")
            self._reserve(lay.total_bytes)
            err = L.atoma_prepare_input(*args, C.byref(lay), self._host.data_ptr(), self._host.numel())
        if err:
            raise AtomaError(err.decode())
        if self._dev is not self._host:   # THE host-to-device copy of the step
            n = int(lay.total_bytes)
            self._dev[:n].copy_(self._host[:n], non_blocking=True)
            ev = self._events[self._turn] or torch.cuda.Event()
            ev.record()
            self._events[self._turn] = ev
        return lay

    def _views(self, lay: ModelInputLayout, prefix_caching: bool) -> ModelInput:
        d = self._dev

        def view(off, count, dtype):
            item = torch.empty((), dtype=dtype).element_size()
            return d[off: off + count * item].view(dtype)

        B = lay.num_sequences
        tokens = view(lay.input_tokens_off, lay.num_tokens, torch.int32)
        positions = view(lay.input_positions_off, lay.num_tokens, torch.int64)
        slots = view(lay.slot_mapping_off, lay.num_slots, torch.int64)
        seq_lens = view(lay.sequence_lengths_off, B, torch.int32)
        ctx_lens = view(lay.context_lengths_off, B, torch.int32)
        seq_start = view(lay.sequence_start_loc_off, B + 1, torch.int32)
        query_start = view(lay.query_start_loc_off, B + 1, torch.int32)
        tables = view(lay.block_tables_off, B * lay.max_block_table_len, torch.int32).reshape(B, lay.max_block_table_len)
        md = FlashAttentionMetadata.new(ctx_lens, slots, query_start, lay.num_prefill_tokens, lay.num_decode_tokens,
                                        lay.max_query_len, lay.max_decode_seq_len, lay.max_prefill_seq_len,
                                        lay.num_prefills, seq_start, seq_lens, tables, prefix_caching)
        return ModelInput(tokens.unsqueeze(0), positions.unsqueeze(0), lay.num_decode_tokens, lay.num_prefills,
                          query_start, md, lay)


def compute_selected_token_indices(cumulative_query_lengths: torch.Tensor) -> torch.Tensor:
    """worker.rs:688-697: index of the last token of every sequence in the flattened batch."""
    return cumulative_query_lengths[1:] - 1
