"""TEST INFRASTRUCTURE — CPU restatement of `ModelWorker::prepare_input_tensors`
(reference backends/vllm/src/worker.rs:224-460) in plain Python, step by step as the Rust does it
(ten growing vectors, then padding and cumulative sums).  Only tests import this; the product path
is atoma_infer_b200/csrc/input_prep.cpp.

Parity pinning: the reference has NO test of this function (grep `prepare_input_tensors`: one call
site, worker.rs:164) — "parity unpinned" apart from the one known answer the reference itself
states in a comment (worker.rs:380-384: prompt of 10 tokens, sliding window 8, block size 4 ->
slot mapping [-1, -1, 2, 3, 4, 5, 6, 7, 0, 1]), which tests/test_input_prep.py replays.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

PAD_SLOT_ID = -1  # worker.rs:13


@dataclass
class Seq:
    token_ids: List[int]
    num_computed_tokens: int
    token_chunk_size: int
    block_table: Optional[List[int]]   # None: the group's block_tables map is empty (profiling)
    is_prompt: bool


class EmptyPrompt(Exception):
    pass


def prepare_input_tensors(seqs: List[Seq], block_size: int, sliding_window: Optional[int],
                          enable_chunked_prefill: bool) -> dict:
    input_tokens, input_positions, slot_mapping = [], [], []
    sequence_lengths, prefill_sequence_lengths, decode_sequence_lengths = [], [], []
    context_lengths, query_lengths, block_tables = [], [], []
    num_prefills = num_prefill_tokens = num_decode_tokens = 0
    for s in seqs:
        length = len(s.token_ids)
        context_length = s.num_computed_tokens if s.is_prompt else length - 1          # :247-256
        sequence_length = min(length, context_length + s.token_chunk_size)             # :259-261
        if s.is_prompt:                                                                  # :264-276
            tokens = s.token_ids[context_length:sequence_length]
        else:
            if length == 0:
                raise EmptyPrompt("Empty prompts should not be received in `ModelWorker`")
            tokens = [s.token_ids[-1]]
        query_length = sequence_length - context_length if s.is_prompt else 1            # :279-283
        sliding_sequence_length = sequence_length                                        # :291-305
        if sliding_window is not None and not s.is_prompt:
            sliding_sequence_length = min(sliding_window, sequence_length)
        if enable_chunked_prefill or not s.is_prompt:                                    # :308-331
            assert s.block_table is not None, "Block table should be allocated for sequence on decoding phase"
            block_table = list(s.block_table)
            if sliding_window is not None:
                sw_block_num = (sliding_window + block_size - 1) // block_size
                start = max(0, len(block_table) - sw_block_num)
                block_table = block_table[start:]
        else:
            block_table = []
        block_tables.append(block_table)                                                 # :333-340
        sequence_lengths.append(sliding_sequence_length)
        context_lengths.append(context_length)
        query_lengths.append(query_length)
        input_tokens.extend(tokens)
        input_positions.extend(range(context_length, sequence_length))
        if s.is_prompt:                                                                  # :344-363
            num_prefills += 1
            num_prefill_tokens += len(tokens)
            prefill_sequence_lengths.append(sequence_length)
        else:
            num_decode_tokens += query_length
            decode_sequence_lengths.append(sliding_sequence_length)
        if s.block_table is None:                                                        # :365-371
            slot_mapping.extend([PAD_SLOT_ID] * sequence_length)
            continue
        start_index = max(0, query_length - sliding_window) if sliding_window is not None else 0  # :385-390
        for i in range(context_length, sequence_length):                                 # :392-401
            if i < start_index:
                slot_mapping.append(PAD_SLOT_ID)
            else:
                slot_mapping.append(s.block_table[i // block_size] * block_size + i % block_size)
    max_block_table_len = max(len(bt) for bt in block_tables)                           # :410 (panics when empty)
    padded = [bt + [0] * (max_block_table_len - len(bt)) for bt in block_tables]         # :411-412, :670-683

    def start_loc(xs):                                                                   # :414-434
        out, acc = [0], 0
        for x in xs:
            acc += x
            out.append(acc)
        return out

    return dict(input_tokens=input_tokens, input_positions=input_positions, slot_mapping=slot_mapping,
                sequence_lengths=sequence_lengths, context_lengths=context_lengths,
                sequence_start_loc=start_loc(sequence_lengths), query_start_loc=start_loc(query_lengths),
                block_tables=padded, max_block_table_len=max_block_table_len,
                num_prefills=num_prefills, num_prefill_tokens=num_prefill_tokens, num_decode_tokens=num_decode_tokens,
                max_query_len=max(query_lengths, default=0),
                max_prefill_seq_len=max(prefill_sequence_lengths, default=0),
                max_decode_seq_len=max(decode_sequence_lengths, default=0))
