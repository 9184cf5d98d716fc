"""Step input preparation: `atoma_prepare_input` (C++) against the Python restatement
of `ModelWorker::prepare_input_tensors` (oracle/worker_oracle.py, worker.rs:224-460) on seeded
schedules, the one known answer the reference states (worker.rs:380-384), its error texts, and the
packed-buffer contract (one buffer, 16-byte aligned sections, views == reference tensors)."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import worker
from atoma_infer_b200.ops import AtomaError
from oracle import worker_oracle as orc


def _mk(rng, n_prompt, n_decode, block_size, chunked, with_tables=True, max_len=200):
    seqs_o, seqs_p = [], []
    for i in range(n_prompt + n_decode):
        is_prompt = i < n_prompt          # the scheduler puts prefills first (flash_attention.rs:322-469)
        length = int(rng.integers(1, max_len))
        toks = rng.integers(0, 128256, size=length).astype(np.uint32)
        if is_prompt:
            computed = int(rng.integers(0, length)) if chunked else 0
            chunk = int(rng.integers(1, length - computed + 1)) if chunked else length
        else:
            computed, chunk = 0, 1
        nblk = (length + block_size - 1) // block_size
        bt = rng.permutation(4096)[:nblk].astype(np.uint32) if with_tables else None
        seqs_o.append(orc.Seq(toks.tolist(), computed, chunk, None if bt is None else bt.tolist(), is_prompt))
        seqs_p.append(worker.SequenceInput(toks, computed, chunk, bt, is_prompt))
    return seqs_o, seqs_p


def _check(mi: worker.ModelInput, want: dict):
    lay = mi.layout
    as_list = lambda t: t.cpu().numpy().astype(np.int64).tolist()
    assert as_list(mi.input_tokens_tensor[0].view(torch.int32)) == [int(np.int32(np.uint32(x))) for x in want["input_tokens"]]
    assert as_list(mi.input_positions[0]) == want["input_positions"]
    md = mi.attention_metadata
    assert as_list(md.slot_mapping) == want["slot_mapping"]
    assert as_list(md.context_lengths) == want["context_lengths"]
    assert as_list(mi.cu_query_lengths) == want["query_start_loc"]
    for k in ("num_prefills", "num_prefill_tokens", "num_decode_tokens", "max_query_len", "max_prefill_seq_len",
              "max_decode_seq_len", "max_block_table_len"):
        assert getattr(lay, k) == want[k], k
    P = want["num_prefills"]
    if want["num_prefill_tokens"] > 0:
        pm = md.prefill_metadata
        assert as_list(pm.sequence_start_locations) == want["sequence_start_loc"][: P + 1]
        assert as_list(pm.sequence_lengths) == want["sequence_lengths"][:P]
        assert pm.max_prefill_sequence_length == want["max_prefill_seq_len"]
    else:
        assert md.prefill_metadata is None
    if want["num_decode_tokens"] > 0:
        dm = md.decoding_metadata
        assert as_list(dm.sequence_lengths) == want["sequence_lengths"][P:]
        assert as_list(dm.block_tables) == want["block_tables"]     # prefix_caching = false: the whole table (flash_attention.rs:108-112)
        assert dm.max_decoding_sequence_length == want["max_decode_seq_len"]
    else:
        assert md.decoding_metadata is None
    # packed-buffer contract
    offs = [lay.input_tokens_off, lay.input_positions_off, lay.slot_mapping_off, lay.sequence_lengths_off,
            lay.context_lengths_off, lay.sequence_start_loc_off, lay.query_start_loc_off, lay.block_tables_off]
    assert all(o % 16 == 0 for o in offs) and offs == sorted(offs) and lay.total_bytes >= offs[-1]


@pytest.mark.parametrize("n_prompt,n_decode,block_size,window,chunked", [
    (0, 32, 16, None, False),      # BASELINE config 2 shape: decode only
    (3, 0, 16, None, False),       # prefill only, no block tables in the tensor (worker.rs:329-331)
    (2, 5, 16, None, True),        # chunked prefill mixed with decodes
    (1, 4, 8, 24, False),          # sliding window: decode lengths capped, tables trimmed
    (2, 3, 4, 8, True),            # sliding window + chunked prefill: masked prompt slots
    (0, 1, 32, None, False),
])
def test_matches_the_restated_reference(n_prompt, n_decode, block_size, window, chunked):
    rng = np.random.default_rng(n_prompt * 100 + n_decode * 10 + block_size)
    st = worker.InputStager(None, block_size, window, chunked)
    for trial in range(5):
        so, sp = _mk(rng, n_prompt, n_decode, block_size, chunked)
        _check(st.prepare_input_tensors(sp), orc.prepare_input_tensors(so, block_size, window, chunked))


def test_known_answer_from_the_reference_comment():
    """worker.rs:380-384: prompt of 10 tokens, sliding window 8, block size 4 ->
    [-1, -1, 2, 3, 4, 5, 6, 7, 0, 1] (the block table wraps around: [0, 1, 0])."""
    toks = np.arange(10, dtype=np.uint32)
    s = worker.SequenceInput(toks, 0, 10, np.array([0, 1, 0], dtype=np.uint32), True)
    mi = worker.InputStager(None, 4, 8, False).prepare_input_tensors([s])
    assert mi.attention_metadata.slot_mapping.tolist() == [-1, -1, 2, 3, 4, 5, 6, 7, 0, 1]
    want = orc.prepare_input_tensors([orc.Seq(toks.tolist(), 0, 10, [0, 1, 0], True)], 4, 8, False)
    assert want["slot_mapping"] == [-1, -1, 2, 3, 4, 5, 6, 7, 0, 1]


def test_profiling_run_without_block_tables_pads_slots():
    """worker.rs:365-371: no block tables -> `sequence_length` pad slots per sequence."""
    rng = np.random.default_rng(5)
    so, sp = _mk(rng, 3, 0, 16, False, with_tables=False)
    mi = worker.InputStager(None, 16, None, False).prepare_input_tensors(sp)
    want = orc.prepare_input_tensors(so, 16, None, False)
    assert mi.attention_metadata.slot_mapping.tolist() == want["slot_mapping"]
    assert set(want["slot_mapping"]) == {-1}


def test_error_texts():
    st = worker.InputStager(None, 16, None, False)
    with pytest.raises(AtomaError, match="Empty prompts should not be received in `ModelWorker`"):
        st.prepare_input_tensors([worker.SequenceInput(np.zeros(0, np.uint32), 0, 1, np.zeros(1, np.uint32), False)])
    with pytest.raises(AtomaError, match="Block table should be allocated for sequence on decoding phase"):
        st.prepare_input_tensors([worker.SequenceInput(np.arange(5, dtype=np.uint32), 0, 1, None, False)])
    with pytest.raises(AtomaError, match="no sequences"):
        st.prepare_input_tensors([])
    with pytest.raises(AtomaError, match="needs block 1 of a block table with 1 entries"):
        st.prepare_input_tensors([worker.SequenceInput(np.arange(20, dtype=np.uint32), 0, 1, np.zeros(1, np.uint32), False)])


def test_decode_step_addresses_are_stable_for_graph_replay():
    """A fixed-shape decode step lands at the same offsets every time (what a captured graph needs)."""
    rng = np.random.default_rng(9)
    st = worker.InputStager(None, 16, None, False)
    bts = [rng.permutation(8192)[:256].astype(np.uint32) for _ in range(32)]
    ptrs = []
    for step in range(3):
        seqs = [worker.SequenceInput(rng.integers(0, 1000, size=4000 + step).astype(np.uint32), 0, 1, bts[b], False)
                for b in range(32)]
        mi = st.prepare_input_tensors(seqs)
        md = mi.attention_metadata
        ptrs.append((md.slot_mapping.data_ptr(), md.decoding_metadata.block_tables.data_ptr(),
                     md.decoding_metadata.sequence_lengths.data_ptr(), mi.input_positions.data_ptr()))
        assert md.decoding_metadata.sequence_lengths.tolist() == [4000 + step] * 32
        last = 4000 + step - 1
        assert md.slot_mapping.tolist() == [int(bts[b][last // 16]) * 16 + last % 16 for b in range(32)]
    assert ptrs[0] == ptrs[1] == ptrs[2]


def test_selected_token_indices():
    assert worker.compute_selected_token_indices(torch.tensor([0, 4, 10, 11])).tolist() == [3, 9, 10]


@pytest.mark.gpu
def test_one_copy_ships_the_step_to_the_gpu():
    rng = np.random.default_rng(3)
    so, sp = _mk(rng, 2, 6, 16, True)
    st = worker.InputStager("cuda:0", 16, None, True)
    mi = st.prepare_input_tensors(sp)
    torch.cuda.synchronize()
    assert mi.attention_metadata.slot_mapping.is_cuda
    _check(mi, orc.prepare_input_tensors(so, 16, None, True))


def test_refresh_never_reallocates_the_staging_image():
    """Round-1 advisor (medium): refresh() used to swap the host / device images when a step outgrew them,
    leaving views handed out earlier (and graphs captured over them) pointing at a dead image.  Now only
    prepare_input_tensors() may grow (new views, `generation` bumped); refresh() raises."""
    from atoma_infer_b200.ops import AtomaError
    InputStager, SequenceInput = worker.InputStager, worker.SequenceInput
    st = InputStager(None, 16, initial_bytes=4096)
    small = [SequenceInput(np.arange(20, dtype=np.uint32), 0, 1, np.arange(2, dtype=np.uint32), False)]
    mi = st.prepare_input_tensors(small)
    gen = st.generation
    assert st.refresh(st.marshal(small)).total_bytes == mi.layout.total_bytes and st.generation == gen
    big = [SequenceInput(np.arange(5000, dtype=np.uint32), 0, 1, np.arange(400, dtype=np.uint32), False) for _ in range(8)]
    with pytest.raises(AtomaError, match="views taken earlier"):
        st.refresh(st.marshal(big))
    assert st.generation == gen
    mi2 = st.prepare_input_tensors(big)          # allowed to grow: hands out fresh views
    assert st.generation == gen + 1 and mi2.attention_metadata.decoding_metadata.block_tables.shape == (8, 400)


def test_random_sweep_against_the_restated_reference():
    """300 random schedules over block sizes, sliding windows (smaller and larger than a block, than the prompts),
    chunked prefill on / off, prefill-only / decode-only / mixed batches: every tensor and scalar of the packed
    buffer equals the step-by-step restatement of worker.rs:224-460."""
    rng = np.random.default_rng(2024)
    for trial in range(300):
        block_size = int(rng.choice([1, 4, 8, 16, 32]))
        window = [None, 1, 5, 8, 24, 64][int(rng.integers(0, 6))]
        chunked = bool(rng.integers(0, 2))
        n_prompt, n_decode = int(rng.integers(0, 4)), int(rng.integers(0, 6))
        if n_prompt + n_decode == 0:
            n_decode = 1
        so, sp = _mk(rng, n_prompt, n_decode, block_size, chunked, max_len=int(rng.choice([2, 9, 70, 300])))
        st = worker.InputStager(None, block_size, window, chunked)
        _check(st.prepare_input_tensors(sp), orc.prepare_input_tensors(so, block_size, window, chunked))
