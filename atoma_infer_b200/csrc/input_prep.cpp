// Decode/prefill step input preparation: a C++ restatement of
// `ModelWorker::prepare_input_tensors` (reference backends/vllm/src/worker.rs:224-460) that lays
// every tensor of a step's `ModelInput` / `FlashAttentionMetadata` out in ONE host buffer, so that the
// step's metadata reaches the GPU with ONE host-to-device copy.  The reference builds ten
// host vectors, uploads each with its own `Tensor::new` (>= 8 H2D copies), pads the block tables
// with one `Tensor::from_vec` PER SEQUENCE + `cat` (worker.rs:670-683) and computes the two start-
// location vectors with GPU cumsums through f32 (worker.rs:414-434).  Pure host code: no CUDA calls.
#include "atoma_ops.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace {

thread_local char t_prep_msg[512];

const char *prep_fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_prep_msg, sizeof(t_prep_msg), fmt, ap);
    va_end(ap);
    return t_prep_msg;
}

constexpr int64_t kPadSlotId = -1;  // worker.rs:13

inline int64_t align16(int64_t x) { return (x + 15) / 16 * 16; }

struct SeqView {  // the per-sequence quantities of worker.rs:245-285
    int64_t context_length, sequence_length, query_length, sliding_sequence_length;
    int64_t bt_start, bt_len;  // slice of the block table that goes into the padded tensor
};

const char *derive(const atoma_sequence &s, int64_t idx, int64_t block_size, int64_t sliding_window,
                   bool chunked, SeqView *v) {
    if (s.length < 0 || s.num_computed_tokens < 0 || s.token_chunk_size < 0 || s.num_blocks < 0)
        return prep_fail("sequence %lld: negative length / computed tokens / chunk size / block count", (long long)idx);
    if (!s.is_prompt && s.length == 0)  // worker.rs:265-271 (`get_last_token_id().is_none()`)
        return prep_fail("Empty prompts should not be received in `ModelWorker`");
    if (!s.is_prompt && s.token_chunk_size < 1)  // a decode step schedules exactly one token per sequence
        return prep_fail("sequence %lld: a decoding sequence needs token_chunk_size >= 1", (long long)idx);
    v->context_length = s.is_prompt ? s.num_computed_tokens : s.length - 1;           // :247-256
    v->sequence_length = std::min(s.length, v->context_length + s.token_chunk_size);   // :259-261
    if (s.is_prompt && v->sequence_length < v->context_length)
        return prep_fail("sequence %lld: %lld computed tokens exceed its length %lld", (long long)idx,
                         (long long)v->context_length, (long long)s.length);
    v->query_length = s.is_prompt ? v->sequence_length - v->context_length : 1;        // :279-283
    v->sliding_sequence_length = v->sequence_length;                                   // :291-305
    if (sliding_window > 0 && !s.is_prompt) v->sliding_sequence_length = std::min(sliding_window, v->sequence_length);
    v->bt_start = 0;
    v->bt_len = 0;
    if (chunked || !s.is_prompt) {                                                     // :308-331
        if (!s.block_table)
            return prep_fail("Block table should be allocated for sequence on decoding phase");
        v->bt_len = s.num_blocks;
        if (sliding_window > 0) {
            const int64_t sw_blocks = (sliding_window + block_size - 1) / block_size;
            v->bt_start = std::max<int64_t>(0, s.num_blocks - sw_blocks);
            v->bt_len = s.num_blocks - v->bt_start;
        }
    }
    return nullptr;
}

}  // namespace

extern "C" const char *atoma_prepare_input(const atoma_sequence *seqs, int64_t num_sequences, int64_t block_size,
                                           int64_t sliding_window, int enable_chunked_prefill,
                                           atoma_model_input_layout *layout, void *buffer, int64_t buffer_bytes) {
    if (!layout) return prep_fail("atoma_prepare_input: layout is NULL");
    if (!seqs || num_sequences <= 0)  // worker.rs:410 unwraps the maximum over an empty list
        return prep_fail("atoma_prepare_input: no sequences to prepare");
    if (block_size <= 0) return prep_fail("atoma_prepare_input: block_size %lld", (long long)block_size);
    if (sliding_window < 0) sliding_window = 0;
    const bool chunked = enable_chunked_prefill != 0;

    // ---- pass 1: sizes and scalars ----
    atoma_model_input_layout L;
    std::memset(&L, 0, sizeof(L));
    L.num_sequences = num_sequences;
    for (int64_t i = 0; i < num_sequences; ++i) {
        SeqView v;
        if (const char *e = derive(seqs[i], i, block_size, sliding_window, chunked, &v)) return e;
        const int64_t ntok = seqs[i].is_prompt ? v.query_length : 1;
        L.num_tokens += ntok;
        // worker.rs:365-371: without block tables (memory profiling) `sequence_length` pad slots
        L.num_slots += seqs[i].block_table ? (v.sequence_length - v.context_length) : v.sequence_length;
        L.max_block_table_len = std::max(L.max_block_table_len, v.bt_len);
        L.max_query_len = std::max(L.max_query_len, v.query_length);
        if (seqs[i].is_prompt) {
            L.num_prefills += 1;
            L.num_prefill_tokens += ntok;
            L.max_prefill_seq_len = std::max(L.max_prefill_seq_len, v.sequence_length);
        } else {
            L.num_decode_tokens += 1;
            L.max_decode_seq_len = std::max(L.max_decode_seq_len, v.sliding_sequence_length);
        }
    }
    int64_t off = 0;
    auto take = [&off](int64_t bytes) { const int64_t o = off; off = align16(off + bytes); return o; };
    L.input_tokens_off = take(L.num_tokens * 4);
    L.input_positions_off = take(L.num_tokens * 8);
    L.slot_mapping_off = take(L.num_slots * 8);
    L.sequence_lengths_off = take(num_sequences * 4);
    L.context_lengths_off = take(num_sequences * 4);
    L.sequence_start_loc_off = take((num_sequences + 1) * 4);
    L.query_start_loc_off = take((num_sequences + 1) * 4);
    L.block_tables_off = take(num_sequences * L.max_block_table_len * 4);
    L.total_bytes = std::max<int64_t>(off, 16);
    *layout = L;
    if (!buffer) return nullptr;
    if (buffer_bytes < L.total_bytes)
        return prep_fail("atoma_prepare_input: buffer of %lld bytes, %lld needed", (long long)buffer_bytes,
                         (long long)L.total_bytes);

    // ---- pass 2: fill ----
    unsigned char *base = static_cast<unsigned char *>(buffer);
    uint32_t *tokens = reinterpret_cast<uint32_t *>(base + L.input_tokens_off);
    int64_t *positions = reinterpret_cast<int64_t *>(base + L.input_positions_off);
    int64_t *slots = reinterpret_cast<int64_t *>(base + L.slot_mapping_off);
    uint32_t *seq_lens = reinterpret_cast<uint32_t *>(base + L.sequence_lengths_off);
    uint32_t *ctx_lens = reinterpret_cast<uint32_t *>(base + L.context_lengths_off);
    uint32_t *seq_start = reinterpret_cast<uint32_t *>(base + L.sequence_start_loc_off);
    uint32_t *query_start = reinterpret_cast<uint32_t *>(base + L.query_start_loc_off);
    uint32_t *tables = reinterpret_cast<uint32_t *>(base + L.block_tables_off);
    int64_t t = 0, sl = 0;
    uint32_t seq_acc = 0, query_acc = 0;
    seq_start[0] = 0;    // worker.rs:414-422 (cumsum, exact here; the reference's goes through f32)
    query_start[0] = 0;  // worker.rs:428-434
    for (int64_t i = 0; i < num_sequences; ++i) {
        const atoma_sequence &s = seqs[i];
        SeqView v;
        derive(s, i, block_size, sliding_window, chunked, &v);
        if (s.is_prompt) {  // worker.rs:264-276
            if (v.query_length > 0 && !s.token_ids)
                return prep_fail("sequence %lld: token_ids is NULL", (long long)i);
            for (int64_t j = v.context_length; j < v.sequence_length; ++j) tokens[t + (j - v.context_length)] = s.token_ids[j];
        } else {
            if (!s.token_ids) return prep_fail("sequence %lld: token_ids is NULL", (long long)i);
            tokens[t] = s.token_ids[s.length - 1];
        }
        // worker.rs:340: positions context_length..sequence_length (one entry for a decode token)
        for (int64_t j = v.context_length; j < v.sequence_length; ++j) positions[t + (j - v.context_length)] = j;
        t += s.is_prompt ? v.query_length : 1;
        seq_lens[i] = static_cast<uint32_t>(v.sliding_sequence_length);   // :334
        ctx_lens[i] = static_cast<uint32_t>(v.context_length);            // :335
        seq_acc += static_cast<uint32_t>(v.sliding_sequence_length);
        query_acc += static_cast<uint32_t>(v.query_length);
        seq_start[i + 1] = seq_acc;
        query_start[i + 1] = query_acc;
        uint32_t *row = tables + i * L.max_block_table_len;               // :410-412, pad value 0
        for (int64_t j = 0; j < L.max_block_table_len; ++j) row[j] = j < v.bt_len ? s.block_table[v.bt_start + j] : 0u;
        if (!s.block_table) {                                             // :365-371
            for (int64_t j = 0; j < v.sequence_length; ++j) slots[sl++] = kPadSlotId;
            continue;
        }
        // :385-401 — the UNTRIMMED block table; tokens before `start_index` are masked
        const int64_t start_index = sliding_window > 0 ? std::max<int64_t>(0, v.query_length - sliding_window) : 0;
        for (int64_t j = v.context_length; j < v.sequence_length; ++j) {
            if (j < start_index) {
                slots[sl++] = kPadSlotId;
                continue;
            }
            const int64_t blk = j / block_size;
            if (blk >= s.num_blocks)  // the reference would panic on the slice index
                return prep_fail("sequence %lld: token %lld needs block %lld of a block table with %lld entries",
                                 (long long)i, (long long)j, (long long)blk, (long long)s.num_blocks);
            slots[sl++] = static_cast<int64_t>(s.block_table[blk]) * block_size + j % block_size;
        }
    }
    return nullptr;
}
