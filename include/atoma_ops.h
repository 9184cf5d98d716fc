/*
 * atoma_ops.h — host-side operator layer above the kernel ABI (atoma_flash.h).
 *
 * The reference's operator layer is Rust (crate `csrc`: csrc/src/lib.rs, csrc/src/cache_manager.rs;
 * operator `models::FlashAttention`: models/src/flash_attention.rs) and cannot be built in this
 * environment, so it is restated here in C++ behind a C ABI.  Every function performs the same
 * validation (same error texts where the reference's tests assert them), derives the same
 * `run_mha` argument tuple (csrc/src/lib.rs:283, :1043, :1801) and then calls the kernel ABI.
 *
 * Tensors are described by value (pointer already offset by the tensor's start offset, shape and
 * strides in elements), the way candle hands `CudaStorage + Layout` to a CustomOp.  Functions
 * return NULL on success or a pointer to a thread-local error message (the analogue of
 * `candle_core::bail!`).  Nothing here allocates device memory: outputs are caller-allocated, as
 * candle allocates them in the reference.
 */
#ifndef ATOMA_OPS_H_
#define ATOMA_OPS_H_

#include <stdbool.h>
#include <stdint.h>

#include "atoma_flash.h"

#ifdef __cplusplus
extern "C" {
#endif

enum atoma_dtype {
    ATOMA_F16 = 0,
    ATOMA_BF16 = 1,
    ATOMA_F32 = 2,
    ATOMA_U32 = 3,
    ATOMA_I64 = 4,
    ATOMA_U8 = 5,
    ATOMA_F64 = 6
};

#define ATOMA_MAX_RANK 5
#define ATOMA_DEVICE_CPU (-1)

typedef struct atoma_tensor {
    void *data;      /* device (or host) pointer of element [0,...,0] */
    int32_t dtype;   /* enum atoma_dtype */
    int32_t device;  /* ATOMA_DEVICE_CPU or CUDA ordinal */
    int32_t rank;
    int64_t shape[ATOMA_MAX_RANK];
    int64_t stride[ATOMA_MAX_RANK]; /* in elements */
} atoma_tensor;

/* The exact argument tuple of `run_mha` (csrc/src/ffi.rs:4-66), by name. */
typedef struct atoma_mha_args {
    void *q_ptr, *k_ptr, *v_ptr, *o_ptr, *softmax_lse_ptr, *alibi_slopes_ptr;
    int32_t *cu_seqlens_q_ptr, *cu_seqlens_k_ptr;
    bool is_seqlens_k_cumulative;
    uint32_t q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride;
    uint32_t alibi_slopes_batch_stride;
    uint32_t q_row_stride, k_row_stride, v_row_stride, o_row_stride;
    uint32_t q_head_stride, k_head_stride, v_head_stride, o_head_stride;
    uint32_t num_splits;
    uint32_t b, h, h_k, d, d_rounded;
    float softmax_scale, scale_softmax_log2;
    int *block_table;
    uint32_t block_table_batch_stride;
    int page_block_size;
    int *seqused_k;
    uint32_t seqlen_q, seqlen_k, seqlen_q_rounded, seqlen_k_rounded;
    int is_bf16, is_causal;
    int window_size_left, window_size_right;
    float softcap;
    bool unpadded_lse, force_split_kernel;
    void *softmax_lseaccum_ptr, *oaccum_ptr;
} atoma_mha_args;

/* What the reference wrapper would allocate around the call (for parity tests and callers that
 * want to pass real buffers): output shape, LSE / split-accumulator element counts. */
typedef struct atoma_mha_plan {
    atoma_mha_args args;   /* o_ptr, softmax_lse_ptr, *accum_ptr are NULL until bound */
    int32_t out_rank;
    int64_t out_shape[4];
    int64_t lse_elems;       /* f32 */
    int64_t lseaccum_elems;  /* f32, 0 when num_splits <= 1 */
    int64_t oaccum_elems;    /* f32, 0 when num_splits <= 1 */
    int32_t seqlenq_ngroups_swapped;
} atoma_mha_plan;

/* -1 in a window argument means `None`. num_sms <= 0: query the current device. */

/* csrc::flash_attn* family (csrc/src/lib.rs:31-343, 392-571): q [b,sq,h,d], k/v [b,sk,hk,d]. */
const char *atoma_flash_attn_plan(const atoma_tensor *q, const atoma_tensor *k,
                                  const atoma_tensor *v, const atoma_tensor *alibi_slopes,
                                  float softmax_scale, int64_t window_size_left,
                                  int64_t window_size_right, float softcap, int num_sms,
                                  atoma_mha_plan *plan);

/* csrc::flash_attn_varlen* family (csrc/src/lib.rs:606-1103, 1160-1501): q [total_q,h,d];
 * k/v [total_k,hk,d] or, with a block table, [num_blocks,page,hk,d]. */
const char *atoma_flash_attn_varlen_plan(const atoma_tensor *q, const atoma_tensor *k,
                                         const atoma_tensor *v, const atoma_tensor *alibi_slopes,
                                         const atoma_tensor *seqlens_q,
                                         const atoma_tensor *seqlens_k,
                                         const atoma_tensor *seqused_k,
                                         const atoma_tensor *block_table, int64_t max_seqlen_q,
                                         int64_t max_seqlen_k, float softmax_scale,
                                         int64_t window_size_left, int64_t window_size_right,
                                         float softcap, int num_sms, atoma_mha_plan *plan);

/* csrc::flash_attn_kv_cache* family (csrc/src/lib.rs:1523-1855, 1907-2105): q [b,sq,h,d];
 * k/v cache [b_cache,sk,hk,d] or, with a block table, [num_blocks,page,hk,d]. */
const char *atoma_flash_attn_kv_cache_plan(const atoma_tensor *q, const atoma_tensor *k_cache,
                                           const atoma_tensor *v_cache,
                                           const atoma_tensor *alibi_slopes,
                                           const atoma_tensor *block_table,
                                           const atoma_tensor *seqlens_k, float softmax_scale,
                                           int64_t window_size_left, int64_t window_size_right,
                                           int num_sms, atoma_mha_plan *plan);

/* Bind the output (required) and optional LSE / accumulators, then launch (run_mha_stream).
 * `use_legacy_stream` != 0 calls the reference symbol `run_mha` (stream 0) instead. */
const char *atoma_mha_execute(const atoma_mha_plan *plan, void *out, void *softmax_lse,
                              void *softmax_lseaccum, void *oaccum, int use_legacy_stream,
                              atoma_stream_t stream);

/* utils::compute_num_splits (csrc/src/lib.rs:2122-2199). */
uint32_t atoma_compute_num_splits(int64_t batch_size, int64_t num_heads, int64_t head_size,
                                  int64_t max_seqlen_k, int64_t max_seqlen_q, int num_sms);

/* csrc::reshape_and_cache_flash (csrc/src/cache_manager.rs:319-535). */
const char *atoma_reshape_and_cache_flash(const atoma_tensor *key, const atoma_tensor *value,
                                          const atoma_tensor *key_cache,
                                          const atoma_tensor *value_cache,
                                          const atoma_tensor *slot_mapping, atoma_stream_t stream);

/* csrc::copy_blocks (csrc/src/cache_manager.rs:148-307).  key_caches/value_caches: arrays of
 * per-layer tensors [num_blocks, block_size, hk, d]; block_mapping: device i64 [num_pairs, 2].
 * ptr_scratch: device buffer of >= 2*num_layers int64 the per-layer base addresses are uploaded
 * to (the reference builds two candle tensors for this, cache_manager.rs:221-222). */
const char *atoma_copy_blocks(const atoma_tensor *key_caches, int64_t num_key_caches,
                              const atoma_tensor *value_caches, int64_t num_value_caches,
                              const atoma_tensor *block_mapping, void *ptr_scratch,
                              atoma_stream_t stream);

/* csrc::swap_blocks (csrc/src/cache_manager.rs:18-128): dst[block d] = src[block s] for every
 * (s, d) of the mapping; src/dst [num_blocks, block_size, hk, d] on the same GPU or CPU<->GPU. */
const char *atoma_swap_blocks(const atoma_tensor *src, const atoma_tensor *dst,
                              const int64_t *src_blocks, const int64_t *dst_blocks,
                              int64_t num_pairs, atoma_stream_t stream);

/* models::FlashAttention::forward (models/src/flash_attention.rs:322-469). */
typedef struct atoma_attention_config {
    int64_t num_heads, num_kv_heads, head_dim;
    float softmax_scale;
    int64_t sliding_window; /* -1 = None */
} atoma_attention_config;

typedef struct atoma_attention_metadata {
    const atoma_tensor *slot_mapping;              /* i64 [T] */
    int64_t num_prefill_tokens, num_decoding_tokens;
    /* prefill (may be absent: has_prefill = 0) */
    int32_t has_prefill;
    const atoma_tensor *prefill_block_tables;      /* NULL = no prefix cache */
    int64_t max_prefill_sequence_length;
    const atoma_tensor *query_start_locations;     /* u32, may be NULL */
    const atoma_tensor *sequence_start_locations;  /* u32 [Bp+1] */
    int64_t max_prefill_sequence_length_k;         /* only with prefill_block_tables */
    /* decode (may be absent: has_decode = 0) */
    int32_t has_decode;
    const atoma_tensor *decode_block_tables;       /* u32 [Bd, max_blocks] */
    const atoma_tensor *decode_sequence_lengths;   /* u32 [Bd] */
} atoma_attention_metadata;

/* q [T,h,d], k/v [T,hk,d], kv_cache [2,num_blocks,block_size,hk,d]; out [T,h*d] (caller
 * allocated, fully written).  alibi_slopes may be NULL. */
const char *atoma_flash_attention_forward(const atoma_attention_config *cfg, const atoma_tensor *q,
                                          const atoma_tensor *k, const atoma_tensor *v,
                                          const atoma_tensor *kv_cache,
                                          const atoma_tensor *alibi_slopes,
                                          const atoma_attention_metadata *md,
                                          const atoma_tensor *out, atoma_stream_t stream);

/*
 * Step input preparation — `ModelWorker::prepare_input_tensors`
 * (backends/vllm/src/worker.rs:224-460), SURVEY §8 row f2.  One flattened entry per sequence of
 * the scheduled sequence groups (group fields repeated per sequence), in the order the reference
 * iterates them.  Every tensor of the step's `ModelInput` / `FlashAttentionMetadata` is laid out in
 * ONE host buffer (16-byte aligned sections, offsets in `atoma_model_input_layout`), so one
 * host-to-device copy ships the whole step; device tensors are views at the same offsets.
 */
typedef struct atoma_sequence {
    const uint32_t *token_ids;   /* sequence_data.get_token_ids(): prompt + generated tokens */
    int64_t length;              /* sequence_data.length() */
    int64_t num_computed_tokens; /* sequence_data.get_num_computed_tokens() (prompts only) */
    int64_t token_chunk_size;    /* sequence_group_metadata.token_chunk_size */
    const uint32_t *block_table; /* physical block numbers; NULL = the group has no block tables
                                    (memory profiling run, worker.rs:365-371) */
    int64_t num_blocks;
    int32_t is_prompt;           /* sequence_group_metadata.is_prompt */
} atoma_sequence;

typedef struct atoma_model_input_layout {
    int64_t total_bytes;
    int64_t input_tokens_off;       /* u32 [num_tokens]              ModelInput::input_tokens_tensor */
    int64_t input_positions_off;    /* i64 [num_tokens]              ModelInput::input_positions */
    int64_t slot_mapping_off;       /* i64 [num_slots]               FlashAttentionMetadata::slot_mapping */
    int64_t sequence_lengths_off;   /* u32 [num_sequences]           sequence_lengths (window-capped for decodes) */
    int64_t context_lengths_off;    /* u32 [num_sequences]           context_lengths */
    int64_t sequence_start_loc_off; /* u32 [num_sequences + 1]       cumulative sequence lengths */
    int64_t query_start_loc_off;    /* u32 [num_sequences + 1]       cumulative query lengths */
    int64_t block_tables_off;       /* u32 [num_sequences, max_block_table_len], padded with 0 */
    int64_t num_tokens, num_slots, num_sequences, max_block_table_len;
    int64_t num_prefills, num_prefill_tokens, num_decode_tokens;
    int64_t max_query_len, max_prefill_seq_len, max_decode_seq_len;
} atoma_model_input_layout;

/* sliding_window: 0 = none.  With buffer == NULL only `layout` (sizes, offsets, scalars) is
 * computed; otherwise `buffer` (host memory of at least layout->total_bytes, ideally pinned) is
 * filled as well.  Returns NULL or an error text (the reference's where it has one: "Empty prompts
 * should not be received in `ModelWorker`", "Block table should be allocated for sequence on
 * decoding phase"). */
const char *atoma_prepare_input(const atoma_sequence *seqs, int64_t num_sequences, int64_t block_size,
                                int64_t sliding_window, int enable_chunked_prefill,
                                atoma_model_input_layout *layout, void *buffer, int64_t buffer_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ATOMA_OPS_H_ */
