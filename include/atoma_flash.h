/*
 * atoma_flash.h — C ABI of the B200-native attention / KV-cache library (libflashattention).
 *
 * Part 1 is the DROP-IN BOUNDARY: the four symbols the reference's Rust crate `csrc` binds in
 * csrc/src/ffi.rs:3-102 (C side: csrc/kernels/flash_api.cu:22-82, cache_manager.cu:44-51, 64-71,
 * 215-233).  Names, argument order and argument types are identical, so the reference's
 * `libflashattention.a` can be swapped for this library without touching ffi.rs.
 *
 * Part 2 are extensions a B200 build wants behind the same Rust functions (explicit stream,
 * batched swap, dense GEMM); each one names the reference call site it replaces.
 *
 * Conventions shared with the reference: every entry returns void, never allocates or frees
 * caller memory, never synchronises the host; strides are in ELEMENTS; all pointers are device
 * pointers unless said otherwise.  Errors (unsupported configuration, CUDA launch failure) are
 * reported on stderr and recorded in a thread-local status readable with atoma_last_status().
 */
#ifndef ATOMA_FLASH_H_
#define ATOMA_FLASH_H_

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque CUDA stream handle (cudaStream_t / CUstream). */
typedef void *atoma_stream_t;

/* ------------------------------------------------------------------------------------------ */
/* Part 1 — the reference's C ABI (csrc/src/ffi.rs)                                            */
/* ------------------------------------------------------------------------------------------ */

/*
 * Attention forward.  Replaces csrc/kernels/flash_api.cu:22-159 (`run_mha`), bound at
 * csrc/src/ffi.rs:4-66 and called from csrc/src/lib.rs:283 (dense), :1043 (varlen / paged varlen)
 * and :1801 (kv-cache decode).  Launches on the legacy default stream, as the reference does
 * (flash_api.cu:157).
 *
 *   O = softmax(softmax_scale * Q K^T [+ alibi] [+ causal mask]) V        (fp32 softmax, 16-bit io)
 *
 *  - q/o are indexed  [batch * batch_stride | cu_seqlens_q[b] * row_stride] + row*row_stride
 *    + head*head_stride + d;  k/v likewise, or through `block_table` when it is non-NULL:
 *    token t of sequence b lives at k + block_table[b*block_table_batch_stride + t/page_block_size]
 *    * k_batch_stride + (t % page_block_size) * k_row_stride          (reference utils.h:296-314).
 *  - cu_seqlens_k is cumulative [b+1] when is_seqlens_k_cumulative, else per-sequence lengths [b]
 *    (reference block_info.h:14-24); seqused_k, when non-NULL, overrides the key length.
 *  - kv head of query head h is h / (h / h_k)                         (flash_fwd_kernel.h:593).
 *  - causal mask is bottom-right aligned: key j visible to query i iff j <= i + len_k - len_q
 *    (mask.h:171-172).  window_size_*, softcap are accepted and ignored exactly as the reference
 *    ignores them (static_switch.h:7-11).
 *  - softmax_lse (fp32) is written as [b,h,seqlen_q] when !unpadded_lse, as [h,total_q] otherwise.
 *  - num_splits / softmax_lseaccum_ptr / oaccum_ptr are accepted for ABI compatibility; the split
 *    policy is owned by the library (in-kernel reduction), the caller's accumulators are not used.
 */
void run_mha(void *q_ptr, void *k_ptr, void *v_ptr, void *o_ptr, void *softmax_lse_ptr,
             void *alibi_slopes_ptr, int32_t *cu_seqlens_q_ptr, int32_t *cu_seqlens_k_ptr,
             bool is_seqlens_k_cumulative, uint32_t q_batch_stride, uint32_t k_batch_stride,
             uint32_t v_batch_stride, uint32_t o_batch_stride, uint32_t alibi_slopes_batch_stride,
             uint32_t q_row_stride, uint32_t k_row_stride, uint32_t v_row_stride,
             uint32_t o_row_stride, uint32_t q_head_stride, uint32_t k_head_stride,
             uint32_t v_head_stride, uint32_t o_head_stride, uint32_t num_splits, uint32_t b,
             uint32_t h, uint32_t h_k, uint32_t d, uint32_t d_rounded, float softmax_scale,
             float scale_softmax_log2, int *block_table, uint32_t block_table_batch_stride,
             int page_block_size, int *seqused_k, uint32_t seqlen_q, uint32_t seqlen_k,
             uint32_t seqlen_q_rounded, uint32_t seqlen_k_rounded, int is_bf16, int is_causal,
             int window_size_left, int window_size_right, float softcap, bool unpadded_lse,
             bool force_split_kernel, void *softmax_lseaccum_ptr, void *oaccum_ptr);

/*
 * For every layer l < num_layers and pair p < num_pairs copy block src=block_mapping[2p] onto
 * dst=block_mapping[2p+1] in BOTH the key cache and the value cache of layer l.
 * key_cache_ptrs / value_cache_ptrs: device arrays of int64 base addresses [num_layers];
 * block_mapping: device int64 [num_pairs, 2]; numel_per_block counts 2-byte elements.
 * Replaces cache_manager.cu:16-81, bound at ffi.rs:68-86, called from cache_manager.rs:276, 287.
 */
void copy_blocks_f16(void *key_cache_ptrs, void *value_cache_ptrs, const void *block_mapping,
                     int64_t num_layers, int64_t num_pairs, int64_t numel_per_block,
                     atoma_stream_t stream);
void copy_blocks_bf16(void *key_cache_ptrs, void *value_cache_ptrs, const void *block_mapping,
                      int64_t num_layers, int64_t num_pairs, int64_t numel_per_block,
                      atoma_stream_t stream);

/*
 * Scatter freshly projected K/V rows into the paged caches:
 *   key_cache[slot / block_size][slot % block_size][h][:] = key[t][h][:]   with slot = slot_mapping[t],
 * tokens with slot < 0 are skipped.  key/value: [num_tokens, num_heads, head_size] with token
 * stride key_stride / value_stride; caches: [num_blocks, block_size, num_heads, head_size] with
 * block stride block_stride; dtype: 0 = f16, 1 = bf16.
 * Replaces cache_manager.cu:140-170, 215-242, bound at ffi.rs:88-101, called from
 * cache_manager.rs:512.
 */
void reshape_and_cache_flash(void *key, void *value, void *key_cache, void *value_cache,
                             int64_t *slot_mapping, int64_t block_stride, int64_t num_tokens,
                             int64_t num_heads, int64_t head_size, int64_t block_size,
                             int64_t key_stride, int64_t value_stride, uint32_t dtype,
                             atoma_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Part 2 — extensions (same semantics, B200-friendly plumbing)                                */
/* ------------------------------------------------------------------------------------------ */

/* run_mha on an explicit stream (so a decode step can be captured in a CUDA graph).  Same
 * arguments as run_mha followed by the stream.  The reference hard-codes stream 0
 * (flash_api.cu:157). */
void run_mha_stream(void *q_ptr, void *k_ptr, void *v_ptr, void *o_ptr, void *softmax_lse_ptr,
                    void *alibi_slopes_ptr, int32_t *cu_seqlens_q_ptr, int32_t *cu_seqlens_k_ptr,
                    bool is_seqlens_k_cumulative, uint32_t q_batch_stride,
                    uint32_t k_batch_stride, uint32_t v_batch_stride, uint32_t o_batch_stride,
                    uint32_t alibi_slopes_batch_stride, uint32_t q_row_stride,
                    uint32_t k_row_stride, uint32_t v_row_stride, uint32_t o_row_stride,
                    uint32_t q_head_stride, uint32_t k_head_stride, uint32_t v_head_stride,
                    uint32_t o_head_stride, uint32_t num_splits, uint32_t b, uint32_t h,
                    uint32_t h_k, uint32_t d, uint32_t d_rounded, float softmax_scale,
                    float scale_softmax_log2, int *block_table, uint32_t block_table_batch_stride,
                    int page_block_size, int *seqused_k, uint32_t seqlen_q, uint32_t seqlen_k,
                    uint32_t seqlen_q_rounded, uint32_t seqlen_k_rounded, int is_bf16,
                    int is_causal, int window_size_left, int window_size_right, float softcap,
                    bool unpadded_lse, bool force_split_kernel, void *softmax_lseaccum_ptr,
                    void *oaccum_ptr, atoma_stream_t stream);

/*
 * Fused rotary embedding + cache write (the reference's apply_rotary_embed x2 + transposes +
 * reshape_and_cache_flash, models/src/llama.rs:218-251, 269-304 and flash_attention.rs:361, as one
 * kernel).  q [T, num_heads, head_size] and k [T, num_kv_heads, head_size] (token strides q_stride /
 * k_stride, elements) are rotated IN PLACE with the non-interleaved ("rotate half") convention of
 * candle_nn::rotary_emb::rope: out[:d/2] = x1*cos - x2*sin, out[d/2:] = x1*sin + x2*cos, where
 * cos/sin [max_pos, head_size/2] (row stride rot_stride, same dtype as q) are indexed by
 * positions[t]; arithmetic in fp32, one rounding.  The rotated k and v [T, num_kv_heads, head_size]
 * are then written to key_cache / value_cache at slot_mapping[t] exactly like
 * reshape_and_cache_flash (slot < 0: token skipped; slot_mapping may be NULL: rotate only).
 */
void atoma_rope_reshape_and_cache(void *q, void *k, void *v, void *key_cache, void *value_cache,
                                  const int64_t *positions, const void *cos, const void *sin,
                                  const int64_t *slot_mapping, int64_t num_tokens, int64_t num_heads,
                                  int64_t num_kv_heads, int64_t head_size, int64_t q_stride,
                                  int64_t k_stride, int64_t v_stride, int64_t rot_stride,
                                  int64_t block_stride, int64_t block_size, uint32_t dtype,
                                  atoma_stream_t stream);

/*
 * Batched block swap: for p < num_pairs copy block_bytes bytes from
 * src_base + src_blocks[p]*block_bytes to dst_base + dst_blocks[p]*block_bytes.
 * kind: 0 = device->device (one launch of the bulk-copy kernel, cp.async.bulk through shared memory;
 * src_blocks/dst_blocks are HOST arrays, uploaded by the call), 1 = host->device, 2 = device->host: with
 * pinned (device-mapped) host memory the SAME kernel reads / writes the host buffer over the host link in
 * one launch, scattered blocks included; with pageable host memory contiguous runs of blocks are merged
 * into single cudaMemcpyAsync calls on `stream`.
 * Replaces the per-block memcpy loop of csrc/src/cache_manager.rs:18-128 + ops.rs:16-220, which
 * has no C symbol in the reference.
 */
void swap_blocks_batched(void *dst_base, const void *src_base, const int64_t *src_blocks,
                         const int64_t *dst_blocks, int64_t num_pairs, int64_t block_bytes,
                         int kind, atoma_stream_t stream);

/*
 * y[T,N] = x[T,K] * w[N,K]^T, bf16 in, fp32 accumulate, bf16 out (no bias): the q/k/v/o
 * projection of models/src/llama.rs:269-271, 311, which the reference sends to cuBLAS through
 * candle `Linear`.  ldx/ldw/ldy are row strides in elements.
 */
void atoma_gemm_bf16_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K,
                        int64_t ldx, int64_t ldw, int64_t ldy, atoma_stream_t stream);

/*
 * One-shot all-reduce(sum) over NVLink peer memory for the tensor-parallel path.
 * The reference calls NCCL through cudarc (`comm.all_reduce(src, &mut dst, &ReduceOp::Sum)`,
 * models/src/multi_gpu.rs:141-179; one communicator per GPU thread) for the [T, hidden] payload after
 * o_proj / down_proj; these entries serve the same call behind `AllReduce::cuda_fwd`.
 *
 * Every rank owns one REGION (atoma_p2p_region_bytes(slot_bytes): a header of flags and epochs plus
 * two data slots), allocated with atoma_p2p_alloc, exported as a 64-byte CUDA IPC handle
 * (atoma_p2p_export), exchanged out of band like the NCCL unique id, and mapped by every peer
 * (atoma_p2p_import).  regions[r] is rank r's region as seen from THIS process (own rank: the local
 * pointer).  The producer writes its addend into atoma_p2p_slot(regions[rank], slot_bytes, slot);
 * atoma_allreduce_sum_p2p then leaves sum_r slot_r in `out` (local, 16-byte aligned, must not alias a
 * slot), added in fp32 in rank order and rounded once, bit-identical on every rank.  Consecutive
 * all-reduces of a communicator MUST alternate slot 0 / 1; every rank must make the same calls in
 * the same order with the same count.  Stream-ordered, no host synchronisation, graph-capturable.
 * dtype: 0 = f16, 1 = bf16, 2 = f32; world 2, 4 or 8.  A peer that does not arrive within 30 s sets
 * the region's error word (atoma_p2p_error != 0; it synchronises the stream) instead of hanging.
 */
int64_t atoma_p2p_region_bytes(int64_t slot_bytes);
int atoma_p2p_alloc(void **ptr, int64_t bytes);
int atoma_p2p_free(void *ptr);
int atoma_p2p_export(void *ptr, void *handle64);
int atoma_p2p_import(const void *handle64, void **peer_ptr);
int atoma_p2p_close(void *peer_ptr);
void *atoma_p2p_slot(void *region, int64_t slot_bytes, int slot);
int atoma_p2p_error(void *region, atoma_stream_t stream);
void atoma_allreduce_sum_p2p(void *const *regions, int rank, int world, int64_t slot_bytes, int slot,
                             void *out, int64_t count, uint32_t dtype, atoma_stream_t stream);
/* Diagnostics: 0 = library's choice (default: one-shot), 1 = one-shot (every addend to every rank),
 * 2 = two-shot (addends to the vector's owner, finished vectors back).  Same results either way. */
void atoma_debug_allreduce_algo(int algo);

/*
 * Process-wide launch mode, a bit set (default 0):
 *   ATOMA_LAUNCH_PDL (1)           launch this library's kernels with programmatic dependent launch:
 *                                  a kernel's prologue (barrier / TMEM / descriptor set-up) overlaps
 *                                  the tail of the kernel before it in the stream; every kernel
 *                                  waits for its predecessor to complete before it touches memory
 *                                  the predecessor may produce.
 *   ATOMA_LAUNCH_CONST_WEIGHTS (2) the `w` operand of atoma_gemm_bf16_tn is never written by work
 *                                  in flight on the stream (inference weights): with PDL the GEMM
 *                                  starts streaming W into shared memory before the predecessor ends.
 *   (4, formerly ATOMA_LAUNCH_STABLE_CACHE, is ignored: that promise is per call, see below.)
 */
#define ATOMA_LAUNCH_PDL 1
#define ATOMA_LAUNCH_CONST_WEIGHTS 2
#define ATOMA_LAUNCH_STABLE_CACHE 4 /* deprecated, masked off by atoma_set_launch_mode */
void atoma_set_launch_mode(int flags);
int atoma_get_launch_mode(void);

/*
 * Per-call flags for the NEXT run_mha / run_mha_stream / atoma_mha_execute call made by the calling thread
 * (one-shot: consumed by that call, thread-local — no process-global state):
 *   ATOMA_CALL_STABLE_CACHE (1)  for a paged decode call (block table, seqlen_q == 1): the work still in
 *       flight on the call's stream writes no sequence length, no block-table entry and, in the pages this
 *       call's block tables reference, no cache row other than rows of each sequence's FINAL page (the slot
 *       of the token being decoded: reshape_and_cache_flash / atoma_rope_reshape_and_cache /
 *       atoma_qkv_rope_cache just before the attention call, flash_attention.rs:361 then :455; pages are
 *       exclusive to a sequence while they are written — copy-on-write).  With PDL the decode kernel then
 *       streams all older pages while its predecessor (the cache write) is still running and waits for it
 *       only before it loads q, the final page of a sequence, or writes.
 * The library itself withdraws the promise where it can see it broken: copy_blocks_f16/bf16 and
 * swap_blocks_batched into a device cache mark their stream, and the next paged decode launched on a marked
 * stream ignores the flag, waits for everything before it BEFORE releasing its own dependents, and clears
 * the mark.  atoma_llama_decode_step sets the flag itself for its attention calls (it knows the predecessor
 * is its own qkv projection).
 */
#define ATOMA_CALL_STABLE_CACHE 1
void atoma_mha_call_flags(int flags);

/* Diagnostics: when device_buf is non-null every CTA of the paged decode kernel writes eight
 * uint64 {sm id, start, end of its last KV segment, end, partial published, arrival counted,
 * reduction started, reduction finished} (globaltimer ns; the last two only on CTAs that reduce)
 * at device_buf[8*cta..]; the buffer must hold 8 * #SM entries.  NULL switches it off (default). */
void atoma_debug_decode_trace(void *device_buf);
/* Diagnostics: 0 = the decode kernel always uses separate K and V tensor maps (two ops per page);
 * 1 (default) = one map and one op for K and V when they are views of one allocation. */
void atoma_debug_decode_merge(int enable);
/* Step timeline: while device_buf (max_kernels x 1280 u64, zeroed by the caller) is set, every GEMM, paged-decode
 * and add+RMSNorm launch writes per-CTA globaltimer stamps into its own 1280-word slice, in launch order;
 * atoma_debug_step_trace_names copies the '\n'-separated kernel labels of the slices (returns their total length).
 * Pass NULL to stop.  Launch eagerly (not from a graph captured earlier) while tracing. */
void atoma_debug_step_trace(void *device_buf, int max_kernels);
int atoma_debug_step_trace_names(char *out, int cap);
/* Same for the projection GEMM: {start, first operand tile landed, last accumulator complete, end}. */
void atoma_debug_gemm_trace(void *device_buf);
/* Diagnostics: cap the GEMM's shared-memory ring at `stages` stages (0 = as many as fit). */
void atoma_debug_gemm_stages(int stages);
/* Diagnostics: 0 (default) = T <= 128 with K % 64 == 0 on the decode-time kernel (gemm_skinny.cu), everything
 * else on the stream-K kernel; 2 = everything on the stream-K kernel (A/B measurements, regression tests). */
void atoma_debug_gemm_algo(int algo);

/* 0 = last call on this thread succeeded; otherwise a negative atoma code or a cudaError_t. */
int atoma_last_status(void);
/* Human-readable text for the last non-zero status on this thread ("" if none). */
const char *atoma_last_error(void);
/* Name of the kernel path the last run_mha on this thread dispatched to (for tests/bench). */
const char *atoma_last_mha_path(void);
/* How the last paged decode launch of this thread ran: bit 0 = early KV streaming (ATOMA_CALL_STABLE_CACHE
 * honoured), bit 1 = a cache-maintenance op was pending on the stream, so the kernel waited for its
 * predecessors before releasing its dependents and the promise was ignored. */
int atoma_last_decode_mode(void);
/* Number of kernels launched by this library on this thread since the last reset. */
int64_t atoma_launch_count(void);
void atoma_reset_launch_count(void);
/* Library version string; ends in "src:<hash>" = atoma_source_hash(). */
const char *atoma_version(void);
/* SHA-256 (first 16 hex digits) of the sources the library was built from (csrc/Makefile computes it over
 * csrc/ and include/ in sorted order); __graft_entry__.build() and smoke() check it against the tree. */
const char *atoma_source_hash(void);

#ifdef __cplusplus
}
#endif
#endif /* ATOMA_FLASH_H_ */
