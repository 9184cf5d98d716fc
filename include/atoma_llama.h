/*
 * atoma_llama.h — the decode-step operators around the attention path and the step executor
 * (SURVEY §8 rows a11, f1, f2, f3, f4).
 *
 * In the reference every one of these is a candle call issued from Rust: the projections are
 * `Linear::forward` -> cuBLAS (models/src/llama.rs:269-271, 311, 364-365, 476), RoPE is
 * candle_nn::rotary_emb::rope behind 2 index_select + 4 transpose/contiguous copies
 * (llama.rs:218-304), the cache write is csrc::reshape_and_cache_flash (flash_attention.rs:361),
 * SwiGLU is candle_nn::ops::silu + a multiply (llama.rs:364), the norms are candle RmsNorm
 * (llama.rs:398-410, 474), sampling is a per-sequence host copy of the logits
 * (backends/vllm/src/model_executor.rs:227-262), and under tensor parallelism the row-parallel
 * projections are followed by ncclAllReduce (models/src/multi_gpu.rs:141-179).  There is no C
 * boundary for any of them in the reference; these symbols are what a `CustomOp` wrapper in
 * `models::llama` binds instead (INTEGRATION.md shows the Rust side).
 *
 * Conventions: row-major 16-bit tensors (dtype 0 = f16, 1 = bf16), strides in elements, pointers and
 * row strides 16-byte aligned, everything stream-ordered and CUDA-graph capturable, nothing
 * allocates.  Errors: atoma_last_status() / atoma_last_error() (atoma_flash.h).
 */
#ifndef ATOMA_LLAMA_H_
#define ATOMA_LLAMA_H_

#include <stdint.h>

#include "atoma_flash.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * y[T,N] = x[T,K] w[N,K]^T for any 16-bit dtype (atoma_gemm_bf16_tn with a dtype argument).
 * T <= 128 with K % 64 == 0 runs the decode-time weight-streaming kernel (gemm_skinny.cu: no split-K,
 * tile width chosen to fill one wave of SMs); other shapes run the stream-K kernel (gemm_bf16.cu).
 */
void atoma_gemm_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx,
                   int64_t ldw, int64_t ldy, uint32_t dtype, atoma_stream_t stream);

/*
 * Fused q/k/v projection + rotary embedding + KV-cache write, ONE kernel (row f1):
 *   qkv[T, (H + 2 Hkv) D] = x[T,K] wqkv^T, wqkv = [Wq; Wk; Wv] stacked along N (multi_gpu.rs:26-35
 *   `load_multi` builds exactly this concatenation), head-major rows;
 *   q and k heads are rotated (non-interleaved rope: out[:D/2] = x1 cos - x2 sin, out[D/2:] = x1 sin +
 *   x2 cos on the projection rounded to the model dtype, fp32 arithmetic, one rounding; cos/sin
 *   [max_pos, D/2] rows selected by positions[t]) and everything is stored to qkv; the rotated k and
 *   v of token t are also written to key_cache / value_cache [NB, block_size, Hkv, D] at
 *   slot_mapping[t] (slot < 0: skipped; slot_mapping or the caches may be NULL: no cache write) —
 *   bit-identical to atoma_gemm_tn + atoma_rope_reshape_and_cache.
 * T > 128 is processed in 128-row passes.
 */
void atoma_qkv_rope_cache(const void *x, const void *wqkv, void *qkv, int64_t T, int64_t K, int64_t ldx,
                          int64_t ldw, int64_t ldy, int64_t num_heads, int64_t num_kv_heads,
                          int64_t head_size, const int64_t *positions, const void *cos, const void *sin,
                          int64_t rot_stride, const int64_t *slot_mapping, void *key_cache,
                          void *value_cache, int64_t block_stride, int64_t block_size, uint32_t dtype,
                          atoma_stream_t stream);

/*
 * Fused gate/up projection + SwiGLU, ONE kernel (row f3; llama.rs:364):
 *   y[T, inter] = silu(x w_gate^T) * (x w_up^T), w_gate_up = [W_gate; W_up] stacked along N
 *   ([2 inter, K]); both projections are rounded to the model dtype first (the reference
 *   materialises them), silu(g) * u is evaluated in fp32 and rounded once.
 */
void atoma_gate_up_silu(const void *x, const void *w_gate_up, void *y, int64_t T, int64_t inter, int64_t K,
                        int64_t ldx, int64_t ldw, int64_t ldy, uint32_t dtype, atoma_stream_t stream);

/*
 * The MLP of a block (llama.rs:362-366): act[T, inter] = silu(x w_gate^T) * (x w_up^T) and y[T, hidden] = act w_down^T.
 * With programmatic dependent launch enabled, T <= 128 and one wave of tiles per projection both run in ONE kernel
 * (gemm_chain2_kernel: the CTAs keep their ring, barriers and tensor memory across the two projections and the second
 * one's weights start streaming while the first drains; a grid-wide arrival counter gates only its activation loads)
 * — bit-identical to atoma_gate_up_silu + atoma_gemm_tn, which is what runs otherwise.  act is written either way.
 */
void atoma_mlp_gate_up_down(const void *x, const void *w_gate_up, void *act, const void *w_down, void *y, int64_t T,
                            int64_t inter, int64_t hidden, int64_t ldx, int64_t ldw_gate_up, int64_t ld_act,
                            int64_t ldw_down, int64_t ldy, uint32_t dtype, atoma_stream_t stream);

/*
 * Residual add + RMSNorm, ONE kernel (llama.rs:398-410):
 *   h = x + residual (model dtype; either may be NULL = absent), residual_out = h (may be NULL, may
 *   alias x or residual), normed = dtype( h * rsqrt(mean(h^2) + eps) * weight ) in fp32 (may be NULL).
 */
void atoma_add_rmsnorm(const void *x, const void *residual, void *residual_out, const void *weight,
                       void *normed, int64_t T, int64_t hidden, int64_t ldx, int64_t ldr, int64_t ldro,
                       int64_t ldn, float eps, uint32_t dtype, atoma_stream_t stream);

/*
 * Row-parallel projection + residual add + RMSNorm, ONE kernel (llama.rs:311 + 401-407; 365 + 408 + 399):
 *   y = x w^T (as atoma_gemm_tn; T <= 128, K % 64 == 0), then h = y + residual -> residual_out and
 *   normed = rmsnorm(h) * weight exactly as atoma_add_rmsnorm computes them (bit-identical to the two calls).  The
 *   CTAs of the GEMM that finish last normalise the rows: no second launch.  N = the normalised width (<= 8192).
 */
void atoma_gemm_tn_add_rmsnorm(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx,
                               int64_t ldw, int64_t ldy, const void *residual, void *residual_out, const void *weight,
                               void *normed, int64_t ldr, int64_t ldro, int64_t ldn, float eps, uint32_t dtype,
                               atoma_stream_t stream);

/* out[t, :] = table[tokens[t], :] (llama.rs:470 `wte.forward`). */
void atoma_embedding(const uint32_t *tokens, const void *table, void *out, int64_t T, int64_t hidden,
                     int64_t vocab, int64_t ld_table, int64_t ldo, uint32_t dtype, atoma_stream_t stream);

/*
 * Greedy sampling on the device (model_executor.rs:227-262 + candle LogitsProcessor::sample_argmax):
 * token_out[t] = the LAST index of the maximum of logits[t, :] under f32 total order, logit_out[t]
 * (may be NULL) = that logit as f32 (what the reference reports as the token's "logprob"),
 * logits_f32 (may be NULL) = the row converted to f32 (llama.rs:477 `to_dtype(F32)`).
 */
void atoma_sample_greedy(const void *logits, int64_t T, int64_t vocab, int64_t ld, uint32_t dtype,
                         uint32_t *token_out, float *logit_out, float *logits_f32, int64_t ld_f32,
                         atoma_stream_t stream);

/*
 * Greedy sampling over a vocabulary sharded across tensor-parallel ranks: logits [T, rows] is this rank's slice
 * (global indices row0 .. row0 + rows); every rank stores its row candidates {order key, global index, epoch} into
 * every rank's region (first staging slot; one 16-byte unit per candidate) and keeps the best of all ranks' — the
 * argmax of the full row with atoma_sample_greedy's tie rule, identical on every rank.  Must be called by every
 * rank in the same order; T <= 128; slot_bytes >= 32 KiB.
 */
void atoma_sample_greedy_tp(const void *logits, int64_t T, int64_t rows, int64_t row0, int64_t ld, uint32_t dtype,
                            uint32_t *token_out, float *logit_out, void *const *regions, int rank, int world,
                            int64_t slot_bytes, atoma_stream_t stream);

/*
 * Tensor-parallel row-parallel projection with the all-reduce fused in (row f4; multi_gpu.rs:141-179).
 * Send side: atoma_gemm_tn_push computes this rank's partial y[T,N] = x[T,K_shard] w^T and its
 * epilogue stores the bf16/f16 partial tile straight into every rank's receive buffer over NVLink
 * (regions[r]: the peer-mapped regions of atoma_p2p_alloc/export/import, slot_bytes >= T*N*2, N % 32 == 0;
 * 8-byte units {4 bytes of data, epoch}, laid out in 32-column blocks so that a CTA's tile is one contiguous
 * run of the peer's memory and a warp's stores coalesce) — no staging buffer in global memory, no separate
 * all-reduce launch.
 * Receive side: atoma_allreduce_recv_add_rmsnorm polls the local receive buffers, adds the `world`
 * addends in fp32 in rank order and rounds once (bit-identical on every rank), then continues exactly
 * like atoma_add_rmsnorm with that sum as x.  The pair must be called by every rank in the same order;
 * the receive buffer used alternates with the parity of a device-resident epoch counter that the
 * receive side advances, so consecutive all-reduces never collide and graphs replay.  T <= 128.
 */
void atoma_gemm_tn_push(const void *x, const void *w, int64_t T, int64_t N, int64_t K, int64_t ldx,
                        int64_t ldw, void *const *regions, int rank, int world, int64_t slot_bytes,
                        uint32_t dtype, atoma_stream_t stream);
void atoma_allreduce_recv_add_rmsnorm(void *const *regions, int rank, int world, int64_t slot_bytes,
                                      const void *residual, void *residual_out, const void *weight,
                                      void *normed, int64_t T, int64_t hidden, int64_t ldr, int64_t ldro,
                                      int64_t ldn, float eps, uint32_t dtype, atoma_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Decode-step executor (rows f2 + f3): models/src/llama.rs:456-478 for a decode-only batch     */
/* ------------------------------------------------------------------------------------------ */

typedef struct atoma_llama_layer {
    const void *wqkv;      /* [(H + 2 Hkv) D, hidden]  (this rank's heads) */
    const void *wo;        /* [hidden, H D]            (this rank's columns) */
    const void *w_gate_up; /* [2 inter, hidden]        (this rank's inter = intermediate / world) */
    const void *w_down;    /* [hidden, inter] */
    const void *rms_attn;  /* [hidden] input_layernorm */
    const void *rms_mlp;   /* [hidden] post_attention_layernorm */
    void *key_cache;       /* [NB, block_size, Hkv, D] */
    void *value_cache;
} atoma_llama_layer;

typedef struct atoma_llama_model {
    int64_t hidden, inter, num_layers, num_heads, num_kv_heads, head_dim, vocab; /* per-rank heads / inter */
    int64_t block_size, block_stride; /* KV page geometry (elements) */
    int64_t rot_stride;               /* row stride of cos / sin */
    float rms_eps, softmax_scale;
    uint32_t dtype;
    const void *embed;   /* [vocab, hidden] or NULL (then the step starts from `hidden_in`) */
    const void *norm_f;  /* [hidden] */
    const void *lm_head; /* [vocab, hidden] or NULL (the step ends with the final norm) */
    const void *cos, *sin;
    const atoma_llama_layer *layers; /* host array [num_layers] */
    /* tensor parallelism: world 1 (no communication) or 2/4/8 with the regions of atoma_p2p_* */
    int32_t tp_rank, tp_world;
    void *const *p2p_regions;
    int64_t p2p_slot_bytes;
    /* vocab-parallel head (tp_world > 1 only; 0 rows = lm_head is the full [vocab, hidden] matrix, replicated as in
     * llama_nccl.rs:320-321): lm_head points at rows [lm_head_row0, lm_head_row0 + lm_head_rows) of the matrix, a
     * multiple of 8 rows; the step then returns next_tokens / next_logits only (each rank takes the argmax of its
     * slice and the ranks exchange their candidates over peer memory — atoma_sample_greedy_tp). */
    int64_t lm_head_rows, lm_head_row0;
} atoma_llama_model;

typedef struct atoma_decode_batch {
    int64_t num_tokens;              /* T <= 128: one token per sequence */
    const uint32_t *tokens;          /* device [T] (used when model.embed != NULL) */
    const void *hidden_in;           /* device [T, hidden] (used when model.embed == NULL) */
    const int64_t *positions;        /* device [T] */
    const int64_t *slot_mapping;     /* device [T] */
    const int32_t *block_tables;     /* device [T, max_blocks] */
    const int32_t *sequence_lengths; /* device [T], including the token being decoded */
    int64_t max_blocks;
    void *scratch;                   /* device, >= atoma_llama_decode_scratch_bytes() */
    int64_t scratch_bytes;
    /* outputs (device; each may be NULL) */
    void *hidden_out;      /* [T, hidden] final normed hidden state */
    void *logits;          /* [T, vocab] model dtype */
    float *logits_f32;     /* [T, vocab] */
    uint32_t *next_tokens; /* [T] greedy */
    float *next_logits;    /* [T] */
} atoma_decode_batch;

/* sizeof of the structs above as this library was compiled (0 = atoma_llama_layer, 1 = atoma_llama_model,
 * 2 = atoma_decode_batch): lets a binding in another language assert its mirror of the layout at start-up. */
int64_t atoma_llama_abi_sizeof(int which);
int64_t atoma_llama_decode_scratch_bytes(const atoma_llama_model *model, int64_t num_tokens);
/* One decode step: embedding -> num_layers x [add+rmsnorm, qkv+rope+cache write, paged decode
 * attention, o_proj (+all-reduce), add+rmsnorm, gate_up+silu, down_proj (+all-reduce)] -> add +
 * final norm -> lm_head -> greedy sample: 6 launches per layer with programmatic dependent launch enabled (gate_up and
 * down_proj share one kernel), 7 without; no host synchronisation. */
void atoma_llama_decode_step(const atoma_llama_model *model, const atoma_decode_batch *batch,
                             atoma_stream_t stream);

/* Health of the step's in-kernel grid barriers on `stream` (synchronises it; not during capture).  The chained MLP
 * kernel and the fused norm tail wait inside the kernel for every CTA of their own grid; the wait is bounded (2 s): if
 * it ever expires — a CTA trapped, or another grid with such a barrier was running on the same device at the same
 * time and neither became fully resident — the kernel finishes with an unusable result and sets a sticky word.
 * Returns 0 = healthy, bit 0 = fused norm tail timed out, bit 1 = chained MLP timed out (atoma_last_error() says so),
 * -1 = the query itself failed.  Serving loops call it where they already synchronise (the sampled-token copy). */
int atoma_llama_step_status(atoma_stream_t stream);

/* Diagnostics for the decode-time GEMM: per-CTA timestamps (globaltimer ns) at device_buf[8*cta..]: {0 start,
 * 1 first stage landed, 2 first accumulator complete, 3 end, 4 the SM id, 5 tensor memory released, 6 x producer's
 * grid dependency resolved, 7 last stage landed}; tile width / k blocks per stage / ring depth overrides
 * (0 = automatic). */
void atoma_debug_skinny_trace(void *device_buf);
void atoma_debug_skinny_tuning(int np, int kbs, int stages);
/* Timing experiments only (results are garbage): skip_loads bit 0 = issue no x loads, bit 1 = no W loads;
 * w_ops = 1 | 2 weight TMA ops per stage. */
void atoma_debug_skinny_mode(int skip_loads, int w_ops);
/* Shared-memory ring budget of the decode-time GEMM in KiB (0 = default).  <= 100 selects the half-footprint
 * configuration: two CTAs fit on an SM, so the next kernel of the stream becomes resident and streams its
 * weights while this one drains (programmatic dependent launch). */
void atoma_debug_skinny_budget(int kib);
/* Cross-kernel L2 weight prefetch of atoma_llama_decode_step (each weight-streaming kernel prefetches the start
 * of what the next one streams while its own tail drains).  mask bits: 1 attention -> Wo, 2 o_proj -> Wgate_up,
 * 4 gate_up -> Wdown, 8 down_proj -> next Wqkv / lm_head, 16 qkv -> Wo; default 0 (measured: no gain).  cap_mib: bytes per edge
 * (default 32; 0 keeps the current value). */
void atoma_debug_step_prefetch(int mask, int cap_mib);
/* atoma_llama_decode_step, bit 1 (value 2): gate_up + SwiGLU and down_proj as two kernels instead of the chained one
 * (default: chained, 6 launches per layer).  Bit 0: 1 = the add+RMSNorm after o_proj / down_proj runs in the tail of those GEMMs
 * (5 launches per layer, under tensor parallelism the receive side of the all-reduce included); 0 (default) =
 * separate kernels (7 launches per layer; measured 1.5 us per norm faster).  Same bits either way. */
void atoma_debug_step_fusion(int flags);

#ifdef __cplusplus
}
#endif
#endif /* ATOMA_LLAMA_H_ */
