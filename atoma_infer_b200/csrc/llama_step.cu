// C ABI of include/atoma_llama.h: the decode-step operators and the decode-step executor.
//
// The executor restates, for a decode-only batch, what the reference assembles from candle calls in
// models/src/llama.rs:253-314 (CausalSelfAttention::forward), :362-366 (Mlp::forward), :398-410
// (Block::forward), :456-478 (Llama::forward) and, under tensor parallelism, llama_nccl.rs:97-139 +
// multi_gpu.rs:141-179 — as 6 kernel launches per layer on one stream (7 without programmatic dependent launch):
//   add + RMSNorm | qkv GEMM (+RoPE +cache write) | paged decode attention | o_proj GEMM (+push) |
//   add + RMSNorm (+all-reduce receive) | gate_up GEMM (+SwiGLU) chained with down GEMM (+push)
// It owns no memory (weights, caches, scratch and outputs are the caller's) and never synchronises,
// so the caller captures it in a CUDA graph once per batch shape.
#include "atoma_llama.h"
#include "common.cuh"
#include "mha_params.h"
#include "skinny_args.h"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a tool (nsys, ncu) is attached

#include <cmath>
#include <cstring>

namespace atoma {

bool launch_gemm_bf16_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx,
                         int64_t ldw, int64_t ldy, cudaStream_t stream);
bool launch_add_rmsnorm(const void *x, const void *residual, void *residual_out, const void *weight, void *normed,
                        int64_t T, int64_t hidden, int64_t ldx, int64_t ldr, int64_t ldro, int64_t ldn, float eps,
                        uint32_t dtype, void *p2p_region, int64_t p2p_slot_bytes, int world, cudaStream_t stream);
bool launch_embedding(const uint32_t *tokens, const void *table, void *out, int64_t T, int64_t hidden, int64_t vocab,
                      int64_t ld_table, int64_t ldo, uint32_t dtype, cudaStream_t stream);
bool launch_argmax(const void *logits, int64_t T, int64_t vocab, int64_t ld, uint32_t dtype, uint32_t *token_out,
                   float *logit_out, float *logits_f32, int64_t ld_f32, cudaStream_t stream);
bool launch_argmax_tp(const void *logits, int64_t T, int64_t rows, int64_t row0, int64_t ld, uint32_t dtype,
                      uint32_t *token_out, float *logit_out, void *const *regions, int rank, int world, int64_t slot_bytes,
                      cudaStream_t stream);
void p2p_layout(int64_t slot_bytes, int64_t *recv_offset, int64_t *slot_stride, int64_t *recv_stride);
int try_launch_decode_paged(const MhaParams &p, bool is_bf16, int64_t max_units_hint, cudaStream_t stream);
void set_next_decode_stable_cache(int v);

namespace {

inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Cross-kernel L2 prefetch of the step (see common.cuh: NextPrefetch).  Every weight-streaming kernel of a layer
// is told what the NEXT one will stream first; bit i of the mask enables one edge (diagnostics: A/B runs).
//   1: attention -> Wo    2: o_proj -> Wgate_up    4: gate_up -> Wdown    8: down_proj -> next layer's Wqkv / lm_head
//  16: qkv -> Wo (instead of 1: the prefetch then sits in L2 through the attention kernel)
// Default 0 (off): measured on a B200 (profiles/r02_step_prefetch_ab.json) every combination is within +-1 us per layer of
// "off" — at a kernel boundary the dead time is control latency (CTA hand-over, set-up, descriptor fetch, first
// HBM round trip), not bandwidth, and an L2 hit shortens only the last of those.
int g_step_prefetch_mask = 0;
long long g_step_prefetch_cap = 32ll << 20;   // bytes per edge: ~5 us of HBM time, a quarter of L2

inline void hint(int bit, const void *w, long long bytes) {
    if (g_step_prefetch_mask & bit) set_next_kernel_prefetch(w, std::min(bytes, g_step_prefetch_cap));
}

// NVTX range for the lifetime of the object (the reference marks its forward passes with
// tracing spans — models/src/llama.rs:258, 363, 397; these are the profiler-visible equivalent).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// `tail` (optional): residual add + RMSNorm of y's rows fused into the kernel's tail (norm_row.cuh); tail->x is set here.
bool gemm_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx, int64_t ldw,
             int64_t ldy, uint32_t dtype, cudaStream_t stream, const NormArgs *tail = nullptr) {
    if (T <= 0 || N <= 0) return true;
    if (dtype > 1 || K <= 0 || !x || !w || !y || !al16(x) || !al16(w) || ldx % 8 || ldw % 8 || ldx < K || ldw < K ||
        ldy < N) {
        set_status(kErrInvalidArg,
                   "atoma_gemm_tn: dtype 0=f16 1=bf16, x and w 16-byte aligned with row strides that are multiples of 8 "
                   "elements (K=%lld ldx=%lld ldw=%lld ldy=%lld dtype=%u)",
                   static_cast<long long>(K), static_cast<long long>(ldx), static_cast<long long>(ldw),
                   static_cast<long long>(ldy), dtype);
        return false;
    }
    if (skinny_gemm_supported(T, K, ldx, ldw) && N < (1ll << 31)) {
        SkinnyArgs a{};
        a.epi = kEpiPlain;
        a.y = y;
        a.ldy = ldy;
        if (tail) {
            a.norm_tail = 1;
            a.norm = *tail;
            a.norm.x = y;
            a.norm.ldx = ldy;
        }
        return launch_skinny_gemm(a, x, w, static_cast<int>(T), K, ldx, ldw, 1, static_cast<int>(N), 1, 0, 0, dtype == 1,
                                  stream);
    }
    if (tail) {
        set_status(kErrUnsupported, "atoma_gemm_tn_add_rmsnorm: needs T <= 128 and K %% 64 == 0 (T=%lld K=%lld)",
                   static_cast<long long>(T), static_cast<long long>(K));
        return false;
    }
    if (dtype != 1) {
        set_status(kErrUnsupported, "atoma_gemm_tn: f16 needs T <= 128 and K %% 64 == 0 (T=%lld K=%lld)",
                   static_cast<long long>(T), static_cast<long long>(K));
        return false;
    }
    return launch_gemm_bf16_tn(x, w, y, T, N, K, ldx, ldw, ldy, stream);
}

bool qkv_rope_cache(const void *x, const void *wqkv, void *qkv, int64_t T, int64_t K, int64_t ldx, int64_t ldw,
                    int64_t ldy, int64_t H, int64_t Hkv, int64_t D, const int64_t *positions, const void *cos,
                    const void *sin, int64_t rot_stride, const int64_t *slot_mapping, void *key_cache, void *value_cache,
                    int64_t block_stride, int64_t block_size, uint32_t dtype, cudaStream_t stream) {
    if (T <= 0) return true;
    const int64_t N = (H + 2 * Hkv) * D;
    if (dtype > 1 || H <= 0 || Hkv <= 0 || D <= 0 || D % 16 || !positions || !cos || !sin || !x || !wqkv || !qkv ||
        !al16(x) || !al16(wqkv) || !al16(qkv) || !al16(cos) || !al16(sin) || !al16(key_cache) || !al16(value_cache) ||
        ldx % 8 || ldw % 8 || ldy % 8 || rot_stride % 8 || block_stride % 8 || ldx < K || ldw < K || ldy < N ||
        (key_cache && (block_size <= 0 || !value_cache))) {
        set_status(kErrInvalidArg,
                   "atoma_qkv_rope_cache: 16-bit dtype, head_size a multiple of 16, all pointers and strides 16-byte aligned "
                   "(H=%lld Hkv=%lld D=%lld)",
                   static_cast<long long>(H), static_cast<long long>(Hkv), static_cast<long long>(D));
        return false;
    }
    if (!skinny_gemm_supported(std::min<int64_t>(T, 128), K, ldx, ldw)) {
        set_status(kErrUnsupported, "atoma_qkv_rope_cache: needs K %% 64 == 0 and an sm_100a device (K=%lld)",
                   static_cast<long long>(K));
        return false;
    }
    const int64_t esz = 2;
    for (int64_t t0 = 0; t0 < T; t0 += 128) {
        const int tn = static_cast<int>(std::min<int64_t>(128, T - t0));
        SkinnyArgs a{};
        a.epi = kEpiRope;
        a.y = static_cast<char *>(qkv) + t0 * ldy * esz;
        a.ldy = ldy;
        a.positions = positions + t0;
        a.cos = cos;
        a.sin = sin;
        a.rot_stride = rot_stride;
        a.slot_mapping = slot_mapping ? slot_mapping + t0 : nullptr;
        a.k_cache = key_cache;
        a.v_cache = value_cache;
        a.block_stride = block_stride;
        a.block_size = block_size;
        a.num_heads = static_cast<int>(H);
        a.num_kv_heads = static_cast<int>(Hkv);
        if (!launch_skinny_gemm(a, static_cast<const char *>(x) + t0 * ldx * esz, wqkv, tn, K, ldx, ldw,
                                static_cast<int>(H + 2 * Hkv), static_cast<int>(D / 2), 2, D / 2, D, dtype == 1, stream))
            return false;
    }
    return true;
}

bool gate_up_silu(const void *x, const void *w, void *y, int64_t T, int64_t inter, int64_t K, int64_t ldx, int64_t ldw,
                  int64_t ldy, uint32_t dtype, cudaStream_t stream) {
    if (T <= 0 || inter <= 0) return true;
    if (dtype > 1 || !x || !w || !y || !al16(x) || !al16(w) || !al16(y) || ldx % 8 || ldw % 8 || ldy % 8 || ldx < K ||
        ldw < K || ldy < inter || inter % 8) {
        set_status(kErrInvalidArg, "atoma_gate_up_silu: 16-bit dtype, inter a multiple of 8, 16-byte aligned pointers and strides");
        return false;
    }
    if (!skinny_gemm_supported(std::min<int64_t>(T, 128), K, ldx, ldw)) {
        set_status(kErrUnsupported, "atoma_gate_up_silu: needs K %% 64 == 0 and an sm_100a device (K=%lld)",
                   static_cast<long long>(K));
        return false;
    }
    for (int64_t t0 = 0; t0 < T; t0 += 128) {
        const int tn = static_cast<int>(std::min<int64_t>(128, T - t0));
        SkinnyArgs a{};
        a.epi = kEpiSilu;
        a.y = static_cast<char *>(y) + t0 * ldy * 2;
        a.ldy = ldy;
        if (!launch_skinny_gemm(a, static_cast<const char *>(x) + t0 * ldx * 2, w, tn, K, ldx, ldw, 1,
                                static_cast<int>(inter), 2, inter, 0, dtype == 1, stream))
            return false;
    }
    return true;
}

void fill_push_args(SkinnyArgs &a, void *const *regions, int rank, int world, int64_t slot_bytes, int64_t T, int64_t N) {
    a.epi = kEpiPush;
    for (int r = 0; r < 8; ++r) a.push_region[r] = static_cast<unsigned char *>(regions[r < world ? r : rank]);
    a.push_state = reinterpret_cast<const uint32_t *>(regions[rank]);
    p2p_layout(slot_bytes, &a.push_recv_offset, &a.push_slot_stride, &a.push_recv_stride);
    a.push_plane_bytes = T * N * 2;
    a.push_rank = rank;
    a.push_world = world;
}

bool gemm_tn_push(const void *x, const void *w, int64_t T, int64_t N, int64_t K, int64_t ldx, int64_t ldw,
                  void *const *regions, int rank, int world, int64_t slot_bytes, uint32_t dtype, cudaStream_t stream,
                  const NormArgs *tail = nullptr) {
    if (T <= 0 || N <= 0) return true;
    if (dtype > 1 || !regions || world < 2 || world > 8 || (world & (world - 1)) || rank < 0 || rank >= world || N % 32 ||
        T * N * 2 > slot_bytes || !x || !w || !al16(x) || !al16(w) || ldx % 8 || ldw % 8 || ldx < K || ldw < K) {
        set_status(kErrInvalidArg,
                   "atoma_gemm_tn_push: world 2/4/8, 0 <= rank < world, N a multiple of 32, T*N*2 <= slot_bytes, aligned "
                   "operands (T=%lld N=%lld rank=%d world=%d slot=%lld)",
                   static_cast<long long>(T), static_cast<long long>(N), rank, world, static_cast<long long>(slot_bytes));
        return false;
    }
    if (!skinny_gemm_supported(T, K, ldx, ldw)) {
        set_status(kErrUnsupported, "atoma_gemm_tn_push: needs T <= 128, K %% 64 == 0 and an sm_100a device (T=%lld K=%lld)",
                   static_cast<long long>(T), static_cast<long long>(K));
        return false;
    }
    SkinnyArgs a{};
    fill_push_args(a, regions, rank, world, slot_bytes, T, N);
    if (tail) {   // receive side (the all-reduce's sum + residual + RMSNorm) in the same kernel
        a.norm_tail = 1;
        a.norm = *tail;
        a.norm.x = nullptr;
        norm_args_set_p2p(a.norm, regions[rank], slot_bytes, world);
    }
    return launch_skinny_gemm(a, x, w, static_cast<int>(T), K, ldx, ldw, 1, static_cast<int>(N), 1, 0, 0, dtype == 1,
                              stream);
}

// Arguments of a fused tail: h = y + residual -> residual_out; normed = rmsnorm(h) * weight.
NormArgs norm_tail_args(const void *residual, void *residual_out, const void *weight, void *normed, int64_t T,
                        int64_t hidden, int64_t ldr, int64_t ldro, int64_t ldn, float eps) {
    NormArgs n{};
    n.residual = residual; n.residual_out = residual_out; n.weight = weight; n.normed = normed;
    n.ldr = ldr; n.ldro = ldro; n.ldn = ldn;
    n.hidden = static_cast<int>(hidden);
    n.T = static_cast<int>(T);
    n.eps = eps;
    n.world = 1;
    return n;
}

// 1: the add+RMSNorm after o_proj and after down_proj run in the tail of those GEMMs (5 launches per layer);
// 0 (default): as separate kernels (7 launches per layer) — atoma_debug_step_fusion.  Measured on a B200
// (profiles/r02_step_fusion_ab.json): the fused tail is 1.5 us per norm SLOWER (190.2 vs 187.1 us per layer at tp = 1,
// 66.5 vs 64.2 for a tp8 shard): the norm CTAs must first wait for the slowest CTA of the grid, whereas the separate
// norm kernel is already resident (PDL) and its 32 CTAs start the moment the GEMM's grid completes.  Same bits.
int g_step_norm_tail = 0;

struct Scratch {
    char *resid, *normed, *qkv, *attn, *proj, *act, *logits;
    int64_t bytes;
};

Scratch carve(const atoma_llama_model &m, int64_t T, char *base) {
    auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
    Scratch s{};
    int64_t off = 0;
    auto take = [&](int64_t elems) {
        char *p = base ? base + off : nullptr;
        off += up(elems * 2);
        return p;
    };
    s.resid = take(T * m.hidden);
    s.normed = take(T * m.hidden);
    s.qkv = take(T * (m.num_heads + 2 * m.num_kv_heads) * m.head_dim);
    s.attn = take(T * m.num_heads * m.head_dim);
    s.proj = take(T * m.hidden);
    s.act = take(T * m.inter);
    s.logits = take(T * m.vocab);
    s.bytes = off;
    return s;
}

}  // namespace
}  // namespace atoma

using namespace atoma;

extern "C" void atoma_gemm_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx,
                              int64_t ldw, int64_t ldy, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    gemm_tn(x, w, y, T, N, K, ldx, ldw, ldy, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_qkv_rope_cache(const void *x, const void *wqkv, void *qkv, int64_t T, int64_t K, int64_t ldx,
                                     int64_t ldw, int64_t ldy, int64_t num_heads, int64_t num_kv_heads,
                                     int64_t head_size, const int64_t *positions, const void *cos, const void *sin,
                                     int64_t rot_stride, const int64_t *slot_mapping, void *key_cache,
                                     void *value_cache, int64_t block_stride, int64_t block_size, uint32_t dtype,
                                     atoma_stream_t stream) {
    clear_status();
    qkv_rope_cache(x, wqkv, qkv, T, K, ldx, ldw, ldy, num_heads, num_kv_heads, head_size, positions, cos, sin, rot_stride,
                   slot_mapping, key_cache, value_cache, block_stride, block_size, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_gate_up_silu(const void *x, const void *w_gate_up, void *y, int64_t T, int64_t inter, int64_t K,
                                   int64_t ldx, int64_t ldw, int64_t ldy, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    gate_up_silu(x, w_gate_up, y, T, inter, K, ldx, ldw, ldy, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_add_rmsnorm(const void *x, const void *residual, void *residual_out, const void *weight,
                                  void *normed, int64_t T, int64_t hidden, int64_t ldx, int64_t ldr, int64_t ldro,
                                  int64_t ldn, float eps, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    launch_add_rmsnorm(x, residual, residual_out, weight, normed, T, hidden, ldx, ldr, ldro, ldn, eps, dtype, nullptr, 0, 1,
                       static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_embedding(const uint32_t *tokens, const void *table, void *out, int64_t T, int64_t hidden,
                                int64_t vocab, int64_t ld_table, int64_t ldo, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    launch_embedding(tokens, table, out, T, hidden, vocab, ld_table, ldo, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_sample_greedy(const void *logits, int64_t T, int64_t vocab, int64_t ld, uint32_t dtype,
                                    uint32_t *token_out, float *logit_out, float *logits_f32, int64_t ld_f32,
                                    atoma_stream_t stream) {
    clear_status();
    launch_argmax(logits, T, vocab, ld, dtype, token_out, logit_out, logits_f32, ld_f32, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_sample_greedy_tp(const void *logits, int64_t T, int64_t rows, int64_t row0, int64_t ld, uint32_t dtype,
                                       uint32_t *token_out, float *logit_out, void *const *regions, int rank, int world,
                                       int64_t slot_bytes, atoma_stream_t stream) {
    clear_status();
    launch_argmax_tp(logits, T, rows, row0, ld, dtype, token_out, logit_out, regions, rank, world, slot_bytes,
                     static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_gemm_tn_push(const void *x, const void *w, int64_t T, int64_t N, int64_t K, int64_t ldx,
                                   int64_t ldw, void *const *regions, int rank, int world, int64_t slot_bytes,
                                   uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    gemm_tn_push(x, w, T, N, K, ldx, ldw, regions, rank, world, slot_bytes, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_allreduce_recv_add_rmsnorm(void *const *regions, int rank, int world, int64_t slot_bytes,
                                                 const void *residual, void *residual_out, const void *weight,
                                                 void *normed, int64_t T, int64_t hidden, int64_t ldr, int64_t ldro,
                                                 int64_t ldn, float eps, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    if (!regions || world < 2 || world > 8 || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_status(kErrInvalidArg, "atoma_allreduce_recv_add_rmsnorm: world 2/4/8 and 0 <= rank < world (rank %d, world %d)",
                   rank, world);
        return;
    }
    launch_add_rmsnorm(nullptr, residual, residual_out, weight, normed, T, hidden, 8, ldr, ldro, ldn, eps, dtype,
                       regions[rank], slot_bytes, world, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t atoma_llama_decode_scratch_bytes(const atoma_llama_model *model, int64_t num_tokens) {
    if (!model || num_tokens <= 0) return 0;
    return carve(*model, num_tokens, nullptr).bytes;
}

extern "C" void atoma_llama_decode_step(const atoma_llama_model *model, const atoma_decode_batch *batch,
                                        atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!model || !batch) {
        set_status(kErrInvalidArg, "atoma_llama_decode_step: null model or batch");
        return;
    }
    const atoma_llama_model &m = *model;
    const atoma_decode_batch &b = *batch;
    const int64_t T = b.num_tokens;
    if (T <= 0) return;
    const int world = m.tp_world > 1 ? m.tp_world : 1;
    if (T > 128 || m.dtype > 1 || !m.layers || m.num_layers <= 0 || !m.norm_f || (!m.embed && !b.hidden_in) ||
        (m.embed && !b.tokens) || !b.positions || !b.slot_mapping || !b.block_tables || !b.sequence_lengths ||
        !b.scratch || b.scratch_bytes < carve(m, T, nullptr).bytes || (m.lm_head && !b.logits && !b.next_tokens && !b.logits_f32) ||
        (world > 1 && (!m.p2p_regions || T * m.hidden * 2 > m.p2p_slot_bytes))) {
        set_status(kErrInvalidArg,
                   "atoma_llama_decode_step: need 1 <= T <= 128 (got %lld), a 16-bit dtype, layers, norm_f, tokens or hidden_in, "
                   "positions / slot_mapping / block_tables / sequence_lengths, scratch of atoma_llama_decode_scratch_bytes() "
                   "and, for tp_world > 1, p2p regions with slot_bytes >= T * hidden * 2",
                   static_cast<long long>(T));
        return;
    }
    NvtxRange step_range("atoma_llama_decode_step");
    const Scratch s = carve(m, T, static_cast<char *>(b.scratch));
    const int64_t hid = m.hidden, H = m.num_heads, Hkv = m.num_kv_heads, D = m.head_dim;
    const int64_t nq = H * D, nqkv = (H + 2 * Hkv) * D;
    const uint32_t dt = m.dtype;

    // x = wte(tokens)  (llama.rs:470) -> the residual stream
    const void *resid_in;
    if (m.embed) {
        if (!launch_embedding(b.tokens, m.embed, s.resid, T, hid, m.vocab, hid, hid, dt, stream)) return;
        resid_in = s.resid;
    } else {
        resid_in = b.hidden_in;
    }
    // The add + RMSNorm that follows each row-parallel projection runs in that GEMM's tail (gemm_skinny.cu) unless
    // switched off or the row does not fit the GEMM's CTA: 5 launches per layer instead of 7.
    const bool fuse = g_step_norm_tail && hid <= 8 * 256 * kNormMaxVec && skinny_gemm_supported(T, nq, nq, nq) &&
                      skinny_gemm_supported(T, m.inter, m.inter, m.inter);
    void *final_normed = b.hidden_out ? b.hidden_out : s.normed;
    for (int64_t l = 0; l < m.num_layers; ++l) {
        const atoma_llama_layer &ly = m.layers[l];
        NvtxRange layer_range("block");
        // residual += previous block's MLP output (all-reduced under TP); x = rms_1(residual)  (llama.rs:398-400)
        if (l == 0) {
            if (!launch_add_rmsnorm(nullptr, resid_in, s.resid, ly.rms_attn, s.normed, T, hid, hid, hid, hid, hid, m.rms_eps,
                                    dt, nullptr, 0, 1, stream))
                return;
        } else if (fuse) {
            // done in the tail of the previous layer's down_proj
        } else if (world > 1) {
            if (!launch_add_rmsnorm(nullptr, s.resid, s.resid, ly.rms_attn, s.normed, T, hid, hid, hid, hid, hid, m.rms_eps,
                                    dt, m.p2p_regions[m.tp_rank], m.p2p_slot_bytes, world, stream))
                return;
        } else {
            if (!launch_add_rmsnorm(s.proj, s.resid, s.resid, ly.rms_attn, s.normed, T, hid, hid, hid, hid, hid, m.rms_eps,
                                    dt, nullptr, 0, 1, stream))
                return;
        }
        // q, k, v projections + RoPE + cache write  (llama.rs:269-304, flash_attention.rs:361)
        hint(16, ly.wo, hid * nq * 2);
        if (!qkv_rope_cache(s.normed, ly.wqkv, s.qkv, T, hid, hid, hid, nqkv, H, Hkv, D, b.positions, m.cos, m.sin,
                            m.rot_stride, b.slot_mapping, ly.key_cache, ly.value_cache, m.block_stride, m.block_size, dt,
                            stream))
            return;
        // paged decode attention  (flash_attention.rs:455 -> flash_attn_kv_cache_full -> run_mha; the
        // argument tuple of SURVEY appendix A.1 with q a strided view of the qkv buffer)
        MhaParams p;
        std::memset(&p, 0, sizeof(p));
        p.q = s.qkv;
        p.k = ly.key_cache;
        p.v = ly.value_cache;
        p.o = s.attn;
        p.cu_k = b.sequence_lengths;
        p.block_table = b.block_tables;
        p.q_bs = nqkv; p.q_rs = nqkv; p.q_hs = D;
        p.o_bs = nq; p.o_rs = nq; p.o_hs = D;
        p.k_bs = p.v_bs = m.block_stride;
        p.k_rs = p.v_rs = Hkv * D;
        p.k_hs = p.v_hs = D;
        p.bt_bs = b.max_blocks;
        p.b = static_cast<int>(T); p.h = static_cast<int>(H); p.h_k = static_cast<int>(Hkv); p.d = static_cast<int>(D);
        p.seqlen_q = 1;
        p.seqlen_k = static_cast<int>(b.max_blocks * m.block_size);
        p.page = static_cast<int>(m.block_size);
        p.scale = m.softmax_scale;
        p.scale_log2 = m.softmax_scale * 1.4426950408889634f;
        p.is_causal = 0;       // seqlen_q == 1 folds causal away (csrc/src/lib.rs:1627-1631)
        p.k_cumulative = 0;    // per-sequence lengths (csrc/src/lib.rs:1771)
        // The kernel before this one in the stream (the qkv projection above) writes only the slot of the
        // token being decoded, i.e. rows of each sequence's final page: the decode kernel may stream the
        // older pages before its grid dependency resolves.  A per-call promise, made here where it holds.
        set_next_decode_stable_cache(1);
        hint(1, ly.wo, hid * nq * 2);
        const int r = try_launch_decode_paged(p, dt == 1, 0, stream);
        if (r != 1) {
            if (r == 0)
                set_status(kErrUnsupported, "atoma_llama_decode_step: the paged decode kernel does not support this geometry "
                                            "(head_dim %lld, block_size %lld, H/Hkv %lld)",
                           static_cast<long long>(D), static_cast<long long>(m.block_size), static_cast<long long>(H / Hkv));
            return;
        }
        // o_proj (+ all-reduce send side)  (llama.rs:311; llama_nccl.rs:139)
        hint(2, ly.w_gate_up, 2 * m.inter * hid * 2);
        // x = attn + residual; rms_2(x)  (llama.rs:401-407)
        const NormArgs tail_mlp = norm_tail_args(s.resid, s.resid, ly.rms_mlp, s.normed, T, hid, hid, hid, hid, m.rms_eps);
        if (fuse) {
            if (world > 1) {
                if (!gemm_tn_push(s.attn, ly.wo, T, hid, nq, nq, nq, m.p2p_regions, m.tp_rank, world, m.p2p_slot_bytes, dt, stream,
                                  &tail_mlp))
                    return;
            } else {
                if (!gemm_tn(s.attn, ly.wo, s.proj, T, hid, nq, nq, nq, hid, dt, stream, &tail_mlp)) return;
            }
        } else if (world > 1) {
            if (!gemm_tn_push(s.attn, ly.wo, T, hid, nq, nq, nq, m.p2p_regions, m.tp_rank, world, m.p2p_slot_bytes, dt, stream))
                return;
            if (!launch_add_rmsnorm(nullptr, s.resid, s.resid, ly.rms_mlp, s.normed, T, hid, hid, hid, hid, hid, m.rms_eps, dt,
                                    m.p2p_regions[m.tp_rank], m.p2p_slot_bytes, world, stream))
                return;
        } else {
            if (!gemm_tn(s.attn, ly.wo, s.proj, T, hid, nq, nq, nq, hid, dt, stream)) return;
            // x = attn + residual; rms_2(x)  (llama.rs:401-407)
            if (!launch_add_rmsnorm(s.proj, s.resid, s.resid, ly.rms_mlp, s.normed, T, hid, hid, hid, hid, hid, m.rms_eps, dt,
                                    nullptr, 0, 1, stream))
                return;
        }
        // MLP: silu(gate) * up, down  (llama.rs:362-366) — ONE kernel when the shapes allow (gemm_chain2_kernel)
        int chained = 0;
        if (!fuse && hid % 32 == 0) {
            SkinnyArgs a1{};
            if (world > 1) {
                fill_push_args(a1, m.p2p_regions, m.tp_rank, world, m.p2p_slot_bytes, T, hid);
            } else {
                a1.epi = kEpiPlain;
                a1.y = s.proj;
                a1.ldy = hid;
            }
            chained = try_launch_mlp_chain(s.normed, ly.w_gate_up, s.act, ly.w_down, a1, static_cast<int>(T), m.inter, hid, hid, hid,
                                           m.inter, m.inter, dt == 1, stream);
            if (chained < 0) return;
        }
        if (!chained) {
            hint(4, ly.w_down, hid * m.inter * 2);
            if (!gate_up_silu(s.normed, ly.w_gate_up, s.act, T, m.inter, hid, hid, hid, m.inter, dt, stream)) return;
            if (l + 1 < m.num_layers) hint(8, m.layers[l + 1].wqkv, nqkv * hid * 2);
            else if (m.lm_head) hint(8, m.lm_head, m.vocab * hid * 2);
            if (fuse) {
                // x = mlp + residual, then the NEXT block's rms_1 — or ln_f after the last block  (llama.rs:408, 399, 474)
                const bool last = l + 1 == m.num_layers;
                const NormArgs tail_next = norm_tail_args(s.resid, last ? nullptr : s.resid, last ? m.norm_f : m.layers[l + 1].rms_attn,
                                                          last ? final_normed : s.normed, T, hid, hid, hid, hid, m.rms_eps);
                if (world > 1) {
                    if (!gemm_tn_push(s.act, ly.w_down, T, hid, m.inter, m.inter, m.inter, m.p2p_regions, m.tp_rank, world,
                                      m.p2p_slot_bytes, dt, stream, &tail_next))
                        return;
                } else {
                    if (!gemm_tn(s.act, ly.w_down, s.proj, T, hid, m.inter, m.inter, m.inter, hid, dt, stream, &tail_next)) return;
                }
            } else if (world > 1) {
                if (!gemm_tn_push(s.act, ly.w_down, T, hid, m.inter, m.inter, m.inter, m.p2p_regions, m.tp_rank, world,
                                  m.p2p_slot_bytes, dt, stream))
                    return;
            } else {
                if (!gemm_tn(s.act, ly.w_down, s.proj, T, hid, m.inter, m.inter, m.inter, hid, dt, stream)) return;
            }
        }
    }
    // x = mlp + residual; ln_f  (llama.rs:408, 474)
    if (!fuse &&
        !launch_add_rmsnorm(world > 1 ? nullptr : s.proj, s.resid, nullptr, m.norm_f, final_normed, T, hid, hid, hid, hid, hid,
                            m.rms_eps, dt, world > 1 ? m.p2p_regions[m.tp_rank] : nullptr, m.p2p_slot_bytes, world, stream))
        return;
    if (!m.lm_head) return;
    if (world > 1 && m.lm_head_rows > 0) {
        // vocab-parallel head: this rank's slice of the logits, then the candidates are exchanged over peer memory
        if (b.logits || b.logits_f32 || !b.next_tokens || m.lm_head_rows % 8 || m.lm_head_row0 < 0 ||
            m.lm_head_row0 + m.lm_head_rows > m.vocab) {
            set_status(kErrInvalidArg, "atoma_llama_decode_step: a vocab-parallel head (lm_head_rows = %lld, a multiple of 8 inside "
                                       "the vocabulary) returns next_tokens / next_logits only",
                       static_cast<long long>(m.lm_head_rows));
            return;
        }
        if (!gemm_tn(final_normed, m.lm_head, s.logits, T, m.lm_head_rows, hid, hid, hid, m.lm_head_rows, dt, stream)) return;
        launch_argmax_tp(s.logits, T, m.lm_head_rows, m.lm_head_row0, m.lm_head_rows, dt, b.next_tokens, b.next_logits,
                         m.p2p_regions, m.tp_rank, world, m.p2p_slot_bytes, stream);
        return;
    }
    // logits = lm_head(x)  (llama.rs:476; every token of a decode batch is selected), greedy sample
    void *logits = b.logits ? b.logits : s.logits;
    if (!gemm_tn(final_normed, m.lm_head, logits, T, m.vocab, hid, hid, hid, m.vocab, dt, stream)) return;
    if (b.next_tokens || b.logits_f32) {
        uint32_t *tok = b.next_tokens ? b.next_tokens : reinterpret_cast<uint32_t *>(s.act);
        launch_argmax(logits, T, m.vocab, m.vocab, dt, tok, b.next_logits, b.logits_f32, m.vocab, stream);
    }
}

// The in-kernel grid barriers of the step (the chained MLP's "every CTA has stored its activations" counter, the
// fused norm tail's "every CTA has finished" counter) are bounded: a wait that outlives 2 s — a CTA of the grid
// trapped, or two grids with such a barrier were forced to share the device and neither became fully resident —
// gives up, lets the kernel finish with an unusable result and sets a sticky word in the stream's barrier
// workspace.  This reads those words back (synchronises the stream): bit 0 = fused norm tail, bit 1 = chained MLP.
extern "C" int atoma_llama_step_status(atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
        set_status(kErrUnsupported, "atoma_llama_step_status synchronises the stream: not during capture");
        return -1;
    }
    const uint32_t *ws = static_cast<const uint32_t *>(get_workspace(stream, 256, kWorkspaceSkinnyTail));
    if (!ws) return -1;
    uint32_t w[16] = {0};
    cudaError_t e = cudaMemcpyAsync(w, ws, sizeof(w), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "atoma_llama_step_status: %s", cudaGetErrorString(e));
        return -1;
    }
    const int bits = (w[2] ? 1 : 0) | (w[10] ? 2 : 0);   // tail_sync[2]; the chain's sync = words 8.. -> sync[2] = word 10
    if (bits)
        set_status(kErrUnsupported, "a grid barrier inside a decode-step GEMM timed out (%s%s): the step's results since then are unusable",
                   (bits & 1) ? "fused norm tail" : "", (bits & 2) ? ((bits & 1) ? ", chained MLP" : "chained MLP") : "");
    return bits;
}

extern "C" int64_t atoma_llama_abi_sizeof(int which) {
    switch (which) {
        case 0: return static_cast<int64_t>(sizeof(atoma_llama_layer));
        case 1: return static_cast<int64_t>(sizeof(atoma_llama_model));
        case 2: return static_cast<int64_t>(sizeof(atoma_decode_batch));
        default: return -1;
    }
}

extern "C" void atoma_debug_skinny_trace(void *device_buf) { set_skinny_trace(device_buf); }
extern "C" void atoma_debug_skinny_tuning(int np, int kbs, int stages) { set_skinny_tuning(np, kbs, stages); }
extern "C" void atoma_debug_skinny_mode(int skip_loads, int w_ops) { set_skinny_debug(skip_loads, w_ops); }
extern "C" void atoma_debug_skinny_budget(int kib) { set_skinny_budget(kib); }
extern "C" void atoma_mlp_gate_up_down(const void *x, const void *w_gate_up, void *act, const void *w_down, void *y, int64_t T,
                                       int64_t inter, int64_t hidden, int64_t ldx, int64_t ldw_gate_up, int64_t ld_act,
                                       int64_t ldw_down, int64_t ldy, uint32_t dtype, atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (T <= 0) return;
    if (dtype > 1 || !x || !w_gate_up || !act || !w_down || !y || !al16(x) || !al16(w_gate_up) || !al16(act) || !al16(w_down) ||
        !al16(y) || ldx % 8 || ldw_gate_up % 8 || ld_act % 8 || ldw_down % 8 || ldy % 8 || inter % 8 || ldx < hidden ||
        ldw_gate_up < hidden || ld_act < inter || ldw_down < inter || ldy < hidden) {
        set_status(kErrInvalidArg, "atoma_mlp_gate_up_down: 16-bit dtype, inter a multiple of 8, 16-byte aligned pointers and strides");
        return;
    }
    if (T <= 128) {
        SkinnyArgs a1{};
        a1.epi = kEpiPlain;
        a1.y = y;
        a1.ldy = ldy;
        const int r = try_launch_mlp_chain(x, w_gate_up, act, w_down, a1, static_cast<int>(T), inter, hidden, ldx, ldw_gate_up, ld_act,
                                           ldw_down, dtype == 1, stream);
        if (r != 0) return;
    }
    if (!gate_up_silu(x, w_gate_up, act, T, inter, hidden, ldx, ldw_gate_up, ld_act, dtype, stream)) return;
    gemm_tn(act, w_down, y, T, hidden, inter, ld_act, ldw_down, ldy, dtype, stream);
}

extern "C" void atoma_debug_step_fusion(int flags) {
    g_step_norm_tail = (flags & 1) ? 1 : 0;
    set_skinny_chain((flags & 2) ? 0 : 1);   // bit 1 set: gate_up and down_proj as two kernels
}

extern "C" void atoma_gemm_tn_add_rmsnorm(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K, int64_t ldx,
                                          int64_t ldw, int64_t ldy, const void *residual, void *residual_out,
                                          const void *weight, void *normed, int64_t ldr, int64_t ldro, int64_t ldn, float eps,
                                          uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    if (N % 8 || N > 8 * 256 * kNormMaxVec || (normed && !weight) || !al16(residual) || !al16(residual_out) || !al16(weight) ||
        !al16(normed) || ldr % 8 || ldro % 8 || ldn % 8 || ldy % 8 || !al16(y)) {
        set_status(kErrInvalidArg, "atoma_gemm_tn_add_rmsnorm: N a multiple of 8 and <= %d, 16-byte aligned pointers and row strides",
                   8 * 256 * kNormMaxVec);
        return;
    }
    const NormArgs tail = norm_tail_args(residual, residual_out, weight, normed, T, N, ldr, ldro, ldn, eps);
    gemm_tn(x, w, y, T, N, K, ldx, ldw, ldy, dtype, static_cast<cudaStream_t>(stream), &tail);
}
extern "C" void atoma_debug_step_prefetch(int mask, int cap_mib) {
    g_step_prefetch_mask = mask;
    if (cap_mib > 0) g_step_prefetch_cap = static_cast<long long>(cap_mib) << 20;
}
