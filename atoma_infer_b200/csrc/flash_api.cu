// C ABI entry points (include/atoma_flash.h) and the host-side plumbing behind them: status,
// per-device attribute cache, library workspace, driver entry point for tensor maps, and the
// run_mha dispatcher.  The dispatcher mirrors the reference's run_mha / run_mha_fwd
// (csrc/kernels/flash_api.cu:8-159) in what it accepts; which kernel runs is this library's own
// decision (paged single-token decode -> decode_paged_kernel, everything else -> attn_tile_kernel).
#include "atoma_flash.h"
#include "common.cuh"
#include "mha_params.h"

#include <cstdarg>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

namespace atoma {

bool launch_attention_tile(const MhaParams &p, bool is_bf16, cudaStream_t stream);
int try_launch_decode_paged(const MhaParams &p, bool is_bf16, int64_t max_units_hint,
                            cudaStream_t stream);
int try_launch_prefill_tc(const MhaParams &p, bool is_bf16, cudaStream_t stream);
void set_decode_trace(void *device_buf);
void set_gemm_trace(void *device_buf);
void set_decode_merge(int enable);
void set_gemm_stage_limit(int stages);
void set_gemm_algo(int algo);
bool launch_gemm_bf16_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K,
                         int64_t ldx, int64_t ldw, int64_t ldy, cudaStream_t stream);

namespace {
thread_local int t_status = 0;
thread_local char t_error[512] = "";
thread_local const char *t_path = "";
thread_local int64_t t_launches = 0;
thread_local int t_decode_mode = 0;
}  // namespace
void note_decode_mode(int mode) { t_decode_mode = mode; }

namespace {
std::atomic<int> g_launch_flags{0};
}
int launch_flags() { return g_launch_flags.load(std::memory_order_relaxed); }

// Per-call promise "the work in flight on the stream writes only final-page rows of the sequences
// this decode call reads" (see ATOMA_CALL_STABLE_CACHE): set by the caller that knows it (the decode
// step executor, run_mha_stream_flags), consumed by the next paged-decode launch on this thread.
namespace {
thread_local int t_next_stable_cache = 0;
}
namespace {
std::mutex g_trace_mu;
unsigned long long *g_step_trace = nullptr;
int g_step_trace_cap = 0, g_step_trace_seq = 0;
std::string g_step_trace_names;
}
unsigned long long *next_step_trace_slot(const char *kernel_name) {
    if (!g_step_trace) return nullptr;
    std::lock_guard<std::mutex> lock(g_trace_mu);
    if (!g_step_trace || g_step_trace_seq >= g_step_trace_cap) return nullptr;
    g_step_trace_names += kernel_name;
    g_step_trace_names += '\n';
    return g_step_trace + static_cast<size_t>(g_step_trace_seq++) * kStepTraceU64;
}
namespace {
thread_local NextPrefetch t_next_prefetch = {nullptr, 0};
std::atomic<int> g_next_prefetch_off{0};
}
void set_next_kernel_prefetch(const void *ptr, long long bytes) { t_next_prefetch = NextPrefetch{ptr, ptr ? bytes : 0}; }
void set_next_kernel_prefetch_off(int off) { g_next_prefetch_off.store(off); }
NextPrefetch take_next_kernel_prefetch() {
    NextPrefetch v = t_next_prefetch;
    t_next_prefetch = NextPrefetch{nullptr, 0};
    if (g_next_prefetch_off.load(std::memory_order_relaxed)) v = NextPrefetch{nullptr, 0};
    return v;
}
void set_next_decode_stable_cache(int v) { t_next_stable_cache = v; }
int take_next_decode_stable_cache() {
    const int v = t_next_stable_cache;
    t_next_stable_cache = 0;
    return v;
}

// Streams on which a cache-maintenance op (copy_blocks, swap-in) has been enqueued since the last paged
// decode: such an op rewrites whole pages, which no "only final-page rows are being written" promise
// covers.  The next decode launch on that stream ignores the promise, waits for its predecessors before
// it touches anything AND before it lets its own dependents start (so that nothing downstream can run
// ahead of the maintenance op either), and clears the mark.
namespace {
std::mutex g_dirty_mu;
std::vector<std::pair<int, cudaStream_t>> g_dirty;
}
void mark_cache_dirty(cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_dirty_mu);
    for (const auto &e : g_dirty)
        if (e.first == dev && e.second == stream) return;
    g_dirty.emplace_back(dev, stream);
}
bool take_cache_dirty(cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_dirty_mu);
    for (size_t i = 0; i < g_dirty.size(); ++i)
        if (g_dirty[i].first == dev && g_dirty[i].second == stream) {
            g_dirty.erase(g_dirty.begin() + static_cast<long>(i));
            return true;
        }
    return false;
}

void set_status(int code, const char *fmt, ...) {
    t_status = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    // The ABI has no error channel (every entry returns void, like the reference); be loud.
    std::fprintf(stderr, "[atoma-flash] error %d: %s\n", code, t_error);
}
void clear_status() {
    t_status = 0;
    t_error[0] = 0;
}
void note_launch(int n) { t_launches += n; }
void set_mha_path(const char *name) { t_path = name; }

bool check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
        return false;
    }
    return true;
}

const DeviceInfo &current_device_info() {
    static std::mutex mu;
    static std::vector<DeviceInfo> infos(64);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    DeviceInfo &di = infos[dev];
    if (di.ordinal != dev) {
        cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
        cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        di.ordinal = dev;
    }
    return di;
}

void *get_workspace(cudaStream_t stream, size_t bytes, int tag) {
    struct Buf {
        void *ptr = nullptr;
        size_t size = 0;
    };
    static std::mutex mu;
    static std::map<std::tuple<int, int, cudaStream_t>, Buf> bufs;   // (device, tag, stream handle): exact, no hashing
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Buf &b = bufs[std::make_tuple(dev, tag, stream)];
    if (b.size >= bytes) return b.ptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
        set_status(kErrUnsupported,
                   "workspace of %zu bytes must be allocated before stream capture: run the op once "
                   "eagerly on this stream first",
                   bytes);
        return nullptr;
    }
    if (b.ptr) {
        // Never free a block that has been handed out: a CUDA graph captured earlier on this stream has
        // the old pointer baked into its kernel nodes and may be replayed after this call.  The old
        // block is retired (kept alive until process exit); workspaces grow rarely and only to the
        // largest shape seen.
        static std::vector<void *> retired;
        retired.push_back(b.ptr);
        b.ptr = nullptr;
        b.size = 0;
    }
    const size_t want = (bytes + (1u << 20) - 1) & ~static_cast<size_t>((1u << 20) - 1);
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "workspace cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        b.ptr = nullptr;
        return nullptr;
    }
    // Zero-initialised once; users keep it clean (decode arrival counters are self-resetting).
    cudaMemsetAsync(b.ptr, 0, want, stream);
    b.size = want;
    return b.ptr;
}

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
        if (e == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) set_status(kErrDriver, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

namespace {

void run_mha_impl(void *q_ptr, void *k_ptr, void *v_ptr, void *o_ptr, void *softmax_lse_ptr,
                  void *alibi_slopes_ptr, int32_t *cu_seqlens_q_ptr, int32_t *cu_seqlens_k_ptr,
                  bool is_seqlens_k_cumulative, uint32_t q_batch_stride, uint32_t k_batch_stride,
                  uint32_t v_batch_stride, uint32_t o_batch_stride,
                  uint32_t alibi_slopes_batch_stride, uint32_t q_row_stride, uint32_t k_row_stride,
                  uint32_t v_row_stride, uint32_t o_row_stride, uint32_t q_head_stride,
                  uint32_t k_head_stride, uint32_t v_head_stride, uint32_t o_head_stride,
                  uint32_t b, uint32_t h, uint32_t h_k, uint32_t d, float softmax_scale,
                  float scale_softmax_log2, int *block_table, uint32_t block_table_batch_stride,
                  int page_block_size, int *seqused_k, uint32_t seqlen_q, uint32_t seqlen_k,
                  int is_bf16, int is_causal, bool unpadded_lse, int window_size_left, int window_size_right,
                  float softcap, cudaStream_t stream) {
    clear_status();
    set_mha_path("none");
    if (b == 0 || h == 0 || seqlen_q == 0) return;
    // The reference compiles sliding windows and soft-capping OUT (FLASHATTENTION_DISABLE_LOCAL / _SOFTCAP,
    // static_switch.h:7-11) and silently computes full attention for them; so did round 1 of this library.
    // A window that would actually mask something, or a soft cap, is refused instead of answered wrongly.
    // (The tuples the reference's wrappers build for plain causal / full attention — window (seqlen_k, 0)
    // or (-1, -1) — mask nothing beyond causality and pass.)
    const bool left_masks = window_size_left >= 0 && static_cast<int64_t>(window_size_left) + 1 < static_cast<int64_t>(seqlen_k);
    const bool right_masks = window_size_right >= 0 && !is_causal && seqlen_q > 1 &&
                             static_cast<int64_t>(window_size_right) + 1 < static_cast<int64_t>(seqlen_k);
    if (softcap > 0.f || left_masks || right_masks) {
        set_status(kErrUnsupported,
                   "run_mha: sliding-window (left %d, right %d) and soft-capped (softcap %g) attention are not implemented "
                   "(the reference compiles both out and computes full attention); refusing to return an unmasked result",
                   window_size_left, window_size_right, static_cast<double>(softcap));
        return;
    }
    if (!q_ptr || !k_ptr || !v_ptr || !o_ptr) {
        set_status(kErrInvalidArg, "run_mha: null q/k/v/o pointer");
        return;
    }
    if (h_k == 0 || h % h_k != 0) {
        set_status(kErrInvalidArg, "run_mha: h=%u is not a multiple of h_k=%u", h, h_k);
        return;
    }
    if (d == 0 || d > 256 || d % 8 != 0) {
        // The reference falls through HEADDIM_SWITCH silently for d > 256 (static_switch.h:96-123).
        set_status(kErrUnsupported, "run_mha: head dim %u (need a multiple of 8, at most 256)", d);
        return;
    }
    if (block_table && page_block_size <= 0) {
        set_status(kErrInvalidArg, "run_mha: page_block_size=%d with a block table", page_block_size);
        return;
    }
    MhaParams p;
    std::memset(&p, 0, sizeof(p));
    p.q = q_ptr; p.k = k_ptr; p.v = v_ptr; p.o = o_ptr;
    p.lse = static_cast<float *>(softmax_lse_ptr);
    p.alibi = static_cast<const float *>(alibi_slopes_ptr);
    p.cu_q = cu_seqlens_q_ptr; p.cu_k = cu_seqlens_k_ptr; p.seqused_k = seqused_k;
    p.block_table = block_table;
    p.q_bs = q_batch_stride; p.k_bs = k_batch_stride; p.v_bs = v_batch_stride; p.o_bs = o_batch_stride;
    p.q_rs = q_row_stride; p.k_rs = k_row_stride; p.v_rs = v_row_stride; p.o_rs = o_row_stride;
    p.q_hs = q_head_stride; p.k_hs = k_head_stride; p.v_hs = v_head_stride; p.o_hs = o_head_stride;
    p.alibi_bs = alibi_slopes_batch_stride;
    p.bt_bs = block_table_batch_stride;
    p.b = static_cast<int>(b); p.h = static_cast<int>(h); p.h_k = static_cast<int>(h_k);
    p.d = static_cast<int>(d);
    p.seqlen_q = static_cast<int>(seqlen_q); p.seqlen_k = static_cast<int>(seqlen_k);
    p.page = block_table ? page_block_size : 1;
    p.scale = softmax_scale; p.scale_log2 = scale_softmax_log2;
    p.is_causal = is_causal ? 1 : 0;
    p.k_cumulative = is_seqlens_k_cumulative ? 1 : 0;
    p.unpadded_lse = unpadded_lse ? 1 : 0;

    // 16-byte vector / cp.async accesses: same alignment contract as the reference's 128-bit
    // gmem copies (kernel_traits.h:125-129).
    auto al16 = [](const void *x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
    const bool strides_ok = p.q_rs % 8 == 0 && p.k_rs % 8 == 0 && p.v_rs % 8 == 0 && p.o_rs % 8 == 0 &&
                            p.q_hs % 8 == 0 && p.k_hs % 8 == 0 && p.v_hs % 8 == 0 && p.o_hs % 8 == 0 &&
                            p.q_bs % 8 == 0 && p.k_bs % 8 == 0 && p.v_bs % 8 == 0 && p.o_bs % 8 == 0;
    if (!al16(q_ptr) || !al16(k_ptr) || !al16(v_ptr) || !al16(o_ptr) || !strides_ok) {
        set_status(kErrUnsupported, "run_mha: q/k/v/o pointers and strides must be 16-byte aligned");
        return;
    }

    const int r = try_launch_decode_paged(p, is_bf16 != 0, 0, stream);
    if (r == 1) {
        set_mha_path("decode_paged");
        return;
    }
    if (r < 0) return;
    const int r2 = try_launch_prefill_tc(p, is_bf16 != 0, stream);
    if (r2 == 1) {
        set_mha_path("prefill_tc");
        return;
    }
    if (r2 < 0) return;
    if (launch_attention_tile(p, is_bf16 != 0, stream)) set_mha_path("tile");
}

}  // namespace
}  // namespace atoma

using namespace atoma;

extern "C" void run_mha(void *q_ptr, void *k_ptr, void *v_ptr, void *o_ptr, void *softmax_lse_ptr,
                        void *alibi_slopes_ptr, int32_t *cu_seqlens_q_ptr,
                        int32_t *cu_seqlens_k_ptr, bool is_seqlens_k_cumulative,
                        uint32_t q_batch_stride, uint32_t k_batch_stride, uint32_t v_batch_stride,
                        uint32_t o_batch_stride, uint32_t alibi_slopes_batch_stride,
                        uint32_t q_row_stride, uint32_t k_row_stride, uint32_t v_row_stride,
                        uint32_t o_row_stride, uint32_t q_head_stride, uint32_t k_head_stride,
                        uint32_t v_head_stride, uint32_t o_head_stride, uint32_t num_splits,
                        uint32_t b, uint32_t h, uint32_t h_k, uint32_t d, uint32_t d_rounded,
                        float softmax_scale, float scale_softmax_log2, int *block_table,
                        uint32_t block_table_batch_stride, int page_block_size, int *seqused_k,
                        uint32_t seqlen_q, uint32_t seqlen_k, uint32_t seqlen_q_rounded,
                        uint32_t seqlen_k_rounded, int is_bf16, int is_causal,
                        int window_size_left, int window_size_right, float softcap,
                        bool unpadded_lse, bool force_split_kernel, void *softmax_lseaccum_ptr,
                        void *oaccum_ptr) {
    (void)num_splits; (void)d_rounded; (void)seqlen_q_rounded; (void)seqlen_k_rounded;
    (void)force_split_kernel; (void)softmax_lseaccum_ptr; (void)oaccum_ptr;
    // Legacy default stream, as the reference (flash_api.cu:157).
    run_mha_impl(q_ptr, k_ptr, v_ptr, o_ptr, softmax_lse_ptr, alibi_slopes_ptr, cu_seqlens_q_ptr,
                 cu_seqlens_k_ptr, is_seqlens_k_cumulative, q_batch_stride, k_batch_stride,
                 v_batch_stride, o_batch_stride, alibi_slopes_batch_stride, q_row_stride,
                 k_row_stride, v_row_stride, o_row_stride, q_head_stride, k_head_stride,
                 v_head_stride, o_head_stride, b, h, h_k, d, softmax_scale, scale_softmax_log2,
                 block_table, block_table_batch_stride, page_block_size, seqused_k, seqlen_q,
                 seqlen_k, is_bf16, is_causal, unpadded_lse, window_size_left, window_size_right, softcap,
                 static_cast<cudaStream_t>(0));
}

extern "C" void run_mha_stream(void *q_ptr, void *k_ptr, void *v_ptr, void *o_ptr,
                               void *softmax_lse_ptr, void *alibi_slopes_ptr,
                               int32_t *cu_seqlens_q_ptr, int32_t *cu_seqlens_k_ptr,
                               bool is_seqlens_k_cumulative, uint32_t q_batch_stride,
                               uint32_t k_batch_stride, uint32_t v_batch_stride,
                               uint32_t o_batch_stride, uint32_t alibi_slopes_batch_stride,
                               uint32_t q_row_stride, uint32_t k_row_stride, uint32_t v_row_stride,
                               uint32_t o_row_stride, uint32_t q_head_stride,
                               uint32_t k_head_stride, uint32_t v_head_stride,
                               uint32_t o_head_stride, uint32_t num_splits, uint32_t b, uint32_t h,
                               uint32_t h_k, uint32_t d, uint32_t d_rounded, float softmax_scale,
                               float scale_softmax_log2, int *block_table,
                               uint32_t block_table_batch_stride, int page_block_size,
                               int *seqused_k, uint32_t seqlen_q, uint32_t seqlen_k,
                               uint32_t seqlen_q_rounded, uint32_t seqlen_k_rounded, int is_bf16,
                               int is_causal, int window_size_left, int window_size_right,
                               float softcap, bool unpadded_lse, bool force_split_kernel,
                               void *softmax_lseaccum_ptr, void *oaccum_ptr,
                               atoma_stream_t stream) {
    (void)num_splits; (void)d_rounded; (void)seqlen_q_rounded; (void)seqlen_k_rounded;
    (void)force_split_kernel; (void)softmax_lseaccum_ptr; (void)oaccum_ptr;
    run_mha_impl(q_ptr, k_ptr, v_ptr, o_ptr, softmax_lse_ptr, alibi_slopes_ptr, cu_seqlens_q_ptr,
                 cu_seqlens_k_ptr, is_seqlens_k_cumulative, q_batch_stride, k_batch_stride,
                 v_batch_stride, o_batch_stride, alibi_slopes_batch_stride, q_row_stride,
                 k_row_stride, v_row_stride, o_row_stride, q_head_stride, k_head_stride,
                 v_head_stride, o_head_stride, b, h, h_k, d, softmax_scale, scale_softmax_log2,
                 block_table, block_table_batch_stride, page_block_size, seqused_k, seqlen_q,
                 seqlen_k, is_bf16, is_causal, unpadded_lse, window_size_left, window_size_right, softcap,
                 static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_gemm_bf16_tn(const void *x, const void *w, void *y, int64_t T, int64_t N,
                                   int64_t K, int64_t ldx, int64_t ldw, int64_t ldy,
                                   atoma_stream_t stream) {
    clear_status();
    launch_gemm_bf16_tn(x, w, y, T, N, K, ldx, ldw, ldy, static_cast<cudaStream_t>(stream));
}

extern "C" void atoma_debug_decode_trace(void *device_buf) { set_decode_trace(device_buf); }
extern "C" void atoma_debug_step_trace(void *device_buf, int max_kernels) {
    std::lock_guard<std::mutex> lock(g_trace_mu);
    g_step_trace = static_cast<unsigned long long *>(device_buf);
    g_step_trace_cap = device_buf ? max_kernels : 0;
    g_step_trace_seq = 0;
    if (device_buf) g_step_trace_names.clear();
}
extern "C" int atoma_debug_step_trace_names(char *out, int cap) {
    std::lock_guard<std::mutex> lock(g_trace_mu);
    const int n = static_cast<int>(g_step_trace_names.size());
    if (out && cap > 0) {
        const int m = std::min(n, cap - 1);
        std::memcpy(out, g_step_trace_names.data(), m);
        out[m] = 0;
    }
    return n;
}
extern "C" void atoma_debug_gemm_trace(void *device_buf) { set_gemm_trace(device_buf); }
extern "C" void atoma_debug_decode_merge(int enable) { set_decode_merge(enable); }
extern "C" void atoma_debug_gemm_stages(int stages) { set_gemm_stage_limit(stages); }
extern "C" void atoma_debug_gemm_algo(int algo) { set_gemm_algo(algo); }
// ATOMA_LAUNCH_STABLE_CACHE is no longer honoured as a process-wide bit (a promise about what is in flight
// on ONE stream before ONE call cannot be process-global state): it is masked off here; use
// atoma_mha_call_flags(ATOMA_CALL_STABLE_CACHE) before the call it applies to.
extern "C" void atoma_set_launch_mode(int flags) {
    g_launch_flags.store(flags & (kLaunchPdl | kLaunchConstWeights), std::memory_order_relaxed);
}
extern "C" void atoma_mha_call_flags(int flags) { set_next_decode_stable_cache((flags & 1) ? 1 : 0); }
extern "C" int atoma_get_launch_mode(void) { return launch_flags(); }
extern "C" int atoma_last_status(void) { return t_status; }
extern "C" const char *atoma_last_error(void) { return t_error; }
extern "C" const char *atoma_last_mha_path(void) { return t_path; }
extern "C" int atoma_last_decode_mode(void) { return t_decode_mode; }
extern "C" int64_t atoma_launch_count(void) { return t_launches; }
extern "C" void atoma_reset_launch_count(void) { t_launches = 0; }
#ifndef ATOMA_SOURCE_HASH
#define ATOMA_SOURCE_HASH "unknown"
#endif
// The hash of the sources this binary was built from (csrc/Makefile): __graft_entry__.build() and smoke()
// recompute it from the tree and refuse a library that was built from anything else.
extern "C" const char *atoma_version(void) { return "atoma-flash-b200 0.2.0 (sm_100a) src:" ATOMA_SOURCE_HASH; }
extern "C" const char *atoma_source_hash(void) { return ATOMA_SOURCE_HASH; }
