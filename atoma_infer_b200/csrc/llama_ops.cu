// The small kernels of a Llama decode step around the projections: token
// embedding gather, residual add + RMSNorm (plain, and as the receive side of the tensor-parallel
// all-reduce whose send side is the o_proj / down_proj GEMM epilogue, row f4), and greedy sampling
// of the lm_head logits on the device.
//
// Reference behaviour restated (third-party arithmetic: candle-nn / candle-kernels 0.9.2-alpha.1 are
// registry dependencies, not under /root/reference — parity for these ops is unpinned there; the
// call sites are what is followed):
//   * models/src/llama.rs:398-410  Block::forward:  x = attn(rms_1(x)) + residual;  x = mlp(rms_2(x)) + residual
//     — the adds are in the model dtype (bf16), RMSNorm computes in fp32 on the bf16 input:
//     y = bf16( x * rsqrt(mean(x^2) + eps) * w ).
//   * models/src/llama.rs:456-478  Llama::forward: wte gather, blocks, ln_f, index_select of the selected
//     token rows, lm_head, logits.to_dtype(F32).
//   * backends/vllm/src/model_executor.rs:227-262: per sequence, copy the full-vocab logits row to the
//     host (to_vec1) and sample there; greedy = candle_transformers LogitsProcessor::sample_argmax =
//     `iter().enumerate().max_by(total_cmp)`, i.e. the LAST index among equal maxima; the reported
//     "logprob" is the raw logit of the sampled token.  Here: one kernel, 8 bytes per sequence leave
//     the device instead of vocab * 4.
//   * models/src/multi_gpu.rs:141-179 AllReduce::cuda_fwd: ncclAllReduce(sum) of the [T, hidden] bf16
//     partial projection; result in bf16.
#include "common.cuh"
#include "norm_row.cuh"

#include "atoma_flash.h"

#include <algorithm>

namespace atoma {
namespace {

using namespace normrow;

__device__ __forceinline__ unsigned long long ops_timer_ns() { return timer_ns(); }

// One CTA per token row (norm_row.cuh).
template <typename T, int W>
__global__ void __launch_bounds__(512) add_rmsnorm_kernel(const NormArgs a) {
    __shared__ float red[32];
    __shared__ uint32_t s_state[2];
    if (a.trace && threadIdx.x == 0) a.trace[blockIdx.x * 8 + 1] = ops_timer_ns();
    pdl_launch_dependents();
    uint4 wv[kNormMaxVec];
    if (a.const_weight) norm_weight_preload(a, wv);   // a model weight: one L2 / HBM round trip off the dependent path
    pdl_wait();
    if (a.trace && threadIdx.x == 0) a.trace[blockIdx.x * 8 + 2] = ops_timer_ns();
    const int row = blockIdx.x;
    uint32_t *state = reinterpret_cast<uint32_t *>(a.region);
    uint32_t e = 0;
    if (W > 1) {
        if (threadIdx.x == 0) {
            s_state[0] = *reinterpret_cast<volatile uint32_t *>(&state[0]);
            s_state[1] = *reinterpret_cast<volatile uint32_t *>(&state[2]);
        }
        __syncthreads();
        e = s_state[0] + 1;               // this all-reduce's epoch
        if (s_state[1] != 0) return;      // a peer failed to arrive earlier: do not wait for it again (block-uniform)
    }
    const bool failed = add_rmsnorm_row<T, W>(a, row, e, red, a.const_weight ? wv : nullptr);
    if (a.trace && threadIdx.x == 0) a.trace[blockIdx.x * 8 + 3] = ops_timer_ns();
    if (W > 1) {
        if (failed && threadIdx.x == 0) atomicExch(&state[2], 1u);   // error word: a peer never arrived
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&state[1], 1u) == gridDim.x - 1) {   // last CTA: the next all-reduce sees the next epoch
                state[1] = 0;
                __threadfence();
                state[0] = e;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) embedding_kernel(const uint32_t *__restrict__ tokens, const T *__restrict__ table,
                                                        T *__restrict__ out, int hidden, int64_t ld_table, int64_t ldo,
                                                        int64_t vocab) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    int64_t tok = tokens[row];
    if (tok >= vocab) tok = vocab - 1;   // candle's index_select would fail the launch; never fault here
    const uint4 *src = reinterpret_cast<const uint4 *>(table + tok * ld_table);
    uint4 *dst = reinterpret_cast<uint4 *>(out + static_cast<int64_t>(row) * ldo);
    for (int v = threadIdx.x; v < hidden / 8; v += blockDim.x) dst[v] = src[v];
}

// Greedy sampling: one CTA per row; ties -> the LAST index (max_by semantics, see the header comment).
// NaN logits: total_cmp orders +NaN above +inf; restated by comparing the sign-magnitude integer key.
__device__ __forceinline__ int32_t total_order_key(float f) {
    int32_t b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
// Row argmax by one CTA over `vocab` logits at x; ties -> the LAST index.  Thread 0 returns the result.
template <typename T>
__device__ __forceinline__ void row_argmax(const T *__restrict__ x, int vocab, float *__restrict__ logits_f32_row,
                                           int32_t *s_key, int *s_idx, int32_t &best_key, int &best_idx) {
    best_key = INT32_MIN;
    best_idx = -1;
    const int nvec = vocab / 8;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[8];
        unpack8f<T>(ld_nc_v4(x + v * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int32_t k = total_order_key(f[j]);
            if (k >= best_key) {   // indices ascend within a thread: >= keeps the last maximum
                best_key = k;
                best_idx = v * 8 + j;
            }
        }
        if (logits_f32_row) {
            float4 *o = reinterpret_cast<float4 *>(logits_f32_row + v * 8);
            o[0] = make_float4(f[0], f[1], f[2], f[3]);
            o[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
    for (int i = nvec * 8 + threadIdx.x; i < vocab; i += blockDim.x) {
        const float f = static_cast<float>(x[i]);
        const int32_t k = total_order_key(f);
        if (k > best_key || (k == best_key && i > best_idx)) {
            best_key = k;
            best_idx = i;
        }
        if (logits_f32_row) logits_f32_row[i] = f;
    }
    auto better = [](int32_t k1, int i1, int32_t k2, int i2) { return k1 > k2 || (k1 == k2 && i1 > i2); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t k = __shfl_xor_sync(0xffffffffu, best_key, o);
        const int i = __shfl_xor_sync(0xffffffffu, best_idx, o);
        if (better(k, i, best_key, best_idx)) {
            best_key = k;
            best_idx = i;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) {
        s_key[warp] = best_key;
        s_idx[warp] = best_idx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w)
            if (better(s_key[w], s_idx[w], best_key, best_idx)) {
                best_key = s_key[w];
                best_idx = s_idx[w];
            }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) argmax_kernel(const T *__restrict__ logits, int64_t ld, int vocab,
                                                      uint32_t *__restrict__ token_out, float *__restrict__ logit_out,
                                                      float *__restrict__ logits_f32, int64_t ld_f32) {
    __shared__ int32_t s_key[32];
    __shared__ int s_idx[32];
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const T *x = logits + static_cast<int64_t>(row) * ld;
    int32_t best_key;
    int best_idx;
    row_argmax<T>(x, vocab, logits_f32 ? logits_f32 + static_cast<int64_t>(row) * ld_f32 : nullptr, s_key, s_idx, best_key, best_idx);
    if (threadIdx.x == 0) {
        token_out[row] = static_cast<uint32_t>(best_idx < 0 ? 0 : best_idx);
        if (logit_out) logit_out[row] = best_idx < 0 ? 0.f : static_cast<float>(x[best_idx]);
    }
}

// Greedy sampling over a vocabulary sharded across tensor-parallel ranks (vocab-parallel lm_head): every rank
// takes the argmax of its own slice of the logits, stores {key, global index, epoch} into EVERY rank's exchange
// area over NVLink peer memory as one 16-byte unit (flag and data travel together, as in the all-reduce), polls its
// own area for the other ranks' candidates and keeps the best under the same order (ties -> the last index): the
// result is the argmax of the full row, identical on every rank, and `vocab * 2` bytes of lm_head per rank become
// `vocab * 2 / world`.  Region header words: [16] epoch of the last completed exchange, [17] CTA arrival counter.
struct ArgmaxTpArgs {
    const void *logits;        // [T, rows] this rank's slice
    int64_t ld;
    int rows, row0;            // slice size and its first global index
    uint32_t *token_out;
    float *logit_out;
    unsigned char *region[8];  // every rank's region (own: local pointer)
    int64_t exch_offset;       // byte offset of the exchange area: [2 parities][8 ranks][128 rows] x 16 bytes
    int rank, world;
    unsigned long long timeout_ns;
};
constexpr int kArgmaxMaxRows = 128;

template <typename T>
__global__ void __launch_bounds__(1024) argmax_tp_kernel(const ArgmaxTpArgs a) {
    __shared__ int32_t s_key[32];
    __shared__ int s_idx[32];
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const T *x = static_cast<const T *>(a.logits) + static_cast<int64_t>(row) * a.ld;
    int32_t best_key;
    int best_idx;
    row_argmax<T>(x, a.rows, nullptr, s_key, s_idx, best_key, best_idx);
    if (threadIdx.x != 0) return;
    uint32_t *state = reinterpret_cast<uint32_t *>(a.region[a.rank]);
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(&state[16]) + 1;
    if (*reinterpret_cast<volatile uint32_t *>(&state[2]) != 0) return;   // a peer failed earlier
    const int64_t base = a.exch_offset + static_cast<int64_t>(e & 1u) * 8 * kArgmaxMaxRows * 16;
    const uint32_t gidx = static_cast<uint32_t>(a.row0 + (best_idx < 0 ? 0 : best_idx));
    for (int d = 0; d < a.world; ++d) {
        const int r = (a.rank + d) % a.world;
        unsigned char *dst = a.region[r] + base + (static_cast<int64_t>(a.rank) * kArgmaxMaxRows + row) * 16;
        asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(static_cast<uint32_t>(best_key)),
                     "r"(gidx), "r"(e), "r"(0u)
                     : "memory");
    }
    int32_t k_best = INT32_MIN;
    uint32_t i_best = 0;
    bool failed = false;
    const unsigned long long t0 = ops_timer_ns();
    for (int r = 0; r < a.world && !failed; ++r) {
        const unsigned char *src = a.region[a.rank] + base + (static_cast<int64_t>(r) * kArgmaxMaxRows + row) * 16;
        uint4 v;
        for (int spin = 0;; ++spin) {
            v = ld_volatile_v4(src);
            if (v.z == e) break;
            if ((spin & 63) == 63 && ops_timer_ns() - t0 > a.timeout_ns) {
                failed = true;
                break;
            }
        }
        if (failed) break;
        const int32_t k = static_cast<int32_t>(v.x);
        if (r == 0 || k > k_best || (k == k_best && v.y > i_best)) {   // ties -> the last (largest) global index
            k_best = k;
            i_best = v.y;
        }
    }
    if (failed) {
        atomicExch(&state[2], 1u);
    } else {
        a.token_out[row] = i_best;
        if (a.logit_out) a.logit_out[row] = __int_as_float(k_best ^ ((k_best >> 31) & 0x7fffffff));   // the key map is an involution
    }
    __threadfence();
    if (atomicAdd(&state[17], 1u) == gridDim.x - 1) {
        state[17] = 0;
        __threadfence();
        state[16] = e;
    }
}

inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// The small kernels of the step ask for the same (maximum) shared-memory carve-out as the GEMMs and the decode
// kernel they run between: an SM cannot change its L1 / shared split while CTAs of a neighbouring kernel are
// resident, and programmatic dependent launch makes exactly that the normal case.
void prefer_max_carveout_once() {
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev == di.ordinal) return;
    auto set = [](auto *k) { cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); };
    set(add_rmsnorm_kernel<__nv_bfloat16, 1>); set(add_rmsnorm_kernel<__nv_bfloat16, 2>);
    set(add_rmsnorm_kernel<__nv_bfloat16, 4>); set(add_rmsnorm_kernel<__nv_bfloat16, 8>);
    set(add_rmsnorm_kernel<__half, 1>); set(add_rmsnorm_kernel<__half, 2>);
    set(add_rmsnorm_kernel<__half, 4>); set(add_rmsnorm_kernel<__half, 8>);
    set(embedding_kernel<__nv_bfloat16>); set(embedding_kernel<__half>);
    set(argmax_kernel<__nv_bfloat16>); set(argmax_kernel<__half>);
    set(argmax_tp_kernel<__nv_bfloat16>); set(argmax_tp_kernel<__half>);
    cudaGetLastError();
    attr_dev = di.ordinal;
}

template <typename T>
cudaError_t launch_norm(const NormArgs &a, int threads, cudaStream_t stream) {
    switch (a.world) {
        case 1: return launch_ex(add_rmsnorm_kernel<T, 1>, dim3(a.T), dim3(threads), 0, stream, a);
        case 2: return launch_ex(add_rmsnorm_kernel<T, 2>, dim3(a.T), dim3(threads), 0, stream, a);
        case 4: return launch_ex(add_rmsnorm_kernel<T, 4>, dim3(a.T), dim3(threads), 0, stream, a);
        case 8: return launch_ex(add_rmsnorm_kernel<T, 8>, dim3(a.T), dim3(threads), 0, stream, a);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// p2p region geometry shared with allreduce_p2p.cu (kArHeaderBytes = 4096, two staging slots, then
// recv[2 slots][8 source ranks][2 planes of slot bytes]).
void p2p_layout(int64_t slot_bytes, int64_t *recv_offset, int64_t *slot_stride, int64_t *recv_stride) {
    const int64_t s = (slot_bytes + 255) / 256 * 256;
    *recv_stride = 2 * s;
    *slot_stride = 8 * (*recv_stride);
    *recv_offset = 4096 + 2 * s;
}

void norm_args_set_p2p(NormArgs &a, void *p2p_region, int64_t p2p_slot_bytes, int world) {
    a.world = world > 1 ? world : 1;
    a.region = static_cast<unsigned char *>(p2p_region);
    if (a.world > 1) {
        p2p_layout(p2p_slot_bytes, &a.recv_offset, &a.slot_stride, &a.recv_stride);
        a.plane_bytes = static_cast<int64_t>(a.T) * a.hidden * 2;
        a.timeout_ns = 30000000000ull;
    }
}

bool launch_add_rmsnorm(const void *x, const void *residual, void *residual_out, const void *weight, void *normed,
                        int64_t T, int64_t hidden, int64_t ldx, int64_t ldr, int64_t ldro, int64_t ldn, float eps,
                        uint32_t dtype, void *p2p_region, int64_t p2p_slot_bytes, int world, cudaStream_t stream) {
    if (T <= 0) return true;
    if (dtype > 1 || hidden <= 0 || hidden % 8 || hidden > 8 * 512 * kNormMaxVec || (normed && !weight) ||
        !al16(x) || !al16(residual) || !al16(residual_out) || !al16(weight) || !al16(normed) || ldx % 8 || ldr % 8 ||
        ldro % 8 || ldn % 8 || (world > 1 && (!p2p_region || static_cast<int64_t>(T) * hidden * 2 > p2p_slot_bytes))) {
        set_status(kErrInvalidArg,
                   "atoma_add_rmsnorm: dtype 0=f16 1=bf16, hidden a multiple of 8 (<= %d), 16-byte aligned pointers and "
                   "row strides (hidden %lld, dtype %u)",
                   8 * 512 * kNormMaxVec, static_cast<long long>(hidden), dtype);
        return false;
    }
    NormArgs a{};
    a.x = x; a.residual = residual; a.residual_out = residual_out; a.weight = weight; a.normed = normed;
    a.ldx = ldx; a.ldr = ldr; a.ldro = ldro; a.ldn = ldn;
    a.hidden = static_cast<int>(hidden);
    a.T = static_cast<int>(T);
    a.eps = eps;
    norm_args_set_p2p(a, p2p_region, p2p_slot_bytes, world);
    const int nvec = a.hidden / 8;
    // 256 threads whenever they can hold the row (hidden <= 8192): the same thread -> vector assignment and the same
    // reduction order as the fused tail of the row-parallel projections (gemm_skinny.cu), hence the same bits
    int threads = std::min(nvec > 256 * kNormMaxVec ? 512 : 256, (nvec + 31) / 32 * 32);
    prefer_max_carveout_once();
    a.const_weight = (launch_flags() & kLaunchConstWeights) ? 1 : 0;
    a.trace = a.T <= kStepTraceCtas ? next_step_trace_slot(a.world > 1 ? "add_rmsnorm recv" : "add_rmsnorm") : nullptr;
    cudaError_t e = dtype == 1 ? launch_norm<__nv_bfloat16>(a, threads, stream) : launch_norm<__half>(a, threads, stream);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "add_rmsnorm_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("add_rmsnorm_kernel");
}

bool launch_embedding(const uint32_t *tokens, const void *table, void *out, int64_t T, int64_t hidden, int64_t vocab,
                      int64_t ld_table, int64_t ldo, uint32_t dtype, cudaStream_t stream) {
    if (T <= 0) return true;
    if (dtype > 1 || hidden % 8 || !al16(table) || !al16(out) || ld_table % 8 || ldo % 8 || !tokens) {
        set_status(kErrInvalidArg, "atoma_embedding: 16-bit dtype, hidden a multiple of 8, 16-byte aligned rows");
        return false;
    }
    prefer_max_carveout_once();
    cudaError_t e;
    if (dtype == 1)
        e = launch_ex(embedding_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(T)), dim3(256), 0, stream, tokens,
                      static_cast<const __nv_bfloat16 *>(table), static_cast<__nv_bfloat16 *>(out), static_cast<int>(hidden),
                      ld_table, ldo, vocab);
    else
        e = launch_ex(embedding_kernel<__half>, dim3(static_cast<unsigned>(T)), dim3(256), 0, stream, tokens,
                      static_cast<const __half *>(table), static_cast<__half *>(out), static_cast<int>(hidden), ld_table, ldo,
                      vocab);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "embedding_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("embedding_kernel");
}

bool launch_argmax(const void *logits, int64_t T, int64_t vocab, int64_t ld, uint32_t dtype, uint32_t *token_out,
                   float *logit_out, float *logits_f32, int64_t ld_f32, cudaStream_t stream) {
    if (T <= 0) return true;
    if (dtype > 1 || vocab <= 0 || !logits || !token_out || !al16(logits) || ld % 8 || (logits_f32 && (!al16(logits_f32) || ld_f32 % 4))) {
        set_status(kErrInvalidArg, "atoma_sample_greedy: 16-bit logits with 16-byte aligned rows");
        return false;
    }
    prefer_max_carveout_once();
    cudaError_t e;
    if (dtype == 1)
        e = launch_ex(argmax_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(T)), dim3(1024), 0, stream,
                      static_cast<const __nv_bfloat16 *>(logits), ld, static_cast<int>(vocab), token_out, logit_out, logits_f32, ld_f32);
    else
        e = launch_ex(argmax_kernel<__half>, dim3(static_cast<unsigned>(T)), dim3(1024), 0, stream,
                      static_cast<const __half *>(logits), ld, static_cast<int>(vocab), token_out, logit_out, logits_f32, ld_f32);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "argmax_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("argmax_kernel");
}

bool launch_argmax_tp(const void *logits, int64_t T, int64_t rows, int64_t row0, int64_t ld, uint32_t dtype,
                      uint32_t *token_out, float *logit_out, void *const *regions, int rank, int world, int64_t slot_bytes,
                      cudaStream_t stream) {
    if (T <= 0) return true;
    // the exchange area is the first staging slot of the region (unused by the fused push / receive all-reduce)
    if (dtype > 1 || rows <= 0 || !logits || !token_out || !al16(logits) || ld % 8 || T > kArgmaxMaxRows || !regions ||
        world < 2 || world > 8 || rank < 0 || rank >= world || slot_bytes < 2 * 8 * kArgmaxMaxRows * 16) {
        set_status(kErrInvalidArg,
                   "atoma_sample_greedy_tp: 16-bit logits with 16-byte aligned rows, T <= %d, world 2..8, slot_bytes >= %d",
                   kArgmaxMaxRows, 2 * 8 * kArgmaxMaxRows * 16);
        return false;
    }
    prefer_max_carveout_once();
    ArgmaxTpArgs a{};
    a.logits = logits; a.ld = ld;
    a.rows = static_cast<int>(rows); a.row0 = static_cast<int>(row0);
    a.token_out = token_out; a.logit_out = logit_out;
    for (int r = 0; r < 8; ++r) a.region[r] = static_cast<unsigned char *>(regions[r < world ? r : rank]);
    a.exch_offset = 4096;   // kArHeaderBytes: staging slot 0
    a.rank = rank; a.world = world;
    a.timeout_ns = 30000000000ull;
    cudaError_t e = dtype == 1 ? launch_ex(argmax_tp_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(T)), dim3(1024), 0, stream, a)
                               : launch_ex(argmax_tp_kernel<__half>, dim3(static_cast<unsigned>(T)), dim3(1024), 0, stream, a);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "argmax_tp_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("argmax_tp_kernel");
}

}  // namespace atoma
