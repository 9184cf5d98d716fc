// Residual add + RMSNorm of ONE token row by one CTA (models/src/llama.rs:398-410), shared by the standalone
// kernel (llama_ops.cu: add_rmsnorm_kernel) and by the tail of the row-parallel projections (gemm_skinny.cu:
// the CTAs that finish last normalise the rows of the output they have just completed — SURVEY §8 rows f3, f4).
//
//   x        = the addend: a [T, hidden] tensor, or (W > 1) the rank-order fp32 sum of the W addends that the
//              ranks pushed into this rank's receive buffers, rounded to the io type (multi_gpu.rs:141-179), or 0
//   h        = io(x + residual)            (the adds of Block::forward are in the model dtype)
//   normed   = io(h * rsqrt(mean(h^2) + eps) * weight)   statistics and products in fp32 (candle RmsNorm)
#pragma once

#include "common.cuh"

namespace atoma {

constexpr int kNormMaxVec = 4;   // 16-byte vectors per thread: hidden <= 8 * blockDim * 4

struct NormArgs {
    const void *x;         // [T, hidden] addend (projection output) or null
    const void *residual;  // [T, hidden] or null (first layer: h = x)
    void *residual_out;    // [T, hidden] h = x + residual (may alias residual or x), or null
    const void *weight;    // [hidden]
    void *normed;          // [T, hidden] rmsnorm(h) * weight, or null (only the add is wanted)
    int64_t ldx, ldr, ldro, ldn;
    int hidden, T;
    float eps;
    // receive side of the tensor-parallel all-reduce (world > 1)
    unsigned char *region;   // local region
    int64_t recv_offset, slot_stride, recv_stride, plane_bytes;
    int world;
    unsigned long long timeout_ns;
    unsigned long long *trace;   // diagnostics (step timeline): [cta][8] = {-, start, dependency resolved, end}
    int const_weight;            // the norm weight is immutable (model weight): load it before the grid dependency resolves
};

// Receive-buffer layout of the fused all-reduce (one plane of a source rank's [T, N] addend; the second plane lies
// plane_bytes further): 32-column blocks, inside a block row-major 64-byte rows — so that the [T, 32k] tile a GEMM
// CTA owns is ONE contiguous run of the peer's memory and a warp's 32 x 16-byte stores coalesce into full NVLink
// packets (row-major [T, N] made every 16-byte store its own packet: 34 us per 8-rank all-reduce, measured).
__device__ __forceinline__ int64_t p2p_unit_offset(int row, int col, int T) {
    return ((static_cast<int64_t>(col >> 5) * T + row) * 4 + ((col & 31) >> 3)) * 16;
}

namespace normrow {

template <typename T>
__device__ __forceinline__ void unpack8f(const uint4 &v, float (&f)[8]) {
    const T *p = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = static_cast<float>(p[i]);
}
template <typename T>
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
    uint4 o;
    o.x = pack2<T>(f[0], f[1]);
    o.y = pack2<T>(f[2], f[3]);
    o.z = pack2<T>(f[4], f[5]);
    o.w = pack2<T>(f[6], f[7]);
    return o;
}
template <typename T>
__device__ __forceinline__ float round_to(float f) {
    return static_cast<float>(static_cast<T>(f));
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
// L2-coherent load: the addend may have been written by other CTAs of the same grid (GEMM tail)
__device__ __forceinline__ uint4 ld_cg_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
// Sum over the CTA, the same bits in every thread (fixed order).  Contains one __syncthreads().
__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

}  // namespace normrow

// This thread's vectors of the norm weight (call before griddepcontrol.wait when the weight is immutable).
__device__ __forceinline__ void norm_weight_preload(const NormArgs &a, uint4 (&wv)[kNormMaxVec]) {
    const int nvec = a.hidden / 8;
#pragma unroll
    for (int u = 0; u < kNormMaxVec; ++u) {
        const int v = threadIdx.x + u * blockDim.x;
        wv[u] = (v < nvec && a.weight) ? *reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(a.weight) + v * 16)
                                       : make_uint4(0, 0, 0, 0);
    }
}

// One row by the whole CTA (every thread must call it; thread t owns the 16-byte vectors t, t + blockDim, ...).
// `e` = the all-reduce's epoch (W > 1).  `red`: >= 32 floats of shared memory, reusable after the call returns and
// a __syncthreads().  Returns true if a peer failed to arrive in time (W > 1; CTA-uniform) — nothing was normalised.
// `wpre`: the thread's weight vectors from norm_weight_preload(), or null (loaded here).
template <typename T, int W>
__device__ __forceinline__ bool add_rmsnorm_row(const NormArgs &a, int row, uint32_t e, float *red,
                                                const uint4 *wpre = nullptr) {
    using namespace normrow;
    const int nvec = a.hidden / 8;
    bool failed = false;
    float h[kNormMaxVec][8];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < kNormMaxVec; ++u) {
        const int v = threadIdx.x + u * blockDim.x;
        if (v >= nvec) break;
        float x[8];
        if (W > 1) {
            // poll the W addends of vector i (8-byte units {data, epoch}); add in rank order, round once
            const unsigned char *src = a.region + a.recv_offset + (e & 1u) * a.slot_stride + p2p_unit_offset(row, v * 8, a.T);
            uint4 lo[W > 1 ? W : 1], hi[W > 1 ? W : 1];
            bool done[W > 1 ? W : 1];
#pragma unroll
            for (int r = 0; r < W; ++r) done[r] = false;
            const unsigned long long t0 = timer_ns();
            for (int spin = 0;; ++spin) {
                bool all = true;
#pragma unroll
                for (int r = 0; r < W; ++r) {
                    if (!done[r]) {
                        lo[r] = ld_volatile_v4(src + r * a.recv_stride);
                        hi[r] = ld_volatile_v4(src + r * a.recv_stride + a.plane_bytes);
                        done[r] = lo[r].y == e && lo[r].w == e && hi[r].y == e && hi[r].w == e;
                        all = all && done[r];
                    }
                }
                if (all) break;
                if ((spin & 63) == 63 && timer_ns() - t0 > a.timeout_ns) {
                    failed = true;
                    break;
                }
            }
            if (failed) break;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int r = 0; r < W; ++r) {
                float f[8];
                unpack8f<T>(make_uint4(lo[r].x, lo[r].z, hi[r].x, hi[r].z), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = round_to<T>(acc[j]);   // the all-reduce result is in the io type
        } else if (a.x) {
            unpack8f<T>(ld_cg_v4(static_cast<const T *>(a.x) + static_cast<int64_t>(row) * a.ldx + v * 8), x);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = 0.f;
        }
        if (a.residual) {
            float r[8];
            unpack8f<T>(ld_cg_v4(static_cast<const T *>(a.residual) + static_cast<int64_t>(row) * a.ldr + v * 8), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = round_to<T>(x[j] + r[j]);   // the add is in the model dtype
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[u][j] = x[j];
            ss += x[j] * x[j];
        }
        if (a.residual_out)
            *reinterpret_cast<uint4 *>(static_cast<T *>(a.residual_out) + static_cast<int64_t>(row) * a.ldro + v * 8) = pack8f<T>(x);
    }
    int any_failed = 0;
    if (W > 1) any_failed = __syncthreads_or(failed ? 1 : 0);
    const float total = block_sum(ss, red);
    if (a.normed && !any_failed) {
        const float scale = rsqrtf(total / static_cast<float>(a.hidden) + a.eps);
#pragma unroll
        for (int u = 0; u < kNormMaxVec; ++u) {
            const int v = threadIdx.x + u * blockDim.x;
            if (v >= nvec) break;
            float w[8], o[8];
            unpack8f<T>(wpre ? wpre[u] : *reinterpret_cast<const uint4 *>(static_cast<const T *>(a.weight) + v * 8), w);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = scale * h[u][j] * w[j];
            *reinterpret_cast<uint4 *>(static_cast<T *>(a.normed) + static_cast<int64_t>(row) * a.ldn + v * 8) = pack8f<T>(o);
        }
    }
    return any_failed != 0;
}

// Fill the tensor-parallel receive fields of `a` (defined in llama_ops.cu).
void norm_args_set_p2p(NormArgs &a, void *p2p_region, int64_t p2p_slot_bytes, int world);

}  // namespace atoma
