// Decode-time ("skinny") projection GEMM  y[T,N] = x[T,K] . w[N,K]^T  for T <= 128 on sm_100a, with the
// fused epilogues of the decode step.  Call sites in the reference: q/k/v/o projections
// models/src/llama.rs:269-271, 311; MLP gate/up/down :362-366; lm_head :476 — all candle `Linear` ->
// cuBLAS there (no reference kernel to follow; bar = cuBLAS 12.9 on the same box).
//
// Why a second GEMM kernel next to gemm_bf16.cu (stream-K): at T <= 128 the GEMM is a pure weight
// stream (algorithmic bytes 2 (N K + T K + T N), ~T flop/byte) and what separates a kernel from the
// HBM roofline is latency, not bandwidth: round 1's stream-K kernel streamed W at 6.9 TB/s but lost
// ~6.5 us per launch in its split reduction (fp32 partials through L2, fence, arrival counter,
// last-arriver pass).  Here no tile is ever split:
//  * the ACTIVATIONS are the M operand of tcgen05.mma (M = 64 for T <= 64 — rows >= T of the operand
//    are whatever shared memory holds, they only reach accumulator rows nobody reads — else 128) and
//    the WEIGHT tile is the N operand, whose width may be any multiple of 8 up to 256: the host picks
//    the width so that the tiles fill one wave of the SMs (qkv of Llama-3.1-8B: 144 tiles of 48
//    columns, o_proj / down_proj: 128 x 32, gate_up: 138 x 208, lm_head: 573 x 224 in 4 waves), each
//    CTA streams the full K extent of its tile(s) and owns the finished accumulator row block;
//  * a weight tile is `halves` (1 or 2) runs of `np` consecutive rows, `half_stride` rows apart,
//    inside row group g (5-D tensor map {64 k, row, half, k block, group}; one TMA op moves
//    [kbs k blocks][halves][np rows][128 B]): with halves = 2 an epilogue thread holds BOTH elements of
//    a RoPE pair (d, d + D/2) or of a SwiGLU pair (gate_i, up_i), so RoPE + KV-cache scatter
//    (llama.rs:218-304 + flash_attention.rs:361) and silu(gate) * up (llama.rs:364,
//    f3) are applied to the accumulator before it is stored — no extra kernel, no extra pass;
//  * tensor-parallel row-parallel projections (o_proj, down_proj; multi_gpu.rs:141-179) can PUSH the
//    bf16 partial tile straight into every peer's receive buffer over NVLink from the epilogue (8-byte
//    units carrying their own epoch flag, the protocol of allreduce_p2p.cu) — the all-reduce kernel
//    that followed the GEMM shrinks to the receive side, which is fused with residual + RMSNorm
//    (llama_ops.cu), SURVEY §8 f4;
//  * warp 0 = TMA producer (W evict-first, x evict-last; under programmatic dependent launch with
//    immutable weights the ring is filled with W before the grid dependency resolves), warp 1 =
//    tcgen05.mma issuer (accumulators double-buffered in TMEM so a multi-wave CTA's epilogue overlaps
//    its next tile), warps 2-5 = epilogue (tcgen05.ld, thread = token row).
#include "common.cuh"
#include "skinny_args.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace atoma {
namespace {

constexpr int kSkThreads = 256;             // warps: 0 W producer, 1 MMA, 2-5 epilogue, 6 idle, 7 x producer
constexpr int kSkMaxStages = 12;
constexpr int kSkSmemBudget = 200 * 1024;   // ring of a CTA that owns its SM; + tail slack + 1 KiB alignment < 227 KiB
// "Half footprint": ring + slack <= half of an SM's shared memory and <= 256 tensor-memory columns, so that TWO
// CTAs fit on an SM — not for this kernel's own sake (one wave of tiles still gets one CTA per SM) but so that the
// NEXT kernel of the stream (programmatic dependent launch) becomes resident and streams its weights while this
// one drains: HBM does not idle across the kernel boundary.
constexpr int kSkHalfBudget = 100 * 1024;
// One thread issues only ~5 TMA tensor ops per microsecond whatever their size (measured: round 1's
// tma_probe; here 0.25 us per stage with two ops per stage from one thread), so a stage is made of FEW,
// LARGE ops — up to 96 KiB per stage, the weight box split over two issuing warps, x from a third.
constexpr int kSkStageTarget = 96 * 1024;
constexpr int kSkDefaultBudget = kSkSmemBudget;

__device__ __forceinline__ unsigned long long sk_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

// 32 lanes x 8 consecutive 32-bit TMEM columns -> 8 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void st_relaxed_sys_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}

// fp32 accumulator -> the io type and back: every fused epilogue starts from the value the unfused
// path would have stored (the reference materialises each projection in the model dtype).
template <typename T>
__device__ __forceinline__ float round_io(float f);
template <>
__device__ __forceinline__ float round_io<__nv_bfloat16>(float f) {
    return __bfloat162float(__float2bfloat16_rn(f));
}
template <>
__device__ __forceinline__ float round_io<__half>(float f) {
    return __half2float(__float2half_rn(f));
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack2<T>(f[0], f[1]);
    o.y = pack2<T>(f[2], f[3]);
    o.z = pack2<T>(f[4], f[5]);
    o.w = pack2<T>(f[6], f[7]);
    return o;
}
template <typename T>
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    const T *p = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = static_cast<float>(p[i]);
}

// Store 8 consecutive output elements of one row, `valid` of them inside the matrix.
template <typename T>
__device__ __forceinline__ void store8(T *dst, const uint4 &v, int valid, bool vec_ok) {
    if (valid >= 8 && vec_ok) {
        *reinterpret_cast<uint4 *>(dst) = v;
    } else {
        const T *p = reinterpret_cast<const T *>(&v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < valid) dst[i] = p[i];
    }
}

// Per-row state of an epilogue thread (thread = token row).
template <typename T>
struct EpiRow {
    int row;
    bool row_ok;
    T *y;                 // this row of the output
    bool y_vec;
    const T *cos_r, *sin_r;
    int64_t cache_row;
    uint32_t epoch;
    int slot_sel;
    unsigned char *stg;   // push epilogue: shared-memory staging of the tile
};

template <typename T, int M, int EPI>
__device__ __forceinline__ EpiRow<T> epi_row_init(const SkinnyArgs &a, int warp, int lane, unsigned char *stg) {
    EpiRow<T> r;
    const int quad = warp & 3;   // the TMEM lane quadrant this warp may read
    // M = 64: accumulator row r lives in lane (r % 16) + 32 (r / 16); M = 128: row r in lane r.
    r.row = (M == 64) ? quad * 16 + lane : quad * 32 + lane;
    r.row_ok = (M == 64 ? lane < 16 : true) && r.row < a.T;
    r.y = static_cast<T *>(a.y) + static_cast<int64_t>(r.row_ok ? r.row : 0) * a.ldy;
    r.y_vec = (a.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
    r.cos_r = r.sin_r = nullptr;
    r.cache_row = -1;
    if (EPI == kEpiRope && r.row_ok) {
        const int64_t pos = a.positions[r.row];
        r.cos_r = static_cast<const T *>(a.cos) + pos * a.rot_stride;
        r.sin_r = static_cast<const T *>(a.sin) + pos * a.rot_stride;
        const int64_t slot = a.slot_mapping ? a.slot_mapping[r.row] : -1;   // < 0: padded token (cache_manager.cu:151-155)
        if (slot >= 0 && a.k_cache)
            r.cache_row = (slot / a.block_size) * a.block_stride + (slot % a.block_size) * a.num_kv_heads * (2 * a.J);
    }
    r.epoch = 0;
    r.slot_sel = 0;
    r.stg = stg;
    if (EPI == kEpiPush) {
        r.epoch = a.push_state[0] + 1;   // this all-reduce's epoch (advanced by the receive-side kernel)
        r.slot_sel = static_cast<int>(r.epoch & 1u);
    }
    return r;
}

// The epilogue of ONE finished tile (group g, first row p0 of the group) whose accumulator starts at TMEM address
// `tacc` (lane quadrant included), by the four epilogue warps.
template <typename T, int M, int EPI>
__device__ __forceinline__ void epilogue_tile(const SkinnyArgs &a, const EpiRow<T> &er, uint32_t tacc, int g, int p0, int warp,
                                              int lane) {
    const int J = a.J, np = a.np;
    const int row = er.row;
    const bool row_ok = er.row_ok;
    T *y = er.y;
    const bool y_vec = er.y_vec;
    const T *cos_r = er.cos_r, *sin_r = er.sin_r;
    const int64_t cache_row = er.cache_row;
    const uint32_t epoch = er.epoch;
    const int slot_sel = er.slot_sel;
    unsigned char *stg = er.stg;
    // 8 consecutive columns of the tile: the sum of the nacc partial accumulators (fixed order)
    auto load8 = [&](int col, uint32_t (&v)[8]) {
        tmem_ld_32x32b_x8(tacc + col, v);
        for (int i = 1; i < a.nacc; ++i) {
            uint32_t w[8];
            tmem_ld_32x32b_x8(tacc + i * a.acc_stride + col, w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
    };

    if (EPI == kEpiPlain || EPI == kEpiPush) {
        // columns [p0, p0 + np) of the output, 8 at a time, two TMEM loads in flight
        for (int c = 0; c < np; c += 16) {
            uint32_t v0[8], v1[8];
            load8(c, v0);
            if (c + 8 < np) load8(c + 8, v1);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j0 = p0 + c + 8 * u;
                if (c + 8 * u >= np || j0 >= J || !row_ok) continue;
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(u == 0 ? v0[i] : v1[i]);
                const uint4 o = pack8<T>(f);
                if (EPI == kEpiPlain) {
                    store8<T>(y + j0, o, J - j0, y_vec);
                } else {
                    // tensor parallel: stage the tile in shared memory; pushed below, coalesced
                    *reinterpret_cast<uint4 *>(stg + (static_cast<size_t>(row) * np + c + 8 * u) * sizeof(T)) = o;
                }
            }
        }
        if (EPI == kEpiPush) {
            // This rank's addend goes to recv[slot][rank] of EVERY rank (own included) as 8-byte units
            // {4 bytes of data, epoch}: vector (row, col) = two 16-byte stores {w0,e,w1,e} {w2,e,w3,e}, one per
            // plane.  Consecutive lanes take consecutive units of the peer's block layout (norm_row.cuh), so a
            // warp writes 512 contiguous bytes per store instruction.
            named_bar_sync(2, 128);   // the four epilogue warps: the tile is staged
            const int tid_e = (warp - 2) * 32 + lane, vpr = np >> 3, nvec_t = a.T * vpr;
            const bool blocked = (np & 31) == 0;
            const int per_blk = a.T * 4;
            const int64_t base = a.push_recv_offset + slot_sel * a.push_slot_stride + a.push_rank * a.push_recv_stride;
            for (int idx = tid_e; idx < nvec_t; idx += 128) {
                int r_, vec;
                if (blocked) {
                    const int blk = idx / per_blk, rem = idx - blk * per_blk;
                    r_ = rem >> 2;
                    vec = blk * 4 + (rem & 3);
                } else {
                    r_ = idx / vpr;
                    vec = idx - r_ * vpr;
                }
                const int col = p0 + vec * 8;
                if (col >= J) continue;
                const uint4 o = *reinterpret_cast<const uint4 *>(stg + (static_cast<size_t>(r_) * np + vec * 8) * sizeof(T));
                const int64_t off = base + p2p_unit_offset(r_, col, a.T);
#pragma unroll 1
                for (int d = 0; d < a.push_world; ++d) {
                    const int r = (a.push_rank + d) % a.push_world;
                    unsigned char *dst = a.push_region[r] + off;
                    st_relaxed_sys_v4(dst, o.x, epoch, o.y, epoch);
                    st_relaxed_sys_v4(dst + a.push_plane_bytes, o.z, epoch, o.w, epoch);
                }
            }
            named_bar_sync(2, 128);   // the staging buffer is free for the CTA's next tile
        }
    } else {
        // paired tile: accumulator columns [0, np) = first half (rows p0.. of the group), [np, 2 np) =
        // second half (the same rows `half_stride` further down)
        const bool is_q = g < a.num_heads, is_k = !is_q && g < a.num_heads + a.num_kv_heads;
        for (int c = 0; c < np; c += 8) {
            uint32_t lo[8], hi[8];
            load8(c, lo);
            load8(np + c, hi);
            tmem_ld_wait();
            const int j0 = p0 + c;
            if (j0 >= J || !row_ok) continue;
            float f1[8], f2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                f1[i] = round_io<T>(__uint_as_float(lo[i]));
                f2[i] = round_io<T>(__uint_as_float(hi[i]));
            }
            if (EPI == kEpiSilu) {
                // silu(gate) * up (llama.rs:364): fp32 on the model-dtype projections, one rounding
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = (f1[i] / (1.f + __expf(-f1[i]))) * f2[i];
                store8<T>(y + j0, pack8<T>(o), J - j0, y_vec);
            } else {
                // RoPE (non-interleaved, candle_nn::rotary_emb::rope; llama.rs:218-251) on q and k heads;
                // v heads pass through.  Same arithmetic as rope_reshape_and_cache_kernel.
                float o1[8], o2[8];
                if (is_q || is_k) {
                    float fc[8], fs[8];
                    unpack8<T>(*reinterpret_cast<const uint4 *>(cos_r + j0), fc);
                    unpack8<T>(*reinterpret_cast<const uint4 *>(sin_r + j0), fs);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        o1[i] = __fmaf_rn(f1[i], fc[i], -__fmul_rn(f2[i], fs[i]));
                        o2[i] = __fmaf_rn(f1[i], fs[i], __fmul_rn(f2[i], fc[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        o1[i] = f1[i];
                        o2[i] = f2[i];
                    }
                }
                const uint4 w1 = pack8<T>(o1), w2 = pack8<T>(o2);
                T *dst = y + static_cast<int64_t>(g) * (2 * J) + j0;
                store8<T>(dst, w1, J - j0, y_vec);
                store8<T>(dst + J, w2, J - j0, y_vec);
                if (!is_q && cache_row >= 0) {
                    // reshape_and_cache_flash (cache_manager.cu:140-170): k_cache[slot / bs, slot % bs, head, :]
                    T *cache = static_cast<T *>(is_k ? a.k_cache : a.v_cache);
                    const int kvh = g - a.num_heads - (is_k ? 0 : a.num_kv_heads);
                    T *cd = cache + cache_row + static_cast<int64_t>(kvh) * (2 * J) + j0;
                    store8<T>(cd, w1, J - j0, true);
                    store8<T>(cd + J, w2, J - j0, true);
                }
            }
        }
    }
}

template <typename T, int M, int EPI>
__global__ void __launch_bounds__(kSkThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                   const SkinnyArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[kSkMaxStages], empty[kSkMaxStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_ticket;
    __shared__ uint32_t s_tail_state[2];
    __shared__ float s_red[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long t_start = a.trace ? sk_timer_ns() : 0ull;

    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full[i], 2);   // one arrive.expect_tx from the W producer, one from the x producer
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);   // one arrival per epilogue warp
        }
        fence_mbar_init();
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
    }
    // The mbarriers are ready once this barrier releases; the tensor-memory allocation (which can block while a
    // co-resident CTA of the previous kernel still holds its columns) is awaited only by the warps that use it
    // (named barrier 1), so the producers start streaming W at once.
    __syncthreads();
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, a.tmem_cols);
        tc_fence_before_sync();
    }
    if (warp >= 1 && warp <= 5) {
        named_bar_sync(1, 5 * 32);
        tc_fence_after_sync();
    }
    const uint32_t tmem_base = (warp >= 1 && warp <= 5) ? tmem_base_smem : 0u;

    const int my_tiles = (a.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total = my_tiles * a.chunks;   // (tile, k chunk) stages this CTA streams

    // The producer and MMA warps run their loops CONVERGED, all 32 lanes, on warp-uniform values; only the
    // asynchronous instructions themselves sit under elect.sync.  (Inside an `if (lane == 0)` region the
    // compiler keeps every operand in vector registers and wraps each UTMALDG / UTCHMMA — which take uniform
    // registers — in an R2UR + ELECT waterfall loop: ~70 cycles per tcgen05.mma and ~0.2 us per TMA op,
    // measured; that, not HBM, bounded the first version of this kernel.)
    if (warp == 0 || warp == 7) {
        // ============================ TMA producers ============================
        // warp 0: the weight op(s) of every stage; warp 7: the x op.  Each arrives on the stage's full
        // barrier with its own byte count.
        const bool is_w = warp == 0;
        // Release the dependent grid: at once, or (trigger_late) only when this CTA issues its last weight load —
        // a dependent grid that sits in the hardware queue for tens of microseconds waiting for SMs was measured
        // to be placed several microseconds late once they free up (scripts/trace_step.py).
        if (is_w && !a.trigger_late) pdl_launch_dependents();
        const int w_op_bytes = a.w_bytes / a.w_ops;
        // Under programmatic dependent launch with immutable weights the ring is filled with W while the
        // previous kernel drains; x (its output) only after the grid dependency has resolved.
        const int pre = (is_w && a.prefetch_w) ? min(a.stages, total) : 0;
        if (pre == 0) pdl_wait();
        if (a.trace && !is_w && lane == 0) a.trace[blockIdx.x * 8 + 6] = sk_timer_ns();   // x: grid dependency resolved
        int stage = 0;
        uint32_t phase = 0;
        int n = 0;
        // this CTA's share of the next kernel's prefetch region (W producer only)
        long long pf_i = 0, pf_end = 0;
        int pf_step = 0;
        if (is_w && a.pf_bytes > 0) {
            const long long pieces = l2_prefetch_pieces(a.pf_bytes);
            pf_i = pieces * blockIdx.x / gridDim.x;
            pf_end = pieces * (blockIdx.x + 1) / gridDim.x;
            if (a.pf_spread) pf_step = static_cast<int>((pf_end - pf_i + total - 1) / max(total, 1));
        }
        // no integer division inside the stage loop (a runtime division costs this warp hundreds of
        // cycles — more than a whole stage of MMAs)
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int g = tile / a.tiles_per_group, p0 = (tile - g * a.tiles_per_group) * a.np;
            for (int ch = 0; ch < a.chunks; ++ch, ++n) {
                if (n == pre && pre != 0) {
                    pdl_wait();
                    if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + 0] = t_start;
                }
                mbar_wait(&empty[stage], phase ^ 1);   // passes at once during the first trip round the ring
                unsigned char *dst = ring + stage * a.stage_bytes;
                if (is_w && a.trigger_late && n == max(total - a.trigger_late, 0)) pdl_launch_dependents();
                if (a.trace && is_w && lane == 0 && n == 0) {
                    unsigned smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    a.trace[blockIdx.x * 8 + 4] = smid;   // which SM ran this CTA (hand-over analysis: scripts/trace_step.py)
                }
                if (elect_one_sync()) {
                    // W is read exactly once by exactly one CTA: evict-first.  x is re-read by every CTA: keep it.
                    if (is_w) {
                        if (a.debug_skip & 2) {
                            mbar_arrive(&full[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full[stage], a.w_bytes);
                            for (int o = 0; o < a.w_ops; ++o)
                                tma_load_5d(dst + o * w_op_bytes, &map_w, &full[stage], 0, p0, 0, ch * a.kbs + o * a.wkb, g,
                                            kL2EvictFirst);
                        }
                    } else {
                        if (a.debug_skip & 1) {
                            mbar_arrive(&full[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full[stage], a.x_bytes);
                            tma_load_3d(dst + a.w_bytes, &map_x, &full[stage], 0, 0, ch * a.kbs, kL2EvictLast);
                        }
                    }
                }
                __syncwarp();
                if (pf_step > 0 && pf_i < pf_end) {   // warp-uniform
                    const long long e = min(pf_end, pf_i + pf_step);
                    if (elect_one_sync()) l2_prefetch_range(a.pf_ptr, a.pf_bytes, pf_i, e);
                    __syncwarp();
                    pf_i = e;
                }
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        if (pf_i < pf_end) {   // warp-uniform
            if (elect_one_sync()) l2_prefetch_range(a.pf_ptr, a.pf_bytes, pf_i, pf_end);
            __syncwarp();
        }
        if (pre != 0 && n <= pre) pdl_wait();   // everything fitted in the prefetch: still order after the predecessor
        if (a.trace && pre == 0 && is_w && lane == 0) a.trace[blockIdx.x * 8 + 0] = t_start;
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        pdl_wait();
        const uint32_t idesc = umma_idesc_f16(M, a.n_tile, a.is_bf16 != 0);
        const int w_slab = a.n_tile * 128, x_slab = a.t_box * 128;
        const uint32_t ring_u32 = smem_u32(ring);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = a.nbuf == 2 ? (it & 1) : 0, use = a.nbuf == 2 ? (it >> 1) : it;
            mbar_wait(&tempty[buf], (use & 1) ^ 1);   // the epilogue drained this accumulator
            tc_fence_after_sync();
            // `nacc` accumulators per tile, used round-robin (diagnostics; 1 by default)
            const uint32_t tmem_d0 = tmem_base + buf * (a.nacc * a.acc_stride);
            int step = 0;
            for (int ch = 0; ch < a.chunks; ++ch) {
                mbar_wait(&full[stage], phase);
                tc_fence_after_sync();
                if (a.trace && it == 0 && ch == 0 && lane == 0) a.trace[blockIdx.x * 8 + 1] = sk_timer_ns();
                if (a.trace && it == my_tiles - 1 && ch == a.chunks - 1 && lane == 0) a.trace[blockIdx.x * 8 + 7] = sk_timer_ns();
                const uint32_t base = ring_u32 + stage * a.stage_bytes;
                const int nkb = min(a.kbs, a.kb_total - ch * a.kbs);   // k blocks past K are zero-filled: skip
                if (elect_one_sync()) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        const uint64_t da = umma_desc_k_sw128(base + a.w_bytes + kb * x_slab);
                        const uint64_t db = umma_desc_k_sw128(base + kb * w_slab);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {   // 16 elements = 32 bytes along K inside the swizzle row: +2
                            const int acc = (step + kb * 4 + k) & (a.nacc - 1);
                            umma_f16_ss(tmem_d0 + acc * a.acc_stride, da + 2 * k, db + 2 * k, idesc,
                                        (step + kb * 4 + k) >= a.nacc ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);   // the slot is free once these MMAs have read it
                }
                __syncwarp();
                step += nkb * 4;
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one_sync()) umma_commit(&tfull[buf]);
            __syncwarp();
        }
    } else if (warp >= 2 && warp <= 5) {
        // ============================ epilogue warps (2..5) ============================
        pdl_wait();
        const int quad = warp & 3;   // the TMEM lane quadrant this warp may read
        const int np = a.np;
        // push epilogue: [t_box][np] staging behind the ring
        const EpiRow<T> er = epi_row_init<T, M, EPI>(a, warp, lane, ring + a.stages * a.stage_bytes + M * 128);

        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int buf = a.nbuf == 2 ? (it & 1) : 0, use = a.nbuf == 2 ? (it >> 1) : it;
            const int g = tile / a.tiles_per_group, p0 = (tile % a.tiles_per_group) * np;
            mbar_wait(&tfull[buf], use & 1);
            tc_fence_after_sync();
            if (a.trace && tid == 64) a.trace[blockIdx.x * 8 + 2] = sk_timer_ns();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * (a.nacc * a.acc_stride);
            epilogue_tile<T, M, EPI>(a, er, tacc, g, p0, warp, lane);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (a.trace && tid == 0) a.trace[blockIdx.x * 8 + 3] = sk_timer_ns();
    if (warp == 1) {
        // Measured (scripts/trace_step.py, slot 5 - slot 3): this fence + release takes 1.1 us + 50 ns per accumulator
        // column the tile used (2.7 us at 32 columns, 11.5 us at 208) — the fence is where the time goes, and dropping
        // it only moves the wait into tcgen05.dealloc.  It is the SM hand-over cost after every tensor-memory kernel.
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, a.tmem_cols);
        if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + 5] = sk_timer_ns();   // tensor memory released
    }

    if ((EPI == kEpiPlain || EPI == kEpiPush) && a.norm_tail) {
        // ============ fused add + RMSNorm of the finished rows, by the CTAs that finish last ============
        // Every thread's output stores (or peer pushes) are ordered before this CTA's arrival.  Nothing below
        // can deadlock: a waiting CTA waits only for CTAs of this grid, which never wait for it.
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = static_cast<int>(atomicAdd(&a.tail_sync[0], 1u));
        __syncthreads();
        const int grid = static_cast<int>(gridDim.x), nn = min(a.norm.T, grid);
        const int j = s_ticket - (grid - nn);
        if (j < 0) return;
        uint32_t *state = reinterpret_cast<uint32_t *>(a.norm.region);
        if (tid == 0) {
            const unsigned long long t_wait = sk_timer_ns();   // bounded like the chained kernel's wait (sticky word [2])
            while (ld_acquire_gpu_u32(&a.tail_sync[0]) < static_cast<uint32_t>(grid)) {
                __nanosleep(64);
                if (sk_timer_ns() - t_wait > 2000000000ull) {
                    atomicExch(&a.tail_sync[2], 1u);
                    break;
                }
            }
            if (EPI == kEpiPush) {
                s_tail_state[0] = *reinterpret_cast<volatile uint32_t *>(&state[0]);
                s_tail_state[1] = *reinterpret_cast<volatile uint32_t *>(&state[2]);
            }
        }
        __syncthreads();
        __threadfence();
        uint32_t e = 0;
        bool skip = false, failed = false;
        if (EPI == kEpiPush) {
            e = s_tail_state[0] + 1;           // this all-reduce's epoch (the one the epilogue pushed with)
            skip = s_tail_state[1] != 0;       // a peer failed to arrive earlier: do not wait for it again
        }
        if (!skip) {
            for (int row = j; row < a.norm.T; row += nn) {
                if (EPI == kEpiPush) {
                    switch (a.norm.world) {
                        case 2: failed |= add_rmsnorm_row<T, 2>(a.norm, row, e, s_red); break;
                        case 4: failed |= add_rmsnorm_row<T, 4>(a.norm, row, e, s_red); break;
                        default: failed |= add_rmsnorm_row<T, 8>(a.norm, row, e, s_red); break;
                    }
                } else {
                    add_rmsnorm_row<T, 1>(a.norm, row, 0u, s_red);
                }
                __syncthreads();   // s_red is reused by the next row
            }
        }
        if (tid == 0) {
            if (EPI == kEpiPush && failed) atomicExch(&state[2], 1u);   // error word: a peer never arrived
            __threadfence();
            if (atomicAdd(&a.tail_sync[1], 1u) == static_cast<uint32_t>(nn - 1)) {
                // last norm CTA: counters back to zero for the next launch; the next all-reduce sees the next epoch
                a.tail_sync[0] = 0;
                a.tail_sync[1] = 0;
                if (EPI == kEpiPush) {
                    __threadfence();
                    state[0] = e;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Two projections chained in ONE kernel: y1 = epi1( silu_mul(x0 . W0^T) . W1^T )  — the MLP of a Llama block
// (gate_up + SwiGLU, then down_proj; models/src/llama.rs:362-366).
//
// Why: the step timeline (scripts/trace_step.py, per-SM hand-over) shows that an SM whose CTA used tensor memory
// takes 3.3 us (32 accumulator columns) to 12 us (208 columns: gate_up) to accept the next kernel's CTA, against
// 0.5 us after the decode-attention kernel, which uses none — and down_proj's CTAs need exactly the SMs gate_up's
// CTAs leave.  Here the CTA stays: phase 1 keeps the ring, the barriers and the tensor-memory allocation of phase 0,
// its weights start streaming into the ring as phase 0's last stages drain, and the only dependency is a grid-wide
// arrival counter ("every CTA has stored its slice of the activations") that gates phase 1's x loads alone.
// Both phases: one tile per CTA (single wave), common ring geometry (stage size = the larger of the two).
struct ChainArgs {
    SkinnyArgs p[2];
    uint32_t *sync;     // [2] zero between launches: phase-0 tiles stored, phase-1 x producers released
    int stages, stage_bytes, tmem_cols;
    int tmem_col1;      // first accumulator column of phase 1 (phase 0 starts at 0)
};

template <typename T, int M, int EPI1>
__global__ void __launch_bounds__(kSkThreads, 1)
gemm_chain2_kernel(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_x0,
                   const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_x1,
                   const ChainArgs c) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[kSkMaxStages], empty[kSkMaxStages], tfull[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SkinnyArgs &a0 = c.p[0], &a1 = c.p[1];
    const unsigned long long t_start = a0.trace ? sk_timer_ns() : 0ull;
    if (tid == 0) {
        for (int i = 0; i < c.stages; ++i) {
            mbar_init(&full[i], 2);
            mbar_init(&empty[i], 1);
        }
        mbar_init(&tfull[0], 1);
        mbar_init(&tfull[1], 1);
        fence_mbar_init();
        tma_prefetch_desc(&map_w0);
        tma_prefetch_desc(&map_x0);
        tma_prefetch_desc(&map_w1);
        tma_prefetch_desc(&map_x1);
    }
    __syncthreads();
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, c.tmem_cols);
        tc_fence_before_sync();
    }
    if (warp >= 1 && warp <= 5) {
        named_bar_sync(1, 5 * 32);
        tc_fence_after_sync();
    }
    const uint32_t tmem_base = (warp >= 1 && warp <= 5) ? tmem_base_smem : 0u;
    const int b = static_cast<int>(blockIdx.x);
    const bool has0 = b < a0.tiles, has1 = b < a1.tiles;   // one tile per CTA and phase

    if (warp == 0 || warp == 7) {
        // ============================ TMA producers (warp 0: W, warp 7: x) ============================
        const bool is_w = warp == 0;
        if (is_w) pdl_launch_dependents();
        if (!(is_w && a0.prefetch_w)) pdl_wait();   // weights are immutable: they stream ahead of the grid dependency
        if (a0.trace && !is_w && lane == 0) a0.trace[b * 8 + 6] = sk_timer_ns();
        int stage = 0;
        uint32_t phase = 0;
        for (int ph = 0; ph < 2; ++ph) {
            const SkinnyArgs &a = ph == 0 ? a0 : a1;
            if (!(ph == 0 ? has0 : has1)) continue;
            const CUtensorMap *mw = ph == 0 ? &map_w0 : &map_w1, *mx = ph == 0 ? &map_x0 : &map_x1;
            const int g = b / a.tiles_per_group, p0 = (b - g * a.tiles_per_group) * a.np;
            if (ph == 1 && !is_w) {
                // phase 1 reads what EVERY CTA stored in phase 0: wait for all of them, then order the generic-proxy
                // stores (made visible by the acquire) before this thread's async-proxy (TMA) reads
                if (lane == 0) {
                    // bounded: every CTA of this grid is resident (at most one per SM) and none waits for this one, so
                    // the wait ends within the kernel's own run time; the limit only keeps a fault elsewhere (a CTA that
                    // trapped) from turning into a hung device — the sticky word sync[2] records it
                    const unsigned long long t_wait = sk_timer_ns();
                    while (ld_acquire_gpu_u32(&c.sync[0]) < static_cast<uint32_t>(a0.tiles)) {
                        __nanosleep(32);
                        if (sk_timer_ns() - t_wait > 2000000000ull) {
                            atomicExch(&c.sync[2], 1u);
                            break;
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                    if (atomicAdd(&c.sync[1], 1u) == static_cast<uint32_t>(a1.tiles - 1)) {   // last reader: reset for the next launch
                        c.sync[0] = 0;
                        c.sync[1] = 0;
                    }
                }
                __syncwarp();
                if (a1.trace && lane == 0) a1.trace[b * 8 + 6] = sk_timer_ns();
            }
            for (int ch = 0; ch < a.chunks; ++ch) {
                mbar_wait(&empty[stage], phase ^ 1);
                unsigned char *dst = ring + stage * c.stage_bytes;
                if (a.trace && is_w && lane == 0 && ch == 0) {
                    unsigned smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    a.trace[b * 8 + 4] = smid;
                    a.trace[b * 8 + 0] = ph == 0 ? t_start : sk_timer_ns();
                }
                if (elect_one_sync()) {
                    if (is_w) {
                        mbar_arrive_expect_tx(&full[stage], a.w_bytes);
                        tma_load_5d(dst, mw, &full[stage], 0, p0, 0, ch * a.kbs, g, kL2EvictFirst);
                    } else {
                        mbar_arrive_expect_tx(&full[stage], a.x_bytes);
                        tma_load_3d(dst + a.w_bytes, mx, &full[stage], 0, 0, ch * a.kbs, kL2EvictLast);
                    }
                }
                __syncwarp();
                if (++stage == c.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        if (is_w && a0.prefetch_w) pdl_wait();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        pdl_wait();
        const uint32_t ring_u32 = smem_u32(ring);
        int stage = 0;
        uint32_t phase = 0;
        for (int ph = 0; ph < 2; ++ph) {
            const SkinnyArgs &a = ph == 0 ? a0 : a1;
            if (!(ph == 0 ? has0 : has1)) continue;
            const uint32_t idesc = umma_idesc_f16(M, a.n_tile, a.is_bf16 != 0);
            const int w_slab = a.n_tile * 128, x_slab = a.t_box * 128;
            const uint32_t tmem_d = tmem_base + (ph == 0 ? 0 : c.tmem_col1);
            int step = 0;
            for (int ch = 0; ch < a.chunks; ++ch) {
                mbar_wait(&full[stage], phase);
                tc_fence_after_sync();
                if (a.trace && ch == 0 && lane == 0) a.trace[b * 8 + 1] = sk_timer_ns();
                if (a.trace && ch == a.chunks - 1 && lane == 0) a.trace[b * 8 + 7] = sk_timer_ns();
                const uint32_t base = ring_u32 + stage * c.stage_bytes;
                const int nkb = min(a.kbs, a.kb_total - ch * a.kbs);
                if (elect_one_sync()) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        const uint64_t da = umma_desc_k_sw128(base + a.w_bytes + kb * x_slab);
                        const uint64_t db = umma_desc_k_sw128(base + kb * w_slab);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (step + kb * 4 + k) > 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                step += nkb * 4;
                if (++stage == c.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one_sync()) umma_commit(&tfull[ph]);
            __syncwarp();
        }
    } else if (warp >= 2 && warp <= 5) {
        // ============================ epilogue warps ============================
        pdl_wait();
        const int quad = warp & 3;
        unsigned char *stg = ring + c.stages * c.stage_bytes + M * 128;
        if (has0) {
            const EpiRow<T> er = epi_row_init<T, M, kEpiSilu>(a0, warp, lane, stg);
            const int g = b / a0.tiles_per_group, p0 = (b - g * a0.tiles_per_group) * a0.np;
            mbar_wait(&tfull[0], 0);
            tc_fence_after_sync();
            if (a0.trace && tid == 64) a0.trace[b * 8 + 2] = sk_timer_ns();
            epilogue_tile<T, M, kEpiSilu>(a0, er, tmem_base + (static_cast<uint32_t>(quad * 32) << 16), g, p0, warp, lane);
            // this CTA's slice of the activations is stored: publish it to every CTA's phase-1 x producer (the barrier
            // orders the four warps' stores before the one release)
            named_bar_sync(2, 128);
            if (tid == 64) {
                red_release_gpu_add(&c.sync[0], 1u);
                if (a0.trace) a0.trace[b * 8 + 3] = sk_timer_ns();
            }
        }
        if (has1) {
            const EpiRow<T> er = epi_row_init<T, M, EPI1>(a1, warp, lane, stg);
            const int g = b / a1.tiles_per_group, p0 = (b - g * a1.tiles_per_group) * a1.np;
            mbar_wait(&tfull[1], 0);
            tc_fence_after_sync();
            if (a1.trace && tid == 64) a1.trace[b * 8 + 2] = sk_timer_ns();
            epilogue_tile<T, M, EPI1>(a1, er, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c.tmem_col1, g, p0, warp, lane);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (a1.trace && tid == 0 && has1) a1.trace[b * 8 + 3] = sk_timer_ns();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, c.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct MapKey {
    const void *ptr;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    int rank, bf16;
    bool operator==(const MapKey &o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapEntry {
    MapKey key;
    CUtensorMap map;
};

// N-D tiled tensor map of 16-bit elements, 128-byte swizzle, innermost box = 64 elements.
bool make_tiled_map(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                    const uint32_t *box, bool is_bf16) {
    static thread_local std::vector<MapEntry> cache;
    MapKey key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = ptr;
    key.rank = rank;
    key.bf16 = is_bf16;
    for (int i = 0; i < rank; ++i) {
        key.dims[i] = dims[i];
        key.box[i] = box[i];
        if (i + 1 < rank) key.strides[i] = strides[i];
    }
    for (const MapEntry &e : cache)
        if (e.key == key) {
            *out = e.map;
            return true;
        }
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) s[i] = strides[i];
    }
    CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     static_cast<cuuint32_t>(rank), const_cast<void *>(ptr), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_status(kErrDriver, "skinny gemm: cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
        return false;
    }
    if (cache.size() >= 1024) cache.clear();
    cache.push_back(MapEntry{key, *out});
    return true;
}

void *g_sk_trace = nullptr;
int g_sk_force_np = 0;       // diagnostics: override the tile width (rows per half)
int g_sk_force_kbs = 0;      // diagnostics: override the k blocks per stage
int g_sk_stage_limit = 0;
int g_sk_debug_skip = 0;     // diagnostics (timing experiments only; results are garbage): 1 = no x loads, 2 = no W loads
int g_sk_w_ops = 1;          // weight ops per stage (1 | 2)
int g_sk_nacc = 0;           // diagnostics: cap on the independent accumulators per tile (0 = automatic)
int g_sk_budget_kib = 0;     // diagnostics: ring budget in KiB (0 = default)
int g_sk_trigger_late = 0;   // diagnostics: release the dependent grid N stages before the last weight load (0 = at kernel start)
int g_sk_single_buf = 1;     // one accumulator buffer for single-tile CTAs (diagnostics: skip_loads bit 3 turns it off)

// Rows per half per tile (multiple of 8, np * halves <= 256) minimising waves x (tile width + a fixed
// per-tile cost); ties go to the wider tile (fewer CTAs, fewer x re-reads).
// `step`: granularity of the width (8; 32 for the push epilogue, whose tiles should cover whole 32-column blocks of
// the peers' receive layout — a 56-column tile splits every row into 64-byte runs and was measured at 30 us per
// all-reduce on the 70B shapes against 15 us for 32-column tiles on the 8B shapes).
int choose_np(int G, int J, int halves, int sms, int step = 8) {
    int best = step;
    long best_cost = -1;
    const int np_max = std::max(step, std::min(256 / halves, (J + step - 1) / step * step));
    for (int np = step; np <= np_max; np += step) {
        const long tiles = static_cast<long>(G) * ((J + np - 1) / np);
        const long waves = (tiles + sms - 1) / sms;
        const long cost = waves * (static_cast<long>(np) * halves + 8);
        if (best_cost < 0 || cost <= best_cost) {
            best = np;
            best_cost = cost;
        }
    }
    return best;
}

template <typename T, int M>
cudaError_t launch_epi(int epi, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap &mw, const CUtensorMap &mx,
                       const SkinnyArgs &a) {
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev != di.ordinal) {
        const int lim = kSkSmemBudget + 128 * 128 + 1024;
        cudaError_t e = cudaSuccess;
        auto set = [&](auto *k) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            // every kernel of the step asks for the same (maximum) shared-memory carve-out: an SM cannot change
            // its L1 / shared split while CTAs of the previous kernel are resident
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        };
        set(gemm_skinny_kernel<T, M, kEpiPlain>);
        set(gemm_skinny_kernel<T, M, kEpiRope>);
        set(gemm_skinny_kernel<T, M, kEpiSilu>);
        set(gemm_skinny_kernel<T, M, kEpiPush>);
        if (e != cudaSuccess) return e;
        attr_dev = di.ordinal;
    }
    switch (epi) {
        case kEpiPlain: return launch_ex(gemm_skinny_kernel<T, M, kEpiPlain>, grid, dim3(kSkThreads), smem, stream, mw, mx, a);
        case kEpiRope: return launch_ex(gemm_skinny_kernel<T, M, kEpiRope>, grid, dim3(kSkThreads), smem, stream, mw, mx, a);
        case kEpiSilu: return launch_ex(gemm_skinny_kernel<T, M, kEpiSilu>, grid, dim3(kSkThreads), smem, stream, mw, mx, a);
        case kEpiPush: return launch_ex(gemm_skinny_kernel<T, M, kEpiPush>, grid, dim3(kSkThreads), smem, stream, mw, mx, a);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

void set_skinny_trace(void *device_buf) { g_sk_trace = device_buf; }
void set_skinny_tuning(int np, int kbs, int stages) {
    g_sk_force_np = np;
    g_sk_force_kbs = kbs;
    g_sk_stage_limit = stages;
}
void set_skinny_budget(int kib) {
    set_next_kernel_prefetch_off(kib < 0 ? 1 : 0);   // -1: default budget, next-kernel L2 prefetch off (A/B measurements)
    g_sk_budget_kib = kib > 0 ? kib : 0;
}
void set_skinny_debug(int skip, int w_ops) {
    g_sk_debug_skip = skip & 3;
    g_sk_single_buf = (skip & 8) ? 0 : 1;
    g_sk_trigger_late = (skip >> 8) & 0xff;   // bits 8-15
    g_sk_nacc = (skip >> 4) & 7;          // bits 4-6: accumulator cap (1, 2, 4; 0 = automatic)
    g_sk_w_ops = w_ops == 2 ? 2 : 1;
}

bool skinny_gemm_supported(int64_t T, int64_t K, int64_t ldx, int64_t ldw) {
    return T >= 1 && T <= 128 && K >= 64 && K % 64 == 0 && ldx % 8 == 0 && ldw % 8 == 0 &&
           current_device_info().cc_major == 10;
}

// Launch one skinny GEMM.  `a` arrives with the epilogue fields filled in; this fills the tiling.
//   w: G groups x halves x J rows (row stride ldw; halves `half_stride_rows` apart, groups
//   `group_stride_rows` apart), x [T, K] row stride ldx.
bool launch_skinny_gemm(SkinnyArgs a, const void *x, const void *w, int T, int64_t K, int64_t ldx, int64_t ldw, int G,
                        int J, int halves, int64_t half_stride_rows, int64_t group_stride_rows, bool is_bf16,
                        cudaStream_t stream) {
    const DeviceInfo &di = current_device_info();
    a.T = T;
    a.m = T <= 64 ? 64 : 128;
    a.t_box = (T + 7) / 8 * 8;
    a.kb_total = static_cast<int>(K / 64);
    a.J = J;
    a.G = G;
    a.halves = halves;
    a.np = g_sk_force_np > 0 ? g_sk_force_np : choose_np(G, J, halves, di.sm_count, (a.epi == kEpiPush && J % 32 == 0) ? 32 : 8);
    a.n_tile = a.np * halves;
    a.tiles_per_group = (J + a.np - 1) / a.np;
    a.tiles = G * a.tiles_per_group;
    const int per_kb = (a.n_tile + a.t_box) * 128;
    const int budget = g_sk_budget_kib > 0 ? g_sk_budget_kib * 1024 : kSkDefaultBudget;
    const int stage_target = std::min(kSkStageTarget, budget / 2);
    int kbs = 16;
    while (kbs > 1 && (kbs * per_kb > stage_target || kbs > a.kb_total)) kbs >>= 1;
    if (g_sk_force_kbs > 0) kbs = g_sk_force_kbs;
    a.kbs = kbs;
    a.w_ops = (kbs >= 2 && g_sk_w_ops == 2) ? 2 : 1;   // the weight box of a stage is issued as one or two ops
    a.debug_skip = g_sk_debug_skip;
    a.trigger_late = g_sk_trigger_late;
    a.wkb = kbs / a.w_ops;
    a.chunks = (a.kb_total + kbs - 1) / kbs;
    a.w_bytes = kbs * a.n_tile * 128;
    a.x_bytes = kbs * a.t_box * 128;
    a.stage_bytes = a.w_bytes + a.x_bytes;
    // push epilogue: the finished tile is staged in shared memory behind the ring ([t_box][np] elements)
    const int stg_bytes = a.epi == kEpiPush ? a.t_box * a.n_tile * 2 : 0;
    a.stages = std::min(kSkMaxStages, (budget - stg_bytes) / a.stage_bytes);
    if (g_sk_stage_limit > 0) a.stages = std::min(a.stages, g_sk_stage_limit);
    if (a.stages < 2) {
        set_status(kErrUnsupported, "skinny gemm: stage of %d bytes does not fit", a.stage_bytes);
        return false;
    }
    a.acc_stride = (a.n_tile + 31) / 32 * 32;
    // independent accumulation chains (diagnostics only: measured, they do not help — the cost was the
    // issuing thread's instruction stream, not the tensor pipe)
    a.nacc = 1;
    if (g_sk_nacc > 1) a.nacc = std::min(g_sk_nacc, a.acc_stride <= 64 ? 4 : (a.acc_stride <= 128 ? 2 : 1));
    // two accumulator buffers (the epilogue of a tile overlaps the MMAs of the CTA's next tile) when they fit the
    // CTA's share of tensor memory: all 512 columns for a CTA that owns its SM, 256 for a half-footprint CTA
    const bool half = budget <= kSkHalfBudget;
    a.nbuf = (2 * a.nacc * a.acc_stride <= (half ? 256 : 512)) ? 2 : 1;
    // a CTA that owns a single tile never overlaps an epilogue with a next tile: one buffer, half the columns
    if (g_sk_single_buf && a.tiles <= di.sm_count) a.nbuf = 1;
    int cols = 32;
    while (cols < a.nbuf * a.nacc * a.acc_stride) cols <<= 1;
    a.tmem_cols = cols;
    a.is_bf16 = is_bf16 ? 1 : 0;
    a.prefetch_w = ((launch_flags() & (kLaunchPdl | kLaunchConstWeights)) == (kLaunchPdl | kLaunchConstWeights)) ? 1 : 0;
    a.trace = static_cast<unsigned long long *>(g_sk_trace);
    if (!a.trace) a.trace = next_step_trace_slot(a.epi == kEpiRope ? "gemm qkv+rope" : a.epi == kEpiSilu ? "gemm gate_up+silu" :
                                                 a.epi == kEpiPush ? (a.norm_tail ? "gemm push+norm" : "gemm push") :
                                                 (a.norm_tail ? "gemm+norm" : "gemm"));
    {
        const NextPrefetch pf = take_next_kernel_prefetch();
        const bool ok = pf.ptr && pf.bytes >= 16 && (reinterpret_cast<uintptr_t>(pf.ptr) & 15) == 0;
        a.pf_ptr = ok ? static_cast<const unsigned char *>(pf.ptr) : nullptr;
        a.pf_bytes = ok ? pf.bytes : 0;
        // narrow tiles are bound by the tensor-pipe issue rate, not HBM: the prefetch rides along from the start
        a.pf_spread = a.n_tile <= 64 ? 1 : 0;
    }

    if (a.norm_tail) {
        if ((a.epi != kEpiPlain && a.epi != kEpiPush) || a.norm.hidden > 8 * kSkThreads * kNormMaxVec || a.norm.T != T) {
            set_status(kErrInvalidArg, "skinny gemm: the fused add+RMSNorm tail needs a plain / push epilogue, hidden <= %d and "
                                       "one norm row per output row", 8 * kSkThreads * kNormMaxVec);
            return false;
        }
        a.tail_sync = static_cast<uint32_t *>(get_workspace(stream, 256, kWorkspaceSkinnyTail));   // zero-initialised, self-resetting
        if (!a.tail_sync) return false;
    }

    CUtensorMap mw, mx;
    {
        const uint64_t dims[5] = {64, static_cast<uint64_t>(J), static_cast<uint64_t>(halves),
                                  static_cast<uint64_t>(a.kb_total), static_cast<uint64_t>(G)};
        const uint64_t strides[4] = {static_cast<uint64_t>(ldw) * 2, static_cast<uint64_t>(half_stride_rows) * ldw * 2, 128,
                                     static_cast<uint64_t>(group_stride_rows) * ldw * 2};
        // sizes of 1 keep their (unused) stride legal: any multiple of 16
        uint64_t st[4] = {strides[0], halves > 1 ? strides[1] : strides[0], 128, G > 1 ? strides[3] : strides[0]};
        const uint32_t box[5] = {64, static_cast<uint32_t>(a.np), static_cast<uint32_t>(halves), static_cast<uint32_t>(a.wkb), 1};
        if (!make_tiled_map(&mw, w, 5, dims, st, box, is_bf16)) return false;
    }
    {
        const uint64_t dims[3] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(a.kb_total)};
        const uint64_t strides[2] = {static_cast<uint64_t>(ldx) * 2, 128};
        const uint32_t box[3] = {64, static_cast<uint32_t>(a.t_box), static_cast<uint32_t>(kbs)};
        if (!make_tiled_map(&mx, x, 3, dims, strides, box, is_bf16)) return false;
    }
    // ring + the rows an M-row A descriptor reads past the last x slab + alignment slack
    const size_t smem = static_cast<size_t>(a.stages) * a.stage_bytes + a.m * 128 + 1024 + stg_bytes;
    const int ctas_per_sm = (half && smem <= 113 * 1024 && a.tmem_cols <= 256) ? 2 : 1;
    const int grid = std::min(a.tiles, di.sm_count * ctas_per_sm);
    cudaError_t e;
    if (is_bf16) {
        e = a.m == 64 ? launch_epi<__nv_bfloat16, 64>(a.epi, dim3(grid), smem, stream, mw, mx, a)
                      : launch_epi<__nv_bfloat16, 128>(a.epi, dim3(grid), smem, stream, mw, mx, a);
    } else {
        e = a.m == 64 ? launch_epi<__half, 64>(a.epi, dim3(grid), smem, stream, mw, mx, a)
                      : launch_epi<__half, 128>(a.epi, dim3(grid), smem, stream, mw, mx, a);
    }
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "gemm_skinny_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("gemm_skinny_kernel");
}

// ---- chained MLP (gemm_chain2_kernel) ----
namespace {

// Tiling of one phase of a chain: the single-kernel rules of launch_skinny_gemm (tile width from choose_np, as many k
// blocks per stage as fit `stage_cap`), without the diagnostics overrides.
void plan_chain_phase(SkinnyArgs &a, int T, int64_t K, int G, int J, int halves, int stage_cap, int sms, bool is_bf16) {
    a.T = T;
    a.m = T <= 64 ? 64 : 128;
    a.t_box = (T + 7) / 8 * 8;
    a.kb_total = static_cast<int>(K / 64);
    a.J = J;
    a.G = G;
    a.halves = halves;
    a.np = choose_np(G, J, halves, sms, (a.epi == kEpiPush && J % 32 == 0) ? 32 : 8);
    a.n_tile = a.np * halves;
    a.tiles_per_group = (J + a.np - 1) / a.np;
    a.tiles = G * a.tiles_per_group;
    const int per_kb = (a.n_tile + a.t_box) * 128;
    int kbs = 16;
    while (kbs > 1 && (kbs * per_kb > stage_cap || kbs > a.kb_total)) kbs >>= 1;
    a.kbs = kbs;
    a.w_ops = 1;
    a.wkb = kbs;
    a.chunks = (a.kb_total + kbs - 1) / kbs;
    a.w_bytes = kbs * a.n_tile * 128;
    a.x_bytes = kbs * a.t_box * 128;
    a.stage_bytes = a.w_bytes + a.x_bytes;
    a.acc_stride = (a.n_tile + 31) / 32 * 32;
    a.nacc = 1;
    a.nbuf = 1;
    a.is_bf16 = is_bf16 ? 1 : 0;
    a.prefetch_w = ((launch_flags() & (kLaunchPdl | kLaunchConstWeights)) == (kLaunchPdl | kLaunchConstWeights)) ? 1 : 0;
}

template <typename T, int M>
cudaError_t launch_chain_epi(int epi1, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap &mw0, const CUtensorMap &mx0,
                             const CUtensorMap &mw1, const CUtensorMap &mx1, const ChainArgs &c) {
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev != di.ordinal) {
        const int lim = kSkSmemBudget + 128 * 128 + 1024;   // ring + staging <= budget, + operand slack + alignment
        cudaError_t e = cudaSuccess;
        auto set = [&](auto *k) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        };
        set(gemm_chain2_kernel<T, M, kEpiPlain>);
        set(gemm_chain2_kernel<T, M, kEpiPush>);
        if (e != cudaSuccess) return e;
        attr_dev = di.ordinal;
    }
    if (epi1 == kEpiPush)
        return launch_ex(gemm_chain2_kernel<T, M, kEpiPush>, grid, dim3(kSkThreads), smem, stream, mw0, mx0, mw1, mx1, c);
    return launch_ex(gemm_chain2_kernel<T, M, kEpiPlain>, grid, dim3(kSkThreads), smem, stream, mw0, mx0, mw1, mx1, c);
}

bool make_phase_maps(CUtensorMap *mw, CUtensorMap *mx, const SkinnyArgs &a, const void *x, const void *w, int64_t ldx, int64_t ldw,
                     int64_t half_stride_rows, bool is_bf16) {
    const uint64_t wd[5] = {64, static_cast<uint64_t>(a.J), static_cast<uint64_t>(a.halves), static_cast<uint64_t>(a.kb_total), 1};
    const uint64_t row_b = static_cast<uint64_t>(ldw) * 2;
    const uint64_t ws[4] = {row_b, a.halves > 1 ? static_cast<uint64_t>(half_stride_rows) * row_b : row_b, 128, row_b};
    const uint32_t wb[5] = {64, static_cast<uint32_t>(a.np), static_cast<uint32_t>(a.halves), static_cast<uint32_t>(a.kbs), 1};
    if (!make_tiled_map(mw, w, 5, wd, ws, wb, is_bf16)) return false;
    const uint64_t xd[3] = {64, static_cast<uint64_t>(a.T), static_cast<uint64_t>(a.kb_total)};
    const uint64_t xs[2] = {static_cast<uint64_t>(ldx) * 2, 128};
    const uint32_t xb[3] = {64, static_cast<uint32_t>(a.t_box), static_cast<uint32_t>(a.kbs)};
    return make_tiled_map(mx, x, 3, xd, xs, xb, is_bf16);
}

}  // namespace

int g_sk_chain_mlp = 1;   // atoma_debug_step_fusion bit 1: 0 = gate_up and down_proj as two kernels
void set_skinny_chain(int on) { g_sk_chain_mlp = on ? 1 : 0; }

// act[T, inter] = silu(x Wg^T) * (x Wu^T) (stored, as the unchained pair does), then `a1.epi` (plain into a1.y, or
// push) of act . Wdown^T, in one launch.  Returns 0 when the shapes do not qualify (the caller launches the two
// kernels), 1 when launched, -1 on error.
int try_launch_mlp_chain(const void *x, const void *w_gate_up, void *act, const void *w_down, SkinnyArgs a1, int T, int64_t inter,
                         int64_t hidden, int64_t ldx, int64_t ldw_gu, int64_t ld_act, int64_t ldw_down, bool is_bf16,
                         cudaStream_t stream) {
    const DeviceInfo &di = current_device_info();
    if (!g_sk_chain_mlp || g_sk_trace || g_sk_force_np || g_sk_force_kbs || g_sk_stage_limit || g_sk_budget_kib || g_sk_debug_skip ||
        !(launch_flags() & kLaunchPdl))
        return 0;
    if (!skinny_gemm_supported(T, hidden, ldx, ldw_gu) || !skinny_gemm_supported(T, inter, ld_act, ldw_down) || inter % 8) return 0;
    constexpr int kStageCap = 64 * 1024;
    ChainArgs c{};
    SkinnyArgs &a0 = c.p[0];
    a0.epi = kEpiSilu;
    a0.y = act;
    a0.ldy = ld_act;
    plan_chain_phase(a0, T, hidden, 1, static_cast<int>(inter), 2, kStageCap, di.sm_count, is_bf16);
    c.p[1] = a1;
    SkinnyArgs &p1 = c.p[1];
    plan_chain_phase(p1, T, inter, 1, static_cast<int>(hidden), 1, kStageCap, di.sm_count, is_bf16);
    if (a0.tiles > di.sm_count || p1.tiles > di.sm_count) return 0;   // single wave only
    c.stage_bytes = std::max(a0.stage_bytes, p1.stage_bytes);
    const int stg_bytes = p1.epi == kEpiPush ? p1.t_box * p1.n_tile * 2 : 0;
    c.stages = std::min(kSkMaxStages, (kSkSmemBudget - stg_bytes) / c.stage_bytes);
    if (c.stages < 2) return 0;
    a0.stages = p1.stages = c.stages;
    a0.stage_bytes = p1.stage_bytes = c.stage_bytes;
    c.tmem_col1 = a0.acc_stride;
    int cols = 32;
    while (cols < a0.acc_stride + p1.acc_stride) cols <<= 1;
    if (cols > 512) return 0;
    c.tmem_cols = cols;
    a0.tmem_cols = p1.tmem_cols = cols;
    c.sync = static_cast<uint32_t *>(get_workspace(stream, 256, kWorkspaceSkinnyTail));
    if (!c.sync) return -1;
    c.sync += 8;   // words 8, 9 of the tail workspace (0, 1 belong to the fused norm tail)
    a0.trace = next_step_trace_slot("chain gate_up+silu");
    p1.trace = next_step_trace_slot(p1.epi == kEpiPush ? "chain down push" : "chain down");
    take_next_kernel_prefetch();   // hints are not used by the chain
    CUtensorMap mw0, mx0, mw1, mx1;
    if (!make_phase_maps(&mw0, &mx0, a0, x, w_gate_up, ldx, ldw_gu, inter, is_bf16)) return -1;
    if (!make_phase_maps(&mw1, &mx1, p1, act, w_down, ld_act, ldw_down, 0, is_bf16)) return -1;
    const int grid = std::max(a0.tiles, p1.tiles);
    const size_t smem = static_cast<size_t>(c.stages) * c.stage_bytes + a0.m * 128 + 1024 + stg_bytes;
    cudaError_t e;
    if (is_bf16) {
        e = a0.m == 64 ? launch_chain_epi<__nv_bfloat16, 64>(p1.epi, dim3(grid), smem, stream, mw0, mx0, mw1, mx1, c)
                       : launch_chain_epi<__nv_bfloat16, 128>(p1.epi, dim3(grid), smem, stream, mw0, mx0, mw1, mx1, c);
    } else {
        e = a0.m == 64 ? launch_chain_epi<__half, 64>(p1.epi, dim3(grid), smem, stream, mw0, mx0, mw1, mx1, c)
                       : launch_chain_epi<__half, 128>(p1.epi, dim3(grid), smem, stream, mw0, mx0, mw1, mx1, c);
    }
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "gemm_chain2_kernel: %s", cudaGetErrorString(e));
        return -1;
    }
    return check_launch("gemm_chain2_kernel") ? 1 : -1;
}

}  // namespace atoma
