// Prefill / dense attention forward on the 5th-generation tensor cores (tcgen05 + TMEM) for
// sm_100a: contiguous (non-paged) K/V, head dim 64 or 128, causal or not, varlen or batched, GQA.
// This is the path `flash_attn_varlen` takes at prefill (reference csrc/src/lib.rs:606-1103 ->
// run_mha -> flash_fwd_kernel, csrc/kernels/flash_fwd_kernel.h:56-500) and `flash_attn` (:31-343).
//
// Same contract as the reference (fp32 scores, exp2-domain online softmax, P rounded to the io type
// before P.V, fp32 accumulation, one final 1/rowsum, bottom-right aligned causal mask mask.h:171-172,
// rows without a visible key -> 0); the execution model is new:
//  * one CTA = one 128-row query tile of one head; K/V tiles of 128 keys stream through a 2-stage
//    TMA ring (128-byte swizzle boxes straight from the [tokens, heads, d] layout);
//  * warp 1 issues S = Q K^T (tcgen05.mma, A and B from shared memory, 128x128 fp32 accumulator in
//    TENSOR MEMORY, double buffered) and O += P V (A = P read from tensor memory, B = V from shared
//    memory as an MN-major operand, 128 x D fp32 accumulator in tensor memory);
//  * warps 2-5 own one query row per thread (TMEM lane = row): tcgen05.ld the 128 scores of the row,
//    softmax them with a LAZY running maximum (the reference maximum is only moved, and O rescaled in
//    tensor memory, when a row's maximum grows by more than 2^8 — otherwise exp2 stays in range and
//    no correction pass is needed), write P back over S as packed 16-bit pairs (tcgen05.st);
//  * QK^T of tile j+1 is issued before P.V of tile j, so the tensor pipe works while tile j is being
//    exponentiated; the epilogue reads O from tensor memory, scales by 1/rowsum and stores 16-byte
//    vectors;
//  * tiles that stick out of their sequence (ragged tails) are loaded with cp.async + zero fill in
//    the same swizzled layout instead of TMA, so nothing outside a sequence is ever multiplied in.
//
// Bound: tensor pipe.  Algorithmic flops: 4 * sum_b(visible (q,k) pairs) * D per head.
#include "common.cuh"
#include "mha_params.h"

#include <cstring>
#include <type_traits>
#include <vector>

namespace atoma {
namespace {

constexpr int kPfThreads = 384;  // warpgroup 0: loads (warp 0) + MMA issue (warp 1); warpgroups 1, 2: softmax
constexpr int kPfM = 128, kPfN = 128;
constexpr int kQTiles = 2;       // two 128-row query tiles per CTA, one softmax warpgroup each
constexpr int kKvStages = 2;
constexpr int kTmemCols = 512;   // S_A 0..127 | S_B 128..255 | O_A 256.. | O_B 384..
constexpr int kPfIssueRegs = 56, kPfSoftmaxRegs = 224;  // setmaxnreg: 168 * 384 = 56 * 128 + 224 * 256
constexpr float kRescaleThreshold = 8.f;  // log2 units: P <= 2^8 between corrections

template <int D>
struct PfCfg {
    static constexpr int kHalves = D / 64;
    static constexpr int kHalfBytes = 128 * 128;            // 128 rows of 128 bytes
    static constexpr int kTileBytes = kHalves * kHalfBytes;  // one Q / K / V tile
    static constexpr int kSmem = kTileBytes * (kQTiles + 2 * kKvStages) + 1024;
};

struct PfSeq {
    int len_q, len_k;
    int q_row0, k_row0;      // TMA row coordinate of the sequence's first token
    int q_batch, k_batch;    // TMA batch coordinate
    int64_t q_off, o_off, k_off, v_off;  // element offsets of the sequence's first token
    int q_start;             // cu_q[b] or -1
};

__device__ __forceinline__ PfSeq pf_seq(const MhaParams &p, int bidb) {
    PfSeq s;
    // reference block_info.h:14-36
    s.q_start = p.cu_q ? p.cu_q[bidb] : -1;
    s.len_q = p.cu_q ? p.cu_q[bidb + 1] - s.q_start : p.seqlen_q;
    const int k_start = (p.cu_k && p.k_cumulative) ? p.cu_k[bidb] : -1;
    const int len_k_cache = p.cu_k ? (p.k_cumulative ? p.cu_k[bidb + 1] - k_start : p.cu_k[bidb]) : p.seqlen_k;
    s.len_k = p.seqused_k ? p.seqused_k[bidb] : len_k_cache;
    s.q_off = s.q_start < 0 ? bidb * p.q_bs : static_cast<int64_t>(s.q_start) * p.q_rs;
    s.o_off = s.q_start < 0 ? bidb * p.o_bs : static_cast<int64_t>(s.q_start) * p.o_rs;
    s.k_off = k_start < 0 ? bidb * p.k_bs : static_cast<int64_t>(k_start) * p.k_rs;
    s.v_off = k_start < 0 ? bidb * p.v_bs : static_cast<int64_t>(k_start) * p.v_rs;
    s.q_row0 = s.q_start < 0 ? 0 : s.q_start;
    s.q_batch = s.q_start < 0 ? bidb : 0;
    s.k_row0 = k_start < 0 ? 0 : k_start;
    s.k_batch = k_start < 0 ? bidb : 0;
    return s;
}

template <typename T, int D>
__global__ void __launch_bounds__(kPfThreads, 1)
prefill_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const MhaParams p) {
    using C = PfCfg<D>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *sq = tiles;                            // [q tile]
    unsigned char *sk = sq + kQTiles * C::kTileBytes;     // [stage]
    unsigned char *sv = sk + kKvStages * C::kTileBytes;   // [stage]
    __shared__ __align__(8) uint64_t q_full[kQTiles], k_full[kKvStages], k_empty[kKvStages], v_full[kKvStages],
        v_empty[kKvStages], s_full[kQTiles], p_full[kQTiles], pv_done[kQTiles];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_block = gridDim.x - 1 - blockIdx.x;  // heaviest (causal) tiles first
    const int head = blockIdx.y, bidb = blockIdx.z;

    if (tid == 0) {
        for (int i = 0; i < kKvStages; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < kQTiles; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);  // one arrival per softmax warp
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_launch_dependents();
    pdl_wait();

    const PfSeq si = pf_seq(p, bidb);
    const int len_q = si.len_q, len_k = si.len_k;
    const int shift = len_k - len_q;  // mask.h:171-172: key j visible to query i iff j <= i + shift
    // Query tile t of this CTA covers rows m0[t] .. m0[t] + 127 and needs the first nt[t] key tiles.
    int m0[kQTiles], nt[kQTiles];
#pragma unroll
    for (int t = 0; t < kQTiles; ++t) {
        m0[t] = (m_block * kQTiles + t) * kPfM;
        nt[t] = 0;
        if (m0[t] < len_q) {
            nt[t] = (len_k + kPfN - 1) / kPfN;
            if (p.is_causal) {
                const int lim = min(m0[t] + kPfM, len_q) + shift;  // one past the last visible key of the tile
                nt[t] = min(nt[t], lim <= 0 ? 0 : (lim + kPfN - 1) / kPfN);
            }
        }
    }
    const int n_max = max(nt[0], nt[1]);
    const int kvh = head / (p.h / p.h_k);  // flash_fwd_kernel.h:593

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPfIssueRegs));
        if (warp == 0 && n_max > 0) {
            // ================================ loads ================================
            // A tile that lies completely inside its sequence is one TMA transaction (kHalves boxes);
            // a ragged one is filled with cp.async + zero fill in the same 128-byte-swizzled layout.
            auto load_tile = [&](unsigned char *dst, const CUtensorMap *map, uint64_t *bar, const T *gbase,
                                 int64_t rs, int row0, int len, int tma_row, int tma_head, int tma_batch,
                                 uint64_t policy) {
                if (row0 + 128 <= len) {
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(bar, C::kTileBytes);
#pragma unroll
                        for (int hf = 0; hf < C::kHalves; ++hf)
                            tma_load_4d(dst + hf * C::kHalfBytes, map, bar, hf * 64, tma_head, tma_row, tma_batch, policy);
                    }
                    __syncwarp();
                } else {
                    for (int idx = lane; idx < 128 * (D / 8); idx += 32) {
                        const int r = idx / (D / 8), c = idx % (D / 8);
                        const bool ok = row0 + r < len;
                        const T *src = ok ? gbase + static_cast<int64_t>(row0 + r) * rs + c * 8 : gbase;
                        cp_async_16(smem_u32(dst + (c >> 3) * C::kHalfBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4)), src,
                                    ok ? 16 : 0);
                    }
                    cp_async_commit();
                    cp_async_wait_all();
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar);
                }
            };
            const T *qg = static_cast<const T *>(p.q) + si.q_off + static_cast<int64_t>(head) * p.q_hs;
            const T *kg = static_cast<const T *>(p.k) + si.k_off + static_cast<int64_t>(kvh) * p.k_hs;
            const T *vg = static_cast<const T *>(p.v) + si.v_off + static_cast<int64_t>(kvh) * p.v_hs;
#pragma unroll
            for (int t = 0; t < kQTiles; ++t)
                if (nt[t] > 0)
                    load_tile(sq + t * C::kTileBytes, &map_q, &q_full[t], qg, p.q_rs, m0[t], len_q, si.q_row0 + m0[t],
                              head, si.q_batch, kL2EvictFirst);
            for (int j = 0; j < n_max; ++j) {
                const int s = j % kKvStages;
                const uint32_t par = ((j / kKvStages) & 1) ^ 1;
                mbar_wait(&k_empty[s], par);
                load_tile(sk + s * C::kTileBytes, &map_k, &k_full[s], kg, p.k_rs, j * kPfN, len_k,
                          si.k_row0 + j * kPfN, kvh, si.k_batch, kL2EvictLast);
                mbar_wait(&v_empty[s], par);
                load_tile(sv + s * C::kTileBytes, &map_v, &v_full[s], vg, p.v_rs, j * kPfN, len_k,
                          si.k_row0 + j * kPfN, kvh, si.k_batch, kL2EvictLast);
            }
        } else if (warp == 1 && n_max > 0) {
            // ================================ MMA issue ================================
            // The whole warp runs this role converged on warp-uniform values; only the tcgen05
            // instructions sit under elect.sync.  (Inside an `if (lane == 0)` region the compiler keeps
            // the operands in vector registers and wraps every UTCHMMA in an R2UR + ELECT waterfall loop,
            // ~70 cycles per MMA against the 64 cycles a 128x128x16 MMA occupies the tensor pipe — the
            // issue thread, not the pipe, was the limit of round 1's kernel: tensor pipe 49 % active.)
            // Order per key tile j:  P.V(A,j)  QK(A,j+1)  P.V(B,j)  QK(B,j+1)
            // so while one query tile is being exponentiated the tensor pipe works for the other.
            constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
            const uint32_t idesc_qk = umma_idesc_f16(kPfM, kPfN, kBf16, false);
            const uint32_t idesc_pv = umma_idesc_f16(kPfM, D, kBf16, true);
            auto issue_qk = [&](int t, int j) {
                const int s = j % kKvStages;
                const uint32_t q_addr = smem_u32(sq + t * C::kTileBytes);
                const uint32_t k_addr = smem_u32(sk + s * C::kTileBytes);
                const uint32_t tmem_s = tmem_base + t * 128;
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = (kk >> 2) * C::kHalfBytes + (kk & 3) * 32;
                        umma_f16_ss(tmem_s, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off), idesc_qk,
                                    kk != 0 ? 1u : 0u);
                    }
                    umma_commit(&s_full[t]);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int t, int j) {
                const int s = j % kKvStages;
                const uint32_t v_addr = smem_u32(sv + s * C::kTileBytes);
                const uint32_t tmem_p = tmem_base + t * 128;
                const uint32_t tmem_o = tmem_base + 256 + t * 128;
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < kPfN / 16; ++kk) {
                        // 16 keys = 16 rows of 128 bytes of V; 16 P values = 8 packed 32-bit columns
                        umma_f16_ts(tmem_o, tmem_p + kk * 8, umma_desc_mn_sw128(v_addr + kk * 16 * 128, C::kHalfBytes),
                                    idesc_pv, (j | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                }
                __syncwarp();
            };
            // first key tile: both query tiles
            mbar_wait(&k_full[0], 0);
#pragma unroll
            for (int t = 0; t < kQTiles; ++t) {
                if (nt[t] > 0) {
                    mbar_wait(&q_full[t], 0);
                    tc_fence_after_sync();
                    issue_qk(t, 0);
                }
            }
            if (elect_one_sync()) umma_commit(&k_empty[0]);
            __syncwarp();
            for (int j = 0; j < n_max; ++j) {
                const int s = j % kKvStages, sn = (j + 1) % kKvStages;
                mbar_wait(&v_full[s], (j / kKvStages) & 1);
                bool k_next_ready = false;
#pragma unroll
                for (int t = 0; t < kQTiles; ++t) {
                    if (j < nt[t]) {
                        mbar_wait(&p_full[t], j & 1);
                        tc_fence_after_sync();
                        issue_pv(t, j);
                    }
                    if (j + 1 < nt[t]) {
                        if (!k_next_ready) {
                            mbar_wait(&k_full[sn], ((j + 1) / kKvStages) & 1);
                            k_next_ready = true;
                        }
                        tc_fence_after_sync();
                        issue_qk(t, j + 1);
                    }
                }
                if (elect_one_sync()) {
                    umma_commit(&v_empty[s]);                       // both P.V of tile j are issued
                    if (k_next_ready) umma_commit(&k_empty[sn]);    // ... and every QK^T of tile j + 1
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ softmax / epilogue warpgroups ================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPfSoftmaxRegs));
        const int t = (warp >> 2) - 1;          // query tile of this warpgroup
        const int quad = warp & 3;
        const int row = quad * 32 + lane;       // query row of this thread within the tile
        const int m0t = t ? m0[1] : m0[0];
        const int grow = m0t + row;             // ... within the sequence
        const int n_tiles = t ? nt[1] : nt[0];
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tmem_s = tmem_base + lane_addr + t * 128;
        const uint32_t tmem_o = tmem_base + lane_addr + 256 + t * 128;
        T *og = static_cast<T *>(p.o) + si.o_off + static_cast<int64_t>(head) * p.o_hs +
                static_cast<int64_t>(grow) * p.o_rs;
        int64_t lse_idx = 0;
        if (p.lse && grow < len_q) {
            if (p.unpadded_lse && p.cu_q) lse_idx = static_cast<int64_t>(head) * p.cu_q[p.b] + si.q_start + grow;
            else lse_idx = (static_cast<int64_t>(bidb) * p.h + head) * p.seqlen_q + grow;
        }
        if (n_tiles > 0) {
            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = 0; j < n_tiles; ++j) {
                mbar_wait(&s_full[t], j & 1);
                tc_fence_after_sync();
                uint32_t sv_[128];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_s + c * 32, sv_ + c * 32);
                tmem_ld_wait();
                // ---- masks: only tiles that touch the sequence end or the causal diagonal ----
                const int col0 = j * kPfN;
                const bool tail = col0 + kPfN > len_k;
                const bool diag = p.is_causal && (col0 + kPfN - 1 > m0t + shift);
                if (tail || diag) {
                    const int lim = p.is_causal ? min(len_k, grow + shift + 1) : len_k;  // first masked column
#pragma unroll
                    for (int c = 0; c < 128; ++c)
                        if (col0 + c >= lim) sv_[c] = 0xff800000u;  // -inf
                }
                // eight independent chains: a single running max / sum is 128 dependent operations
                float mx8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sv_[i]);
#pragma unroll
                for (int c = 8; c < 128; ++c) mx8[c & 7] = fmaxf(mx8[c & 7], __uint_as_float(sv_[c]));
                const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                       fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
                const float m_tile = mx * p.scale_log2;  // scale_log2 > 0 (checked on the host)
                // ---- lazy running maximum ----
                const bool need = m_tile > m_ref + kRescaleThreshold;  // also the first finite tile (m_ref = -inf)
                if (__any_sync(0xffffffffu, need)) {
                    const float alpha = need ? (m_ref == -INFINITY ? 0.f : fast_exp2(m_ref - m_tile)) : 1.f;
                    if (need) m_ref = m_tile;
                    l_run *= alpha;
                    if (j > 0) {
                        // O of this row lives in tensor memory: wait until P.V of the previous tile has
                        // been accumulated, rescale, write back.  Rare after the first tiles.
                        mbar_wait(&pv_done[t], (j - 1) & 1);
                        tc_fence_after_sync();
#pragma unroll
                        for (int c = 0; c < D / 32; ++c) {
                            uint32_t ov[32];
                            tmem_ld_32x32b_x32(tmem_o + c * 32, ov);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                            tmem_st_32x32b_x32(tmem_o + c * 32, ov);
                        }
                        tmem_st_wait();
                    }
                }
                const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
                // ---- P = exp2(S * scale - m), row sum, packed 16-bit pairs back over S ----
                float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(sv_[c * 32 + 2 * i]), p.scale_log2, -m_use));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(sv_[c * 32 + 2 * i + 1]), p.scale_log2, -m_use));
                        ls[i & 3] += p0 + p1;
                        pk[i] = pack2<T>(p0, p1);
                    }
                    tmem_st_32x32b_x16(tmem_s + c * 16, pk);
                }
                l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
            }
            // ---- epilogue: O / rowsum -> global ----
            mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
            tc_fence_after_sync();
            const bool empty = (l_run == 0.f || l_run != l_run);
            const float inv = empty ? 0.f : 1.f / l_run;
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_o + c * 32, ov);
                tmem_ld_wait();
                if (grow < len_q) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 w;
                        w.x = pack2<T>(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
                        w.y = pack2<T>(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
                        w.z = pack2<T>(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
                        w.w = pack2<T>(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
                        *reinterpret_cast<uint4 *>(og + c * 32 + i * 8) = w;
                    }
                }
            }
            if (p.lse && grow < len_q)
                p.lse[lse_idx] = empty ? INFINITY : m_ref * 0.6931471805599453f + __logf(l_run);
        } else if (grow < len_q) {
            // No visible key for any row of this tile: O = 0, LSE = +inf (flash_fwd_kernel.h:543-582).
#pragma unroll
            for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4 *>(og + c * 8) = make_uint4(0, 0, 0, 0);
            if (p.lse) p.lse[lse_idx] = INFINITY;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
struct Map4Key {
    const void *ptr;
    int64_t hs, rs, bs, rows, batches;
    int d, heads, bf16;
    bool operator==(const Map4Key &o) const { return std::memcmp(this, &o, sizeof(Map4Key)) == 0; }
};
struct Map4Entry {
    Map4Key key;
    CUtensorMap map;
};

// [batches][rows][heads][d] view with element strides; box = 128 rows x 64 d of one head.
bool make_attn_map(CUtensorMap *out, const void *ptr, int d, int heads, int64_t rows, int64_t batches,
                   int64_t hs, int64_t rs, int64_t bs, bool is_bf16) {
    static thread_local std::vector<Map4Entry> cache;
    Map4Key key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.hs = hs; key.rs = rs; key.bs = bs; key.rows = rows; key.batches = batches;
    key.d = d; key.heads = heads; key.bf16 = is_bf16;
    for (const Map4Entry &e : cache)
        if (e.key == key) { *out = e.map; return true; }
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(heads),
                                static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batches)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(hs) * 2, static_cast<cuuint64_t>(rs) * 2,
                                   static_cast<cuuint64_t>(bs) * 2};
    const cuuint32_t box[4] = {64, 1, 128, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_status(kErrDriver, "prefill: cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
        return false;
    }
    if (cache.size() >= 256) cache.clear();
    cache.push_back(Map4Entry{key, *out});
    return true;
}

template <typename T, int D>
bool launch_prefill(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const MhaParams &p,
                    cudaStream_t stream) {
    constexpr size_t smem = PfCfg<D>::kSmem;
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev != di.ordinal) {
        cudaError_t e = cudaFuncSetAttribute(prefill_attn_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) {
            set_status(static_cast<int>(e), "prefill: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return false;
        }
        attr_dev = di.ordinal;
    }
    dim3 grid((p.seqlen_q + kQTiles * kPfM - 1) / (kQTiles * kPfM), p.h, p.b);
    prefill_attn_kernel<T, D><<<grid, kPfThreads, smem, stream>>>(mq, mk, mv, p);
    note_launch();
    return check_launch("prefill_attn_kernel");
}

}  // namespace

// 1 = launched, 0 = outside this kernel's envelope (caller uses the general tile kernel), -1 = error.
int try_launch_prefill_tc(const MhaParams &p, bool is_bf16, cudaStream_t stream) {
    if (p.block_table || p.alibi) return 0;
    if (p.d != 64 && p.d != 128) return 0;
    if (p.seqlen_q < 32) return 0;  // a 128-row tile would be mostly padding
    if (!(p.scale_log2 > 0.f)) return 0;
    if (p.h_k <= 0 || p.h % p.h_k != 0 || p.h > 65535 || p.b > 65535) return 0;
    const DeviceInfo &di = current_device_info();
    if (di.cc_major != 10) return 0;
    // TMA: 16-byte aligned bases and strides (run_mha already checked pointers and strides % 8)
    const bool q_varlen = p.cu_q != nullptr;
    const bool k_varlen = p.cu_k != nullptr && p.k_cumulative;
    const int64_t huge_rows = 1ll << 26;
    auto stride_ok = [](int64_t s) { return s > 0 && s * 2 < (1ll << 40); };
    if (!stride_ok(p.q_rs) || !stride_ok(p.k_rs) || !stride_ok(p.v_rs) || !stride_ok(p.q_hs) ||
        !stride_ok(p.k_hs) || !stride_ok(p.v_hs))
        return 0;
    if (!q_varlen && p.b > 1 && !stride_ok(p.q_bs)) return 0;
    if (!k_varlen && p.b > 1 && (!stride_ok(p.k_bs) || !stride_ok(p.v_bs))) return 0;

    CUtensorMap mq, mk, mv;
    const int64_t q_rows = q_varlen ? huge_rows : p.seqlen_q, q_b = q_varlen ? 1 : p.b;
    const int64_t k_rows = k_varlen ? huge_rows : p.seqlen_k, k_b = k_varlen ? 1 : p.b;
    const int64_t q_bs = (q_varlen || p.b == 1 || p.q_bs == 0) ? p.q_rs : p.q_bs;
    const int64_t k_bs = (k_varlen || p.b == 1 || p.k_bs == 0) ? p.k_rs : p.k_bs;
    const int64_t v_bs = (k_varlen || p.b == 1 || p.v_bs == 0) ? p.v_rs : p.v_bs;
    if (!make_attn_map(&mq, p.q, p.d, p.h, q_rows, q_b, p.q_hs, p.q_rs, q_bs, is_bf16)) return -1;
    if (!make_attn_map(&mk, p.k, p.d, p.h_k, k_rows, k_b, p.k_hs, p.k_rs, k_bs, is_bf16)) return -1;
    if (!make_attn_map(&mv, p.v, p.d, p.h_k, k_rows, k_b, p.v_hs, p.v_rs, v_bs, is_bf16)) return -1;
    bool ok;
    if (is_bf16) ok = p.d == 128 ? launch_prefill<__nv_bfloat16, 128>(mq, mk, mv, p, stream)
                                 : launch_prefill<__nv_bfloat16, 64>(mq, mk, mv, p, stream);
    else ok = p.d == 128 ? launch_prefill<__half, 128>(mq, mk, mv, p, stream)
                         : launch_prefill<__half, 64>(mq, mk, mv, p, stream);
    return ok ? 1 : -1;
}

}  // namespace atoma
