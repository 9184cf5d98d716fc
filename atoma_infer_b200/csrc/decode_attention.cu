// Paged single-token decode attention for sm_100a (B200).
//
// Replaces, for run_mha calls with a block table and seqlen_q == 1, the reference's
// flash_fwd_splitkv_kernel + flash_fwd_splitkv_combine_kernel pair
// (csrc/kernels/flash_fwd_kernel.h:504-1092, :1131-1313, launched from
// flash_fwd_launch_template.h:105-177) — "paged attention v1/v2" in BASELINE.json's wording.
// Same contract: O[b,h,:] = softmax(scale * q[b,h,:] . K_b[:, h/G, :]^T) V_b over the first
// seqlens_k[b] cached tokens, gathered through block_table[b, :]; fp32 scores/softmax, P rounded
// to the 16-bit type before P.V, fp32 accumulation (flash_fwd_kernel.h:949-954).
//
// Design (nothing here is taken from the reference's FA2 kernel):
//  * The work is the flat list of "stage units": (sequence, kv-head group, 128 cache rows).  A
//    persistent grid of min(#SM, units) CTAs takes equal contiguous slices of that list
//    (stream-K), so the 1.73-wave imbalance of one-CTA-per-(sequence, head) never appears and no
//    caller-chosen num_splits is needed.
//  * A producer warpgroup streams K and V with TMA (cp.async.bulk.tensor, 128-byte swizzle) straight
//    out of the reference's [num_blocks, block_size, num_kv_heads, head_size] cache layout.  For
//    D = 128 the cache is described as a 5-D tensor {64 d, token, 2 halves of d, kv head, page}, so
//    ONE op moves [HKT heads][2 halves][TB tokens][64 d] of a physical page — for Llama-3.1-8B one
//    whole contiguous 32 KiB K page (all 8 kv heads).  Measured (scripts/tma_probe.cu): an SM
//    sustains only ~5 tensor ops/us from one issuing warp (~10 from four), so the op size decides
//    the bandwidth: 16-32 KiB ops reach the HBM limit from one warp, the 4 KiB ops of a
//    tensor-parallel shard with one kv head need all four producer warps.  The ops of a stage are
//    dealt round-robin to the four warps.  Page ids are prefetched 32 at a time from the block
//    table with one coalesced load and never past ceil(len / page) (the reference reads up to 7
//    entries past it).
//  * Eight consumer warps each own one (kv head, 16-token) slice of a stage and run the GQA group
//    of that kv head as the M dimension of mma.sync m16n8k16 (all H/Hkv query heads share one pass
//    over K/V; the reference re-streams K/V once per query head).
//  * Split partials (fp32 O, running max, running sum per warp) go to a library-owned L2-resident
//    workspace; the last CTA to arrive for a (sequence, head group) reduces them and writes the
//    16-bit output (+LSE) — one launch, no combine kernel.
#include "common.cuh"
#include "mha_params.h"

#include <cstring>
#include <vector>

namespace atoma {
namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kProducerWarps = 4;
constexpr int kDecodeThreads = kConsumerThreads + 32 * kProducerWarps;  // + the producer warpgroup
constexpr int kProducerRegs = 56;    // setmaxnreg: the producer warpgroup hands its registers ...
constexpr int kConsumerRegs = 216;   // ... to the two consumer warpgroups (168 + 112 * 128 / 256)
constexpr int kStageRows = 128;                        // (token, head) rows per stage
constexpr int kMaxBatch = 2047;
constexpr int kMaxCounters = 16384;                    // B * head groups
constexpr int kPartRows = 16;                          // M of the mma tile
constexpr int kMaxCtas = 256;                          // persistent grid size limit
constexpr int kCombineBatch = 8;                       // sub-partials loaded per combine step

template <int D>
struct Cfg {
    static constexpr int kHalves = D / 64;              // 128-byte column halves per row
    static constexpr int kHalfBytes = kStageRows * 128; // 16 KiB
    static constexpr int kTileBytes = kHalves * kHalfBytes;  // K (or V) bytes per stage
    static constexpr int kStageBytes = 2 * kTileBytes;
    static constexpr int kStages = (192 * 1024) / kStageBytes;  // 3 (D=128) / 6 (D=64)
    // In-CTA merge of the token-slice warps of one kv head (tensor-parallel shards, hkt < 8): up to four
    // writers per round, each G <= 8 rows of fp32 O in register order + (max, sum) per row.
    static constexpr int kTreeOBytes = 4 * (D / 8) * 32 * 8;
    static constexpr int kTreeBytes = kTreeOBytes + 4 * 8 * 8;
};

struct DecodeArgs {
    MhaParams p;
    int hkt;         // kv heads per CTA tile (1, 2, 4, 8)
    int hg;          // kv head groups = h_k / hkt
    int tok;         // tokens per stage = 128 / hkt
    int tb;          // tokens per TMA box (divides page and tok, multiple of 16)
    int g;           // query heads per kv head
    int num_ctas;
    int early;       // PDL + the caller's per-call promise: stream older pages before the grid dependency resolves
    int late_trigger;  // a cache-maintenance op is in flight on the stream: wait BEFORE releasing dependents
    int merged;      // map_k describes K AND V (one allocation, contiguous heads): one op moves both
    int *counters;   // [b * hg], zero between launches
    float *ws_ml;    // [cta][2][8 warps][16 rows][2]
    float *ws_o;     // [cta][2][8 warps][16 rows][D]
    unsigned long long *trace;  // diagnostics (atoma_debug_decode_trace): [cta][4] or null
    // what the next kernel of the stream streams first (o_proj's weights in a decode step): prefetched into L2
    // by the producer warps once their own last load is in flight, i.e. under this kernel's split-reduction tail
    const unsigned char *pf_ptr;
    long long pf_bytes;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
__device__ __forceinline__ unsigned sm_id() {
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

__device__ __forceinline__ int seq_len_k(const MhaParams &p, int b) {
    int len;
    if (p.seqused_k) len = p.seqused_k[b];
    else if (p.cu_k) len = p.k_cumulative ? p.cu_k[b + 1] - p.cu_k[b] : p.cu_k[b];
    else len = p.seqlen_k;
    return max(0, min(len, p.seqlen_k));
}

__device__ __forceinline__ long long cta_start(long long units, int cta, int num_ctas) {
    return units * cta / num_ctas;
}
// Largest c with cta_start(c) <= u.
__device__ __forceinline__ int cta_of_unit(long long units, long long u, int num_ctas) {
    return static_cast<int>(((u + 1) * num_ctas + units - 1) / units - 1);
}

// Reduction of every partial of one (sequence, kv head) by ONE warp (the last to arrive for that
// head).  Partials are merged in CTA-index order whoever arrives last, four outputs per lane and
// four partials in flight at a time (16-byte loads), so the result is bit-reproducible and the cost
// is a few L2 round trips.
template <typename T, int D>
__device__ __forceinline__ void combine_head(const DecodeArgs &a, long long units, int ncta, int fc, int lc,
                                             long long ub, long long ue, int b, int kvh, int hh, int G,
                                             int lane) {
    const MhaParams &p = a.p;
    constexpr int C4 = D / 4;
    const int nparts = lc - fc + 1;
    const int total = G * C4;
    for (int g0 = 0; g0 < total; g0 += 128) {
        bool ok[4];
        int rr[4], d4[4];
        float m_run[4], l_run[4];
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = g0 + lane + 32 * k;
            ok[k] = idx < total;
            rr[k] = ok[k] ? idx / C4 : 0;
            d4[k] = idx % C4;
            m_run[k] = -INFINITY;
            l_run[k] = 0.f;
            acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int pc = 0; pc < nparts; pc += 32) {
            int mybase = 0;  // lane j: workspace index of contributor pc + j
            if (pc + lane < nparts) {
                const int c = fc + pc + lane;
                const long long cs = cta_start(units, c, ncta);
                mybase = (c * 2 + ((cs >= ub && cs < ue) ? 0 : 1)) * kConsumerWarps + hh;
            }
            const int nchunk = min(32, nparts - pc);
            for (int s0 = 0; s0 < nchunk; s0 += 4) {
                float2 ml[4][4];
                float4 ov[4][4];
#pragma unroll
                for (int ss = 0; ss < 4; ++ss) {
                    // Loads are unconditional (clamped to a valid partial) so that all 32 of them are
                    // in flight together; out-of-range ones get weight exp2(-inf) = 0 below.
                    const int64_t base = __shfl_sync(0xffffffffu, mybase, min(s0 + ss, nchunk - 1));
                    const bool pv = s0 + ss < nchunk;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ml[ss][k] = __ldcg(reinterpret_cast<const float2 *>(a.ws_ml + base * (kPartRows * 2) + rr[k] * 2));
                        ov[ss][k] = __ldcg(reinterpret_cast<const float4 *>(a.ws_o + base * (kPartRows * D) + rr[k] * D + d4[k] * 4));
                        if (!(pv && ok[k])) ml[ss][k].x = -INFINITY;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float m_new = m_run[k];
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) m_new = fmaxf(m_new, ml[ss][k].x);
                    if (m_new == -INFINITY) continue;  // nothing but empty partials so far
                    const float c0 = fast_exp2(m_run[k] - m_new);  // m_run = -inf -> 0
                    acc[k].x *= c0; acc[k].y *= c0; acc[k].z *= c0; acc[k].w *= c0;
                    l_run[k] *= c0;
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        const float wgt = fast_exp2(ml[ss][k].x - m_new);  // empty partial: exp2(-inf) = 0
                        acc[k].x += wgt * ov[ss][k].x;
                        acc[k].y += wgt * ov[ss][k].y;
                        acc[k].z += wgt * ov[ss][k].z;
                        acc[k].w += wgt * ov[ss][k].w;
                        l_run[k] += wgt * ml[ss][k].y;
                    }
                    m_run[k] = m_new;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
            const bool empty = (l_run[k] == 0.f || l_run[k] != l_run[k]);
            const float inv = empty ? 1.f : 1.f / l_run[k];
            const int head = kvh * G + rr[k];
            T *op = static_cast<T *>(p.o) + b * p.o_bs + static_cast<int64_t>(head) * p.o_hs + d4[k] * 4;
            uint2 packed;
            packed.x = pack2<T>(acc[k].x * inv, acc[k].y * inv);
            packed.y = pack2<T>(acc[k].z * inv, acc[k].w * inv);
            *reinterpret_cast<uint2 *>(op) = packed;
            if (d4[k] == 0 && p.lse)
                p.lse[static_cast<int64_t>(b) * p.h + head] =
                    empty ? INFINITY : m_run[k] * 0.6931471805599453f + __logf(l_run[k]);
        }
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(kDecodeThreads, 1)
decode_paged_kernel(const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const DecodeArgs a) {
    using C = Cfg<D>;
    const MhaParams &p = a.p;
    extern __shared__ unsigned char smem_dyn[];
    // 128-byte-swizzled TMA destinations need 1024-byte alignment.
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_k[C::kStages], full_v[C::kStages], empty[C::kStages];
    __shared__ int cum[kMaxBatch + 2];  // cum[b] = first unit of sequence b
    __shared__ int combine_flag;
    __shared__ int part_base[kMaxCtas];  // combine: workspace index of each contributing CTA's partial

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.b;
    const unsigned long long t_start = a.trace ? global_timer_ns() : 0ull;

    if (tid == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            mbar_init(&full_k[i], kProducerWarps);
            mbar_init(&full_v[i], kProducerWarps);
            mbar_init(&empty[i], kConsumerWarps);
        }
        fence_mbar_init();
    }
    if (warp == kConsumerWarps && lane == 0) {
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
    }
    // Programmatic dependent launch: everything above overlaps the previous kernel's tail; nothing
    // below (sequence lengths, block table, q, caches, workspace, outputs) is touched before the
    // previous kernel in the stream has completed — except, in "stable cache" mode, what that kernel
    // cannot be writing: lengths, block tables and every page but the final one of a sequence.  The
    // producer warps then run up to a full ring (192 KiB per SM) ahead of the dependency.
    if (a.late_trigger) {
        // a cache-maintenance op precedes this call: wait first, and only then let dependents start
        pdl_wait();
        pdl_launch_dependents();
    } else {
        pdl_launch_dependents();
        if (!a.early) pdl_wait();
    }
    if (a.trace && tid == 0) {
        a.trace[blockIdx.x * 8 + 0] = sm_id();
        a.trace[blockIdx.x * 8 + 1] = t_start;
        a.trace[blockIdx.x * 8 + 2] = 0;
        a.trace[blockIdx.x * 8 + 3] = 0;
    }
    // units per sequence -> exclusive prefix sum in cum[]
    for (int i = tid; i < B; i += kDecodeThreads) {
        const int len = seq_len_k(p, i);
        cum[i + 1] = a.hg * ((len + a.tok - 1) / a.tok);
    }
    if (tid == 0) cum[0] = 0;
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < B; base += 32) {
            const int i = base + lane;
            int v = i < B ? cum[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            if (i < B) cum[i + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    const long long units = cum[B];
    const int cta = blockIdx.x;

    // Sequences without any cached token: O = 0, LSE = +inf (flash_fwd_kernel.h:543-582).  Consumer
    // threads only: they are the ones that wait for the dependency before anything is written.
    if (warp < kConsumerWarps) {
        if (a.early) pdl_wait();
        for (int b = cta; b < B; b += gridDim.x) {
            if (cum[b + 1] != cum[b]) continue;
            T *ob = static_cast<T *>(p.o) + b * p.o_bs;
            for (int i = tid; i < p.h * (D / 8); i += kConsumerThreads) {
                const int hh = i / (D / 8), c = i % (D / 8);
                *reinterpret_cast<uint4 *>(ob + hh * p.o_hs + c * 8) = make_uint4(0, 0, 0, 0);
            }
            if (p.lse)
                for (int i = tid; i < p.h; i += kConsumerThreads) p.lse[static_cast<int64_t>(b) * p.h + i] = INFINITY;
        }
    }

    // The grid was sized from the host-side upper bound of the unit count; only the first
    // min(grid, units) CTAs take part in the partition, so that every participant owns >= 1 unit
    // (the arrival-count logic below relies on it).
    const int ncta = static_cast<int>(min(static_cast<long long>(gridDim.x), units));
    if (cta >= ncta) return;
    const long long u_begin = cta_start(units, cta, ncta), u_end = cta_start(units, cta + 1, ncta);

    // One box = [hkt heads][kHalves][tb tokens] 128-byte lines (one tensor op).
    const int box_rows = a.hkt * C::kHalves * a.tb;
    const uint32_t box_bytes = box_rows * 128u;
    const int boxes_per_stage = a.tok / a.tb;
    const uint32_t half_stride = a.tb * 128u;     // bytes between the two 64-wide halves of a row

    int stage = 0;
    uint32_t phase = 0;

    if (warp >= kConsumerWarps) {
        // =========================== producer warpgroup ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        const int pw = warp - kConsumerWarps;  // ops o = 0..2*nbx-1 of a stage (K boxes, then V boxes): warp o % 4
        bool waited = !a.early;
        long long u = u_begin;
        int b = 0;
        while (u < u_end) {
            while (cum[b + 1] <= u) ++b;
            const int ns = (cum[b + 1] - cum[b]) / a.hg;
            const int r = static_cast<int>(u - cum[b]);
            const int hg = r / ns, s0 = r % ns;
            const int n = static_cast<int>(min(static_cast<long long>(ns - s0), u_end - u));
            const int len = seq_len_k(p, b);
            const int npages = (len + p.page - 1) / p.page;
            const int *bt = p.block_table + b * p.bt_bs;
            int pbase = -(1 << 30), pg_cache = 0;
            for (int s = s0; s < s0 + n; ++s) {
                if (!waited && s == ns - 1) {
                    // the final stage of a sequence holds the page the predecessor is writing
                    pdl_wait();
                    waited = true;
                }
                mbar_wait(&empty[stage], phase ^ 1);
                __syncwarp();
                const int tok0 = s * a.tok;
                int nbx = (len - tok0 + a.tb - 1) / a.tb;
                nbx = min(nbx, boxes_per_stage);
                unsigned char *ks = tiles + stage * C::kStageBytes;
                unsigned char *vs = ks + C::kTileBytes;
                int phys[8];
#pragma unroll
                for (int bx = 0; bx < 8; ++bx) {
                    phys[bx] = 0;
                    if (bx < nbx) {  // warp-uniform
                        const int pi = (tok0 + bx * a.tb) / p.page;
                        if (pi < pbase || pi >= pbase + 32) {
                            pbase = pi;
                            pg_cache = (pbase + lane < npages) ? bt[pbase + lane] : 0;
                        }
                        phys[bx] = __shfl_sync(0xffffffffu, pg_cache, pi - pbase);
                    }
                }
                if (lane == 0 && a.merged) {
                    // K and V of a page in ONE op: box bx goes to warp bx % 4; smem [box][K|V][head][half][token]
                    int nmine = 0;
#pragma unroll
                    for (int bx = 0; bx < 8; ++bx) nmine += (bx < nbx && (bx & 3) == pw) ? 1 : 0;
                    mbar_arrive_expect_tx(&full_k[stage], nmine * 2 * box_bytes);
#pragma unroll
                    for (int bx = 0; bx < 8; ++bx) {
                        if (bx < nbx && (bx & 3) == pw) {
                            const int t_in_page = (tok0 + bx * a.tb) % p.page;
                            tma_load_5d(ks + bx * 2 * box_bytes, &map_k, &full_k[stage], 0, t_in_page,
                                        hg * a.hkt * C::kHalves, 0, phys[bx], kL2EvictFirst);
                        }
                    }
                    mbar_arrive(&full_v[stage]);  // V came with K: nothing more to expect
                } else if (lane == 0) {
                    // this warp's share: K box bx when bx % 4 == pw, V box bx when (nbx + bx) % 4 == pw
                    int nk = 0, nv = 0;
#pragma unroll
                    for (int bx = 0; bx < 8; ++bx) {
                        nk += (bx < nbx && (bx & 3) == pw) ? 1 : 0;
                        nv += (bx < nbx && ((nbx + bx) & 3) == pw) ? 1 : 0;
                    }
                    mbar_arrive_expect_tx(&full_k[stage], nk * box_bytes);
#pragma unroll
                    for (int bx = 0; bx < 8; ++bx) {
                        if (bx < nbx && (bx & 3) == pw) {
                            const int t_in_page = (tok0 + bx * a.tb) % p.page;
                            if (C::kHalves == 2)
                                tma_load_5d(ks + bx * box_bytes, &map_k, &full_k[stage], 0, t_in_page, 0,
                                            hg * a.hkt, phys[bx], kL2EvictFirst);
                            else
                                tma_load_4d(ks + bx * box_bytes, &map_k, &full_k[stage], 0, t_in_page,
                                            hg * a.hkt, phys[bx], kL2EvictFirst);
                        }
                    }
                    mbar_arrive_expect_tx(&full_v[stage], nv * box_bytes);
#pragma unroll
                    for (int bx = 0; bx < 8; ++bx) {
                        if (bx < nbx && ((nbx + bx) & 3) == pw) {
                            const int t_in_page = (tok0 + bx * a.tb) % p.page;
                            if (C::kHalves == 2)
                                tma_load_5d(vs + bx * box_bytes, &map_v, &full_v[stage], 0, t_in_page, 0,
                                            hg * a.hkt, phys[bx], kL2EvictFirst);
                            else
                                tma_load_4d(vs + bx * box_bytes, &map_v, &full_v[stage], 0, t_in_page,
                                            hg * a.hkt, phys[bx], kL2EvictFirst);
                        }
                    }
                }
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            u += n;
        }
        if (a.pf_bytes > 0 && lane == 0) {
            const long long pieces = l2_prefetch_pieces(a.pf_bytes);
            const int part = cta * kProducerWarps + pw, parts = ncta * kProducerWarps;
            l2_prefetch_range(a.pf_ptr, a.pf_bytes, pieces * part / parts, pieces * (part + 1) / parts);
        }
        return;
    }

    // ============================== consumer warps ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    const int g = lane >> 2, t4 = lane & 3;
    const int hh = warp % a.hkt;          // kv head within the group
    const int ts = warp / a.hkt;          // 16-token slice within the stage
    const int bx_w = (ts * 16) / a.tb;    // box holding this warp's tokens
    // 128-byte line index of this warp's first token, first half: box layout [head][half][token]
    const int row_base = hh * C::kHalves * a.tb + (ts * 16) % a.tb;   // ... inside the box
    // byte offsets of this warp's K and V boxes inside a stage
    const uint32_t k_off = a.merged ? bx_w * 2 * box_bytes : bx_w * box_bytes;
    const uint32_t v_off = a.merged ? k_off + box_bytes : C::kTileBytes + bx_w * box_bytes;
    const int G = a.g;

    long long u = u_begin;
    int b = 0;
    while (u < u_end) {
        while (cum[b + 1] <= u) ++b;
        const int ns = (cum[b + 1] - cum[b]) / a.hg;
        const int r = static_cast<int>(u - cum[b]);
        const int hg = r / ns, s0 = r % ns;
        const int n = static_cast<int>(min(static_cast<long long>(ns - s0), u_end - u));
        const int len = seq_len_k(p, b);
        const int kvh = hg * a.hkt + hh;

        // ---- Q fragments (A operand): rows = the G query heads of kv head `kvh` ----
        uint32_t qf[D / 16][4];
        {
            const T *qb = static_cast<const T *>(p.q) + b * p.q_bs;
            const T *q0 = qb + static_cast<int64_t>(kvh * G + g) * p.q_hs;
            const T *q1 = qb + static_cast<int64_t>(kvh * G + g + 8) * p.q_hs;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const int d0 = kk * 16 + 2 * t4;
                qf[kk][0] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + d0) : 0u;
                qf[kk][1] = g + 8 < G ? *reinterpret_cast<const uint32_t *>(q1 + d0) : 0u;
                qf[kk][2] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + d0 + 8) : 0u;
                qf[kk][3] = g + 8 < G ? *reinterpret_cast<const uint32_t *>(q1 + d0 + 8) : 0u;
            }
        }
        float o_acc[D / 8][4];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
        float m_run[2] = {-INFINITY, -INFINITY};
        float l_run[2] = {0.f, 0.f};

        for (int s = s0; s < s0 + n; ++s) {
            const uint32_t ks = smem_u32(tiles + stage * C::kStageBytes) + k_off;
            const uint32_t vs = smem_u32(tiles + stage * C::kStageBytes) + v_off;
            const int nvalid = len - (s * a.tok + ts * 16);  // valid tokens in this warp's slice
            mbar_wait(&full_k[stage], phase);
            if (nvalid > 0) {
                // ---- S = Q K^T : 16 (heads, padded) x 16 tokens ----
                float sc[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                {
                    const int i = (lane >> 4) * 8 + (lane & 7);  // token within the slice
                    const uint32_t line = ks + (row_base + i) * 128;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const int c = kk * 2 + ((lane >> 3) & 1);  // 16-byte chunk over D
                        uint32_t bf[4];
                        ldmatrix_x4(bf, line + (c >> 3) * half_stride + (((c & 7) ^ (lane & 7)) << 4));
                        Mma16816<T>::run(sc[0], qf[kk], bf[0], bf[1]);
                        Mma16816<T>::run(sc[1], qf[kk], bf[2], bf[3]);
                    }
                }
                // ---- online softmax (log2 domain) ----
                float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int tokc = j * 8 + 2 * t4 + (e & 1);
                        float x = sc[j][e] * p.scale_log2;
                        x = tokc < nvalid ? x : -INFINITY;
                        sc[j][e] = x;
                        mx[e >> 1] = fmaxf(mx[e >> 1], x);
                    }
                }
                float corr[2], m_safe[2];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 1));
                    mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 2));
                    const float m_new = fmaxf(m_run[rr], mx[rr]);
                    m_safe[rr] = m_new == -INFINITY ? 0.f : m_new;
                    corr[rr] = fast_exp2(m_run[rr] - m_safe[rr]);
                    m_run[rr] = m_new;
                }
                float rs[2] = {0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float pv = fast_exp2(sc[j][e] - m_safe[e >> 1]);
                        sc[j][e] = pv;
                        rs[e >> 1] += pv;
                    }
                }
                l_run[0] = l_run[0] * corr[0] + rs[0];
                l_run[1] = l_run[1] * corr[1] + rs[1];
                if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
                    for (int i = 0; i < D / 8; ++i) {
                        o_acc[i][0] *= corr[0];
                        o_acc[i][1] *= corr[0];
                        o_acc[i][2] *= corr[1];
                        o_acc[i][3] *= corr[1];
                    }
                }
                uint32_t pa[4];
                pa[0] = pack2<T>(sc[0][0], sc[0][1]);
                pa[1] = pack2<T>(sc[0][2], sc[0][3]);
                pa[2] = pack2<T>(sc[1][0], sc[1][1]);
                pa[3] = pack2<T>(sc[1][2], sc[1][3]);

                mbar_wait(&full_v[stage], phase);
                if (nvalid < 16) {
                    // Ragged tail: cache rows past the sequence may hold anything (NaN included) and
                    // 0 * NaN would poison the accumulator: clear them (the reference zero-fills its
                    // V tile instead, flash_fwd_kernel.h:901-906).
                    for (int idx = lane; idx < 16 * (D / 8); idx += 32) {
                        const int i = idx / (D / 8), c = idx % (D / 8);
                        if (i >= nvalid)
                            *reinterpret_cast<uint4 *>(tiles + (vs - smem_u32(tiles)) +
                                                       (c >> 3) * half_stride + (row_base + i) * 128 +
                                                       (((c & 7) ^ (i & 7)) << 4)) = make_uint4(0, 0, 0, 0);
                    }
                    __syncwarp();
                }
                // ---- O += P V ----
                {
                    const int i = ((lane >> 3) & 1) * 8 + (lane & 7);
                    const uint32_t line = vs + (row_base + i) * 128;
#pragma unroll
                    for (int dj2 = 0; dj2 < D / 16; ++dj2) {
                        const int c = dj2 * 2 + (lane >> 4);
                        uint32_t vf[4];
                        ldmatrix_x4_trans(vf, line + (c >> 3) * half_stride + (((c & 7) ^ (lane & 7)) << 4));
                        Mma16816<T>::run(o_acc[2 * dj2], pa, vf[0], vf[1]);
                        Mma16816<T>::run(o_acc[2 * dj2 + 1], pa, vf[2], vf[3]);
                    }
                }
                if (nvalid < 16) fence_proxy_async();
            } else {
                mbar_wait(&full_v[stage], phase);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == C::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }

        if (a.trace && tid == 0) a.trace[blockIdx.x * 8 + 2] = global_timer_ns();
        // ---- publish this warp's partial ----
        const long long ub = cum[b] + static_cast<long long>(hg) * ns, ue = ub + ns;
        const int slot = (u_begin >= ub && u_begin < ue) ? 0 : 1;  // first segment of the CTA -> 0
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            l_run[rr] += __shfl_xor_sync(0xffffffffu, l_run[rr], 1);
            l_run[rr] += __shfl_xor_sync(0xffffffffu, l_run[rr], 2);
        }
        // Tensor-parallel shards put several token-slice warps on one kv head (hkt < 8).  Merge them
        // inside the CTA first (binary tree through shared memory, fixed order), so that one partial
        // per (CTA, kv head) is published and every head is reduced by one warp, exactly like hkt = 8.
        const bool tree = a.hkt < kConsumerWarps && G <= 8;
        if (tree) {
            float2 *sm_o = reinterpret_cast<float2 *>(tiles + C::kStages * C::kStageBytes);
            float2 *sm_ml = reinterpret_cast<float2 *>(tiles + C::kStages * C::kStageBytes + C::kTreeOBytes);
            const int wpk = kConsumerWarps / a.hkt;
            for (int step = 1; step < wpk; step <<= 1) {
                const int sl = (ts / (2 * step)) * a.hkt + hh;  // < 4
                if ((ts % (2 * step)) == step) {
#pragma unroll
                    for (int i = 0; i < D / 8; ++i)
                        sm_o[(sl * (D / 8) + i) * 32 + lane] = make_float2(o_acc[i][0], o_acc[i][1]);
                    if (t4 == 0) sm_ml[sl * 8 + g] = make_float2(m_run[0], l_run[0]);
                }
                named_bar_sync(1, kConsumerThreads);
                if ((ts % (2 * step)) == 0) {
                    const float2 mlb = sm_ml[sl * 8 + g];
                    const float m_new = fmaxf(m_run[0], mlb.x);
                    if (m_new != -INFINITY) {
                        const float ca = fast_exp2(m_run[0] - m_new), cb = fast_exp2(mlb.x - m_new);
#pragma unroll
                        for (int i = 0; i < D / 8; ++i) {
                            const float2 ob = sm_o[(sl * (D / 8) + i) * 32 + lane];
                            o_acc[i][0] = o_acc[i][0] * ca + ob.x * cb;
                            o_acc[i][1] = o_acc[i][1] * ca + ob.y * cb;
                        }
                        l_run[0] = l_run[0] * ca + mlb.y * cb;
                        m_run[0] = m_new;
                    }
                }
                named_bar_sync(1, kConsumerThreads);
            }
        }
        if (!tree || ts == 0) {
            const int64_t part = (static_cast<int64_t>(cta) * 2 + slot) * kConsumerWarps + warp;
            float *po = a.ws_o + part * (kPartRows * D);
            float *pml = a.ws_ml + part * (kPartRows * 2);
            if (g < G) {
#pragma unroll
                for (int dj = 0; dj < D / 8; ++dj)
                    __stcg(reinterpret_cast<float2 *>(po + g * D + dj * 8 + 2 * t4),
                           make_float2(o_acc[dj][0], o_acc[dj][1]));
                if (t4 == 0) __stcg(reinterpret_cast<float2 *>(pml + g * 2), make_float2(m_run[0], l_run[0]));
            }
            if (g + 8 < G) {
#pragma unroll
                for (int dj = 0; dj < D / 8; ++dj)
                    __stcg(reinterpret_cast<float2 *>(po + (g + 8) * D + dj * 8 + 2 * t4),
                           make_float2(o_acc[dj][2], o_acc[dj][3]));
                if (t4 == 0)
                    __stcg(reinterpret_cast<float2 *>(pml + (g + 8) * 2), make_float2(m_run[1], l_run[1]));
            }
        }
        const int fc = cta_of_unit(units, ub, ncta), lc = cta_of_unit(units, ue - 1, ncta);
        if (a.hkt == kConsumerWarps || tree) {
            // One partial per kv head (Hkv a multiple of 8, or after the in-CTA merge): every head is
            // published and reduced on its own — no CTA-wide barrier, and the reductions of a
            // sequence land on whichever warps arrive last, usually on different SMs.
            if (tree && ts != 0) {
                u += n;
                continue;
            }
            __threadfence();
            __syncwarp();
            if (a.trace && tid == 0) a.trace[blockIdx.x * 8 + 4] = global_timer_ns();
            int *ctr = &a.counters[b * p.h_k + kvh];
            int old = 0;
            if (lane == 0) old = atomicAdd(ctr, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (a.trace && tid == 0) a.trace[blockIdx.x * 8 + 5] = global_timer_ns();
            if (old == lc - fc) {
                __threadfence();
                if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + 6] = global_timer_ns();
                combine_head<T, D>(a, units, ncta, fc, lc, ub, ue, b, kvh, hh, G, lane);
                if (lane == 0) *ctr = 0;  // self-cleaning for the next launch
                if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + 7] = global_timer_ns();
            }
            u += n;
            continue;
        }
        __threadfence();
        named_bar_sync(1, kConsumerThreads);
        if (tid == 0) {
            const int old = atomicAdd(&a.counters[b * a.hg + hg], 1);
            combine_flag = (old == lc - fc) ? 1 : 0;
        }
        named_bar_sync(1, kConsumerThreads);
        if (combine_flag) {
            // ---- last arriver: reduce every partial of (b, hg) and write the output ----
            // Partials are merged in a fixed order (CTA index, then token slice), so the result does
            // not depend on which CTA arrives last.  Loads are issued kCombineBatch sub-partials at a
            // time, all independent, so the reduction costs a few L2 round trips, not one per partial.
            __threadfence();
            const int wpk = kConsumerWarps / a.hkt;  // warps (token slices) per kv head
            const int nparts = lc - fc + 1;
            for (int i = tid; i < nparts; i += kConsumerThreads) {
                const long long cs = cta_start(units, fc + i, ncta);
                const int sl = (cs >= ub && cs < ue) ? 0 : 1;
                part_base[i] = ((fc + i) * 2 + sl) * kConsumerWarps;
            }
            named_bar_sync(1, kConsumerThreads);
            const int nsub = nparts * wpk;
            const int total = a.hkt * G * (D / 4);
            for (int idx = tid; idx < total; idx += kConsumerThreads) {
                const int d4 = idx % (D / 4);
                const int rr = (idx / (D / 4)) % G;
                const int h2 = idx / (D / 4) / G;
                float m_run = -INFINITY, l_run = 0.f;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s0 = 0; s0 < nsub; s0 += kCombineBatch) {
                    float2 ml[kCombineBatch];
                    float4 ov[kCombineBatch];
#pragma unroll
                    for (int j = 0; j < kCombineBatch; ++j) {
                        // unconditional loads (clamped), masked afterwards: all in flight together
                        const int sidx = min(s0 + j, nsub - 1);
                        const int64_t part = part_base[sidx / wpk] + (sidx % wpk) * a.hkt + h2;
                        ml[j] = __ldcg(reinterpret_cast<const float2 *>(a.ws_ml + part * (kPartRows * 2) + rr * 2));
                        ov[j] = __ldcg(reinterpret_cast<const float4 *>(a.ws_o + part * (kPartRows * D) + rr * D + d4 * 4));
                        if (s0 + j >= nsub) ml[j].x = -INFINITY;
                    }
                    float m_new = m_run;
#pragma unroll
                    for (int j = 0; j < kCombineBatch; ++j) m_new = fmaxf(m_new, ml[j].x);
                    if (m_new == -INFINITY) continue;  // nothing but empty partials so far
                    const float c0 = fast_exp2(m_run - m_new);  // m_run = -inf -> 0
                    acc.x *= c0; acc.y *= c0; acc.z *= c0; acc.w *= c0;
                    l_run *= c0;
#pragma unroll
                    for (int j = 0; j < kCombineBatch; ++j) {
                        const float wgt = fast_exp2(ml[j].x - m_new);  // empty partial: exp2(-inf) = 0
                        acc.x += wgt * ov[j].x;
                        acc.y += wgt * ov[j].y;
                        acc.z += wgt * ov[j].z;
                        acc.w += wgt * ov[j].w;
                        l_run += wgt * ml[j].y;
                    }
                    m_run = m_new;
                }
                const bool empty = (l_run == 0.f || l_run != l_run);
                const float inv = empty ? 1.f : 1.f / l_run;
                const int head = (hg * a.hkt + h2) * G + rr;
                T *op = static_cast<T *>(p.o) + b * p.o_bs + static_cast<int64_t>(head) * p.o_hs + d4 * 4;
                uint2 packed;
                packed.x = pack2<T>(acc.x * inv, acc.y * inv);
                packed.y = pack2<T>(acc.z * inv, acc.w * inv);
                *reinterpret_cast<uint2 *>(op) = packed;
                if (d4 == 0 && p.lse)
                    p.lse[static_cast<int64_t>(b) * p.h + head] =
                        empty ? INFINITY : m_run * 0.6931471805599453f + __logf(l_run);
            }
            if (tid == 0) a.counters[b * a.hg + hg] = 0;  // self-cleaning for the next launch
        }
        named_bar_sync(1, kConsumerThreads);  // combine_flag / slot 1 may be reused by the next segment
        u += n;
    }
    if (a.trace && tid == 0) a.trace[blockIdx.x * 8 + 3] = global_timer_ns();
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct MapKey {
    const void *ptr;
    int64_t rs, hs, bs, kv_delta;
    int page, hk, d, tb, hkt, bf16;
    bool operator==(const MapKey &o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapEntry {
    MapKey key;
    CUtensorMap map;
};

// kv_delta > 0: ONE map over K and V (V = K + kv_delta bytes, heads contiguous: hs == d), see below.
bool make_kv_map(CUtensorMap *out, const void *ptr, int64_t rs, int64_t hs, int64_t bs, int page,
                 int hk, int d, int tb, int hkt, bool is_bf16, int64_t kv_delta = 0) {
    static thread_local std::vector<MapEntry> cache;
    MapKey key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.rs = rs; key.hs = hs; key.bs = bs; key.kv_delta = kv_delta; key.page = page; key.hk = hk; key.d = d;
    key.tb = tb; key.hkt = hkt; key.bf16 = is_bf16;
    for (const MapEntry &e : cache)
        if (e.key == key) { *out = e.map; return true; }
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    // dims, fastest first.  The page count is not part of the run_mha ABI; any id the block table can
    // hold must be addressable.
    //   D = 64 : {d, token-in-page, kv head, physical page}, box {64, tb, hkt, 1}
    //   D = 128: {64 d, token-in-page, half of d (stride 128 B), kv head, physical page},
    //            box {64, tb, 2, hkt, 1}: both 128-byte halves of every row in one op, landing as
    //            [head][half][token] lines (token & 7 stays the swizzle phase of a line).
    const CUtensorMapDataType dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    //   K and V in one allocation with contiguous heads (the reference's [2, NB, bs, Hkv, D] cache):
    //            {64 d, token-in-page, 128-byte line of the token row (= head x half, stride 128 B),
    //             K|V (stride V - K), physical page}, box {64, tb, lines of hkt heads, 2, 1}: K AND V
    //            of hkt heads of a page in one op, landing as [K|V][head][half][token] lines.
    CUresult r;
    if (kv_delta > 0) {
        const int halves = d / 64;
        const cuuint64_t dims[5] = {64, static_cast<cuuint64_t>(page), static_cast<cuuint64_t>(hk * halves), 2,
                                    static_cast<cuuint64_t>(1) << 22};
        const cuuint64_t strides[4] = {static_cast<cuuint64_t>(rs) * 2, 128, static_cast<cuuint64_t>(kv_delta),
                                       static_cast<cuuint64_t>(bs) * 2};
        const cuuint32_t box[5] = {64, static_cast<cuuint32_t>(tb), static_cast<cuuint32_t>(hkt * halves), 2, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        r = enc(out, dt, 5, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (d == 128) {
        const cuuint64_t dims[5] = {64, static_cast<cuuint64_t>(page), 2, static_cast<cuuint64_t>(hk),
                                    static_cast<cuuint64_t>(1) << 22};
        const cuuint64_t strides[4] = {static_cast<cuuint64_t>(rs) * 2, 128, static_cast<cuuint64_t>(hs) * 2,
                                       static_cast<cuuint64_t>(bs) * 2};
        const cuuint32_t box[5] = {64, static_cast<cuuint32_t>(tb), 2, static_cast<cuuint32_t>(hkt), 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        r = enc(out, dt, 5, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(page),
                                    static_cast<cuuint64_t>(hk), static_cast<cuuint64_t>(1) << 22};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(hs) * 2,
                                       static_cast<cuuint64_t>(bs) * 2};
        const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(tb), static_cast<cuuint32_t>(hkt), 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        r = enc(out, dt, 4, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        set_status(kErrDriver, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
        return false;
    }
    if (cache.size() >= 256) cache.clear();
    cache.push_back(MapEntry{key, *out});
    return true;
}

// TMA box dimensions are limited to 256 elements each.
inline bool box_lines_ok(int hkt, int d) { return hkt * (d / 64) <= 256; }

void *g_decode_trace = nullptr;
int g_decode_no_merge = 0;  // diagnostics (atoma_debug_decode_merge): 1 = always two tensor maps

template <typename T, int D>
bool launch_decode(const CUtensorMap &mk, const CUtensorMap &mv, const DecodeArgs &a,
                   cudaStream_t stream) {
    constexpr size_t smem = Cfg<D>::kStages * Cfg<D>::kStageBytes + 1024 + Cfg<D>::kTreeBytes;
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev != di.ordinal) {
        cudaFuncSetAttribute(decode_paged_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
        // the same (maximum) carve-out as every other kernel of the decode step: see gemm_skinny.cu
        cudaFuncSetAttribute(decode_paged_kernel<T, D>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        attr_dev = di.ordinal;
    }
    cudaError_t e = launch_ex(decode_paged_kernel<T, D>, dim3(a.num_ctas), dim3(kDecodeThreads), smem, stream, mk, mv, a);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "decode_paged_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("decode_paged_kernel");
}

}  // namespace

// Returns 1 if the decode kernel took the call, 0 if the configuration is outside its envelope
// (caller falls back to the general tile kernel), -1 on error (status set).
int try_launch_decode_paged(const MhaParams &p, bool is_bf16, int64_t max_units_hint,
                            cudaStream_t stream) {
    if (!p.block_table || p.seqlen_q != 1 || p.alibi || p.cu_q) return 0;
    if (p.d != 64 && p.d != 128) return 0;
    if (p.h_k <= 0 || p.h % p.h_k != 0) return 0;
    const int G = p.h / p.h_k;
    if (G > 16 || p.b > kMaxBatch || p.b <= 0) return 0;
    if (p.page <= 0 || p.page % 16 != 0) return 0;
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.k) || !al16(p.v) || !al16(p.o)) return 0;
    if ((reinterpret_cast<uintptr_t>(p.q) & 3) != 0) return 0;
    if (p.k_rs % 8 || p.k_hs % 8 || p.k_bs % 8 || p.v_rs % 8 || p.v_hs % 8 || p.v_bs % 8) return 0;
    if (p.q_hs % 2 || p.q_bs % 2 || p.o_hs % 8 || p.o_bs % 8) return 0;
    if (p.k_rs * 2 >= (1ll << 40) || p.k_bs * 2 >= (1ll << 40)) return 0;

    int hkt = 8;
    while (hkt > 1 && p.h_k % hkt != 0) hkt >>= 1;
    const int hg = p.h_k / hkt;
    if (static_cast<int64_t>(p.b) * p.h_k > kMaxCounters) return 0;
    const int tok = kStageRows / hkt;
    int tb = tok;
    while (tb > 16 && p.page % tb != 0) tb >>= 1;
    if (tok / tb > 8) return 0;

    const DeviceInfo &di = current_device_info();
    if (di.cc_major < 9) {
        set_status(kErrUnsupported, "decode kernel needs TMA (sm_90+), device is sm_%d%d", di.cc_major, di.cc_minor);
        return -1;
    }
    DecodeArgs a;
    a.p = p;
    a.hkt = hkt; a.hg = hg; a.tok = tok; a.tb = tb; a.g = G;
    // early KV streaming needs PDL and the caller's promise for THIS call (or the legacy process-wide bit)
    const int stable_call = take_next_decode_stable_cache();
    const bool dirty = take_cache_dirty(stream);   // copy_blocks / swap-in since the last decode on this stream
    a.early = ((launch_flags() & kLaunchPdl) && stable_call && !dirty) ? 1 : 0;
    a.late_trigger = dirty ? 1 : 0;
    note_decode_mode(a.early | (a.late_trigger << 1));
    // Upper bound on units is known on the host only through seqlen_k (= pages * page size); the
    // kernel derives the real count from the device-side lengths.  Grid: one CTA per SM, but never
    // more CTAs than the upper bound of units.
    const int64_t units_ub = static_cast<int64_t>(p.b) * hg * ((p.seqlen_k + tok - 1) / tok);
    (void)max_units_hint;
    a.num_ctas = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(di.sm_count, kMaxCtas), units_ub)));

    const size_t counters_bytes = kMaxCounters * sizeof(int);
    const size_t ml_bytes = static_cast<size_t>(di.sm_count) * 2 * kConsumerWarps * kPartRows * 2 * sizeof(float);
    const size_t o_bytes = static_cast<size_t>(di.sm_count) * 2 * kConsumerWarps * kPartRows * 128 * sizeof(float);
    char *ws = static_cast<char *>(get_workspace(stream, counters_bytes + ml_bytes + o_bytes, kWorkspaceDecode));
    if (!ws) return -1;
    a.counters = reinterpret_cast<int *>(ws);
    a.ws_ml = reinterpret_cast<float *>(ws + counters_bytes);
    a.ws_o = reinterpret_cast<float *>(ws + counters_bytes + ml_bytes);
    a.trace = static_cast<unsigned long long *>(g_decode_trace);
    if (!a.trace) a.trace = next_step_trace_slot("decode_paged");
    {
        const NextPrefetch pf = take_next_kernel_prefetch();
        const bool pf_ok = pf.ptr && pf.bytes >= 16 && (reinterpret_cast<uintptr_t>(pf.ptr) & 15) == 0;
        a.pf_ptr = pf_ok ? static_cast<const unsigned char *>(pf.ptr) : nullptr;
        a.pf_bytes = pf_ok ? pf.bytes : 0;
    }

    // K and V views of one [2, NB, bs, Hkv, D] allocation (flash_attention.rs:239-246, 274-275): one
    // tensor map over both, one op per page for K and V together (op size x2 — what decides the
    // per-SM bandwidth of tensor-parallel shards, see the header).  Anything else: two maps.
    const int64_t kv_delta = static_cast<const char *>(p.v) - static_cast<const char *>(p.k);
    const bool merged = p.k_hs == p.d && p.v_hs == p.d && p.k_rs == p.v_rs && p.k_bs == p.v_bs && kv_delta > 0 &&
                        kv_delta % 16 == 0 && kv_delta < (1ll << 40) && box_lines_ok(hkt, p.d) && !g_decode_no_merge;
    a.merged = merged ? 1 : 0;
    CUtensorMap mk, mv;
    if (merged) {
        if (!make_kv_map(&mk, p.k, p.k_rs, p.k_hs, p.k_bs, p.page, p.h_k, p.d, tb, hkt, is_bf16, kv_delta)) return -1;
        mv = mk;
    } else {
        if (!make_kv_map(&mk, p.k, p.k_rs, p.k_hs, p.k_bs, p.page, p.h_k, p.d, tb, hkt, is_bf16)) return -1;
        if (!make_kv_map(&mv, p.v, p.v_rs, p.v_hs, p.v_bs, p.page, p.h_k, p.d, tb, hkt, is_bf16)) return -1;
    }

    bool ok;
    if (is_bf16) ok = p.d == 128 ? launch_decode<__nv_bfloat16, 128>(mk, mv, a, stream)
                                 : launch_decode<__nv_bfloat16, 64>(mk, mv, a, stream);
    else ok = p.d == 128 ? launch_decode<__half, 128>(mk, mv, a, stream)
                         : launch_decode<__half, 64>(mk, mv, a, stream);
    return ok ? 1 : -1;
}

void set_decode_trace(void *device_buf) { g_decode_trace = device_buf; }
void set_decode_merge(int enable) { g_decode_no_merge = enable ? 0 : 1; }

}  // namespace atoma
