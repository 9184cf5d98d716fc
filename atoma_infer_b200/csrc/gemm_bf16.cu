// Dense projection GEMM  y[T,N] = x[T,K] . w[N,K]^T  (bf16 in, fp32 accumulate, bf16 out) for
// sm_100a: the q/k/v/o projections of the attention block (reference models/src/llama.rs:269-271,
// 311, which go to cuBLAS through candle `Linear`; there is no reference kernel to follow).
//
// Small-T ("weight streaming") kernel: at decode time T is the batch (<= 128 rows) and the GEMM is
// bound by reading W from HBM exactly once, so
//  * the roles of the operands are swapped: the WEIGHT tile is the 128-row M operand of
//    tcgen05.mma and the T (padded to a multiple of 16) activations are its N operand, so a
//    32-token batch costs a 128x32 accumulator, not a 128x128 one that is 3/4 padding;
//  * work = (128-row weight tile, 64-wide k block) units of 16 KiB of W; a persistent grid of
//    min(#SM, units) CTAs takes equal contiguous slices (stream-K), so 48 or 32 weight tiles still
//    load all 148 SMs evenly; a tile whose k range is shared by several CTAs is reduced from fp32
//    partials in a library workspace by the last CTA to arrive, in a fixed order (deterministic);
//  * warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, multi-stage mbarrier ring),
//    warp 1 = tcgen05.mma issuer (one thread; accumulators double-buffered in TMEM),
//    warps 2-5 = epilogue (tcgen05.ld -> registers -> global).
//
// Algorithmic bytes: 2 (N K + T K + T N); flops 2 T N K.
#include "common.cuh"
#include "skinny_args.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace atoma {
namespace {

constexpr int kTileN = 128;       // weight rows per tile (UMMA M)
constexpr int kBlockK = 64;       // k elements per stage: 128 bytes, one swizzle row
constexpr int kUmmaK = 16;        // k per tcgen05.mma (16-bit inputs)
constexpr int kGemmThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kMaxTPad = 256;     // activation rows per tile (UMMA N); larger T is cut into chunks
constexpr int kMaxStages = 12;
constexpr int kGemmMaxCtas = 256;
constexpr int kWBytes = kTileN * kBlockK * 2;  // 16 KiB

struct GemmArgs {
    void *y;
    int64_t ldy;
    int T, N, K;
    int t_pad;        // T rounded up to a multiple of 16
    int kb_per_tile;  // ceil(K / 64)
    int tiles;        // ceil(N / 128) * t_chunks; tile = n_tile * t_chunks + t_chunk
    int t_chunks;     // ceil(T / t_pad)
    int stages;
    int tmem_cols;    // power of two >= 2 * t_pad
    int is_bf16;
    int *counters;    // [tiles], zero between launches
    float *ws;        // [cta][2][t_pad][128] fp32 partials
    int prefetch_w;   // W is immutable: stream it before the grid dependency resolves
    unsigned long long *trace;  // diagnostics (atoma_debug_gemm_trace): [cta][4] or null
};

__device__ __forceinline__ unsigned long long g_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

__device__ __forceinline__ long long g_cta_start(long long units, int cta, int num_ctas) {
    return units * cta / num_ctas;
}
__device__ __forceinline__ int g_cta_of_unit(long long units, long long u, int num_ctas) {
    return static_cast<int>(((u + 1) * num_ctas + units - 1) / units - 1);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_stream_kernel(const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_x, const GemmArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int last_flag;
    __shared__ int part_slot[kGemmMaxCtas];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long t_start = a.trace ? g_timer_ns() : 0ull;
    const int x_bytes = a.t_pad * 128;
    const int stage_bytes = kWBytes + x_bytes;

    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, a.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;

    const long long units = static_cast<long long>(a.tiles) * a.kb_per_tile;
    const int ncta = static_cast<int>(min(static_cast<long long>(gridDim.x), units));
    const int cta = blockIdx.x;
    const bool active = cta < ncta;
    const long long u_begin = active ? g_cta_start(units, cta, ncta) : 0;
    const long long u_end = active ? g_cta_start(units, cta + 1, ncta) : 0;
    const int KB = a.kb_per_tile;

    // Producer and MMA warps run converged on warp-uniform values with only the asynchronous instructions
    // under elect.sync (see gemm_skinny.cu: inside `if (lane == 0)` every UTMALDG / UTCHMMA is wrapped in
    // an R2UR + ELECT waterfall loop, ~70 cycles per MMA — more than a 128x128x16 MMA occupies the pipe).
    if (warp == 0) {
        // ============================ TMA producer ============================
        pdl_launch_dependents();
        // W is read once when there is a single activation chunk (decode): evict-first.  With
        // several chunks every weight tile is re-read from L2 once per chunk: keep it.
        const uint64_t w_policy = a.t_chunks == 1 ? kL2EvictFirst : kL2EvictLast;
        // Programmatic dependent launch: W does not depend on the previous kernel, so the ring is
        // filled with weight tiles while the predecessor drains; x (its output) follows the wait.
        const int pre = a.prefetch_w ? static_cast<int>(min(static_cast<long long>(a.stages), u_end - u_begin)) : 0;
        int tile = static_cast<int>(u_begin / KB), kb = static_cast<int>(u_begin % KB);   // running (tile, k block)
        int n_tile = tile / a.t_chunks, t_chunk = tile % a.t_chunks;
        auto advance = [&]() {
            if (++kb == KB) {
                kb = 0;
                ++tile;
                if (++t_chunk == a.t_chunks) {
                    t_chunk = 0;
                    ++n_tile;
                }
            }
        };
        if (pre > 0) {
            const int tile0 = tile, kb0 = kb, n0 = n_tile, c0 = t_chunk;
            for (int i = 0; i < pre; ++i) {
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&full[i], stage_bytes);
                    tma_load_2d(tiles + i * stage_bytes, &map_w, &full[i], kb * kBlockK, n_tile * kTileN, w_policy);
                }
                __syncwarp();
                advance();
            }
            tile = tile0; kb = kb0; n_tile = n0; t_chunk = c0;
        }
        pdl_wait();
        if (a.trace && lane == 0) a.trace[blockIdx.x * 4 + 0] = t_start;
        int stage = 0;
        uint32_t phase = 0;
        long long u = u_begin;
        for (int i = 0; i < pre; ++i, ++u) {
            if (elect_one_sync())
                tma_load_2d(tiles + i * stage_bytes + kWBytes, &map_x, &full[i], kb * kBlockK, t_chunk * a.t_pad, kL2EvictLast);
            __syncwarp();
            advance();
        }
        stage = pre % a.stages;
        phase = (pre == a.stages) ? 1u : 0u;
        for (; u < u_end; ++u) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char *ws = tiles + stage * stage_bytes;
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full[stage], stage_bytes);
                // W is read exactly once: evict-first; x is re-read by every CTA: keep it in L2.
                tma_load_2d(ws, &map_w, &full[stage], kb * kBlockK, n_tile * kTileN, w_policy);
                tma_load_2d(ws + kWBytes, &map_x, &full[stage], kb * kBlockK, t_chunk * a.t_pad, kL2EvictLast);
            }
            __syncwarp();
            advance();
            if (++stage == a.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        pdl_wait();
        const uint32_t idesc = umma_idesc_f16(kTileN, a.t_pad, a.is_bf16 != 0);
        const uint32_t ring_u32 = smem_u32(tiles);
        int stage = 0;
        uint32_t phase = 0;
        int seg = 0;
        long long u = u_begin;
        int k0 = static_cast<int>(u_begin % KB);
        while (u < u_end) {
            const int n = static_cast<int>(min(static_cast<long long>(KB - k0), u_end - u));
            const int buf = seg & 1;
            mbar_wait(&tempty[buf], ((seg >> 1) & 1) ^ 1);  // epilogue drained this accumulator
            tc_fence_after_sync();
            const uint32_t tmem_d = tmem_base + buf * a.t_pad;
            for (int i = 0; i < n; ++i) {
                mbar_wait(&full[stage], phase);
                tc_fence_after_sync();
                if (a.trace && seg == 0 && i == 0 && lane == 0) a.trace[blockIdx.x * 4 + 1] = g_timer_ns();
                const uint32_t wa = ring_u32 + stage * stage_bytes;
                if (elect_one_sync()) {
                    const uint64_t da = umma_desc_k_sw128(wa), db = umma_desc_k_sw128(wa + kWBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 16 elements = 32 bytes along K inside the swizzle row: +2 (16-byte units)
                        umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);  // smem slot is free once these MMAs have read it
                }
                __syncwarp();
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one_sync()) umma_commit(&tfull[buf]);  // accumulator complete
            __syncwarp();
            ++seg;
            u += n;
            k0 = 0;
        }
    } else {
        // ============================ epilogue warps (2..5) ============================
        pdl_wait();
        const int quad = warp & 3;             // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;      // weight row within the tile
        const int etid = tid - 64;             // 0..127
        __nv_bfloat16 *yb = static_cast<__nv_bfloat16 *>(a.y);
        __half *yh = static_cast<__half *>(a.y);
        int seg = 0;
        long long u = u_begin;
        while (u < u_end) {
            const int tile = static_cast<int>(u / KB), k0 = static_cast<int>(u % KB);
            const int n = static_cast<int>(min(static_cast<long long>(KB - k0), u_end - u));
            const int buf = seg & 1;
            const bool whole = (n == KB);
            const int n_tile = tile / a.t_chunks, t_base = (tile % a.t_chunks) * a.t_pad;
            const int ncol = n_tile * kTileN + row;  // output column (weight row)
            const int slot = (u == u_begin) ? 0 : 1;
            float *part = a.ws + (static_cast<int64_t>(cta) * 2 + slot) * (a.t_pad * kTileN);

            mbar_wait(&tfull[buf], (seg >> 1) & 1);
            tc_fence_after_sync();
            if (a.trace && etid == 0) a.trace[blockIdx.x * 4 + 2] = g_timer_ns();  // last one wins
            for (int c = 0; c < a.t_pad / 16; ++c) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * a.t_pad + c * 16, v);
                tmem_ld_wait();
                if (whole) {
                    if (ncol < a.N) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int t = t_base + c * 16 + j;
                            if (t < a.T) {
                                const float f = __uint_as_float(v[j]);
                                if (a.is_bf16) yb[static_cast<int64_t>(t) * a.ldy + ncol] = __float2bfloat16_rn(f);
                                else yh[static_cast<int64_t>(t) * a.ldy + ncol] = __float2half_rn(f);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        __stcg(part + (c * 16 + j) * kTileN + row, __uint_as_float(v[j]));
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);

            if (!whole) {
                // ---- publish, and reduce if this CTA is the last contributor of the tile ----
                const long long ub = static_cast<long long>(tile) * KB, ue = ub + KB;
                const int fc = g_cta_of_unit(units, ub, ncta), lc = g_cta_of_unit(units, ue - 1, ncta);
                __threadfence();
                named_bar_sync(1, kEpiThreads);
                if (etid == 0) {
                    const int old = atomicAdd(&a.counters[tile], 1);
                    last_flag = (old == lc - fc) ? 1 : 0;
                }
                named_bar_sync(1, kEpiThreads);
                if (last_flag) {
                    __threadfence();
                    const int nparts = lc - fc + 1;
                    for (int i = etid; i < nparts; i += kEpiThreads) {
                        const long long cs = g_cta_start(units, fc + i, ncta);
                        part_slot[i] = (fc + i) * 2 + ((cs >= ub && cs < ue) ? 0 : 1);
                    }
                    named_bar_sync(1, kEpiThreads);
                    // Thread (tq, r4) reduces rows r4*4..r4*4+3 of the tile for t = tq, tq+4, ...: 16-byte
                    // loads, 4 t values x 4 partials in flight at a time, partials added in CTA order.
                    const int r4 = etid & 31, tq = etid >> 5;
                    const int col4 = n_tile * kTileN + r4 * 4;
                    const bool vec_ok = (a.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 7) == 0) &&
                                        (col4 + 3 < a.N);
                    for (int j0 = 0; j0 < a.t_pad / 4; j0 += 4) {
                        float4 acc[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int i0 = 0; i0 < nparts; i0 += 4) {
                            float4 w[4][4];
#pragma unroll
                            for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int t = tq + 4 * (j0 + jj);
                                    w[ii][jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (i0 + ii < nparts && t < a.t_pad)
                                        w[ii][jj] = __ldcg(reinterpret_cast<const float4 *>(
                                            a.ws + static_cast<int64_t>(part_slot[i0 + ii]) * (a.t_pad * kTileN) +
                                            t * kTileN + r4 * 4));
                                }
                            }
#pragma unroll
                            for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    acc[jj].x += w[ii][jj].x;
                                    acc[jj].y += w[ii][jj].y;
                                    acc[jj].z += w[ii][jj].z;
                                    acc[jj].w += w[ii][jj].w;
                                }
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int t = t_base + tq + 4 * (j0 + jj);
                            if (t >= a.T || tq + 4 * (j0 + jj) >= a.t_pad) continue;
                            const int64_t o = static_cast<int64_t>(t) * a.ldy + col4;
                            const float f[4] = {acc[jj].x, acc[jj].y, acc[jj].z, acc[jj].w};
                            if (vec_ok) {
                                uint2 pk;
                                if (a.is_bf16) {
                                    pk.x = pack2<__nv_bfloat16>(f[0], f[1]);
                                    pk.y = pack2<__nv_bfloat16>(f[2], f[3]);
                                } else {
                                    pk.x = pack2<__half>(f[0], f[1]);
                                    pk.y = pack2<__half>(f[2], f[3]);
                                }
                                *reinterpret_cast<uint2 *>(static_cast<char *>(a.y) + o * 2) = pk;
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    if (col4 + e < a.N) {
                                        if (a.is_bf16) yb[o + e] = __float2bfloat16_rn(f[e]);
                                        else yh[o + e] = __float2half_rn(f[e]);
                                    }
                                }
                            }
                        }
                    }
                    if (etid == 0) a.counters[tile] = 0;  // self-cleaning for the next launch
                }
                named_bar_sync(1, kEpiThreads);  // last_flag / part_slot / slot 1 reuse
            }
            ++seg;
            u += n;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (a.trace && tid == 0) a.trace[blockIdx.x * 4 + 3] = g_timer_ns();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}


// ---------------------------------------------------------------------------------------------
struct Map2Key {
    const void *ptr;
    int64_t rows, cols, ld;
    int box_rows, bf16;
    bool operator==(const Map2Key &o) const { return std::memcmp(this, &o, sizeof(Map2Key)) == 0; }
};
struct Map2Entry {
    Map2Key key;
    CUtensorMap map;
};

// 2-D K-major operand [rows, cols] with row stride ld (elements); box = [box_rows][64], 128B swizzle.
bool make_operand_map(CUtensorMap *out, const void *ptr, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows, bool is_bf16) {
    static thread_local std::vector<Map2Entry> cache;
    Map2Key key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.rows = rows; key.cols = cols; key.ld = ld; key.box_rows = box_rows; key.bf16 = is_bf16;
    for (const Map2Entry &e : cache)
        if (e.key == key) { *out = e.map; return true; }
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_status(kErrDriver, "gemm: cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
        return false;
    }
    if (cache.size() >= 512) cache.clear();
    cache.push_back(Map2Entry{key, *out});
    return true;
}

void *g_gemm_trace = nullptr;
int g_gemm_stage_limit = 0;  // diagnostics (atoma_debug_gemm_stages): 0 = as many as fit
int g_gemm_algo = 0;         // atoma_debug_gemm_algo: 0 = decode shapes on the skinny kernel (default), 2 = everything on stream-K

bool launch_stream_gemm(const void *x, const void *w, void *y, int T, int64_t N, int64_t K, int64_t ldx,
                        int64_t ldw, int64_t ldy, bool is_bf16, cudaStream_t stream) {
    const DeviceInfo &di = current_device_info();
    GemmArgs a;
    a.y = y; a.ldy = ldy; a.T = T; a.N = static_cast<int>(N); a.K = static_cast<int>(K);
    a.t_pad = T > kMaxTPad ? kMaxTPad : (T + 15) / 16 * 16;
    a.t_chunks = (T + a.t_pad - 1) / a.t_pad;
    a.kb_per_tile = static_cast<int>((K + kBlockK - 1) / kBlockK);
    a.tiles = static_cast<int>((N + kTileN - 1) / kTileN) * a.t_chunks;
    a.is_bf16 = is_bf16 ? 1 : 0;
    int cols = 32;
    while (cols < 2 * a.t_pad) cols <<= 1;
    a.tmem_cols = cols;
    const int stage_bytes = kWBytes + a.t_pad * 128;
    a.stages = std::min(kMaxStages, (200 * 1024) / stage_bytes);
    if (g_gemm_stage_limit > 0) a.stages = std::max(2, std::min(a.stages, g_gemm_stage_limit));
    const size_t smem = static_cast<size_t>(a.stages) * stage_bytes + 1024;
    const int64_t units = static_cast<int64_t>(a.tiles) * a.kb_per_tile;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(di.sm_count, kGemmMaxCtas), units)));

    // The arrival counters (zero between launches, self-resetting) live in their OWN workspace: in a
    // shared buffer the boundary between counters and partials would move with the shape, and a launch
    // with more tiles would find an earlier launch's fp32 partials where it expects zeroed counters.
    const size_t counters_bytes = std::max<size_t>(64 * 1024, static_cast<size_t>(a.tiles) * sizeof(int));
    const size_t ws_bytes = static_cast<size_t>(grid) * 2 * a.t_pad * kTileN * sizeof(float);
    char *cnt = static_cast<char *>(get_workspace(stream, counters_bytes, kWorkspaceGemmCounters));
    char *ws = static_cast<char *>(get_workspace(stream, ws_bytes, kWorkspaceGemm));
    if (!ws || !cnt) return false;
    a.counters = reinterpret_cast<int *>(cnt);
    a.ws = reinterpret_cast<float *>(ws);
    a.trace = static_cast<unsigned long long *>(g_gemm_trace);
    a.prefetch_w = ((launch_flags() & (kLaunchPdl | kLaunchConstWeights)) == (kLaunchPdl | kLaunchConstWeights)) ? 1 : 0;

    CUtensorMap mw, mx;
    if (!make_operand_map(&mw, w, N, K, ldw, kTileN, is_bf16)) return false;
    if (!make_operand_map(&mx, x, T, K, ldx, a.t_pad, is_bf16)) return false;

    static thread_local int attr_dev = -1;
    if (attr_dev != di.ordinal) {
        // stages * stage_bytes <= 200 KiB by construction (+ 1 KiB alignment slack)
        cudaError_t e = cudaFuncSetAttribute(gemm_tn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024 + 1024);
        if (e != cudaSuccess) {
            set_status(static_cast<int>(e), "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return false;
        }
        attr_dev = di.ordinal;
    }
    cudaError_t e = launch_ex(gemm_tn_stream_kernel, dim3(grid), dim3(kGemmThreads), smem, stream, mw, mx, a);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "gemm_tn_stream_kernel: %s", cudaGetErrorString(e));
        return false;
    }
    return check_launch("gemm_tn_stream_kernel");
}

}  // namespace

void set_gemm_trace(void *device_buf) { g_gemm_trace = device_buf; }
void set_gemm_stage_limit(int stages) { g_gemm_stage_limit = stages; }
void set_gemm_algo(int algo) { g_gemm_algo = algo; }

bool launch_gemm_bf16_tn(const void *x, const void *w, void *y, int64_t T, int64_t N, int64_t K,
                         int64_t ldx, int64_t ldw, int64_t ldy, cudaStream_t stream) {
    if (T <= 0 || N <= 0) return true;
    if (K <= 0) {
        set_status(kErrInvalidArg, "atoma_gemm_bf16_tn: K = %lld", static_cast<long long>(K));
        return false;
    }
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!x || !w || !y || !al16(x) || !al16(w) || ldx % 8 || ldw % 8 || ldx < K || ldw < K || ldy < N) {
        set_status(kErrUnsupported,
                   "atoma_gemm_bf16_tn: x and w must be 16-byte aligned with row strides that are multiples of 8 "
                   "elements (ldx=%lld ldw=%lld ldy=%lld)",
                   static_cast<long long>(ldx), static_cast<long long>(ldw), static_cast<long long>(ldy));
        return false;
    }
    const DeviceInfo &di = current_device_info();
    if (di.cc_major != 10) {
        set_status(kErrUnsupported, "atoma_gemm_bf16_tn needs tcgen05 (sm_100a), device is sm_%d%d", di.cc_major,
                   di.cc_minor);
        return false;
    }
    if (T > (1 << 24)) {
        set_status(kErrUnsupported, "atoma_gemm_bf16_tn: T = %lld", static_cast<long long>(T));
        return false;
    }
    // Decode shapes go to the weight-streaming kernel without split-K (gemm_skinny.cu); atoma_debug_gemm_algo(2)
    // keeps them on the stream-K kernel (diagnostics / A-B measurements).
    if (g_gemm_algo != 2 && skinny_gemm_supported(T, K, ldx, ldw) && N < (1ll << 31)) {
        SkinnyArgs a{};
        a.epi = kEpiPlain;
        a.y = y;
        a.ldy = ldy;
        return launch_skinny_gemm(a, x, w, static_cast<int>(T), K, ldx, ldw, 1, static_cast<int>(N), 1, 0, 0, true, stream);
    }
    if (!launch_stream_gemm(x, w, y, static_cast<int>(T), N, K, ldx, ldw, ldy, true, stream)) return false;
    return true;
}

}  // namespace atoma
