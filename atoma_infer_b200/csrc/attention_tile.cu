// General attention forward (dense / varlen / paged K-V, causal, ALiBi, GQA, any head dim that is
// a multiple of 8 up to 256) for sm_100a.  This is the path every run_mha call that is not a
// paged single-token decode takes: prefill (`flash_attn_varlen`, reference csrc/src/lib.rs:606-1103),
// dense (`flash_attn`, :31-343), paged-prefix prefill and multi-token kv-cache calls.
//
// What it computes is the reference's contract (flash_fwd_kernel.h:56-500 / :504-1092, mask.h,
// softmax.h): fp32 scores, online softmax in the exp2 domain, P rounded to the 16-bit type before
// P.V, fp32 output accumulation, one final 1/rowsum, rows with no visible key -> 0.  How it is
// computed is new: a 64 x BN tile per CTA, K then V streamed through one shared-memory buffer each
// with cp.async zero-fill (so ragged tails and head dims such as 8 or 40 never read out of bounds,
// unlike the reference's unpredicated Is_even_K loads), block-table lookups only for pages below
// ceil(len_k / page).
#include "common.cuh"
#include "mha_params.h"

namespace atoma {
namespace {

constexpr int kBM = 64;
constexpr int kThreads = 128;

struct SeqInfo {
    int q_start, len_q, len_k;
    int64_t q_off, o_off, k_off, v_off;
};

__device__ __forceinline__ SeqInfo seq_info(const MhaParams &p, int bidb) {
    SeqInfo s;
    // reference block_info.h:14-24
    s.q_start = p.cu_q ? p.cu_q[bidb] : -1;
    s.len_q = p.cu_q ? p.cu_q[bidb + 1] - s.q_start : p.seqlen_q;
    const int k_start = (p.cu_k && p.k_cumulative) ? p.cu_k[bidb] : -1;
    const int len_k_cache =
        p.cu_k ? (p.k_cumulative ? p.cu_k[bidb + 1] - k_start : p.cu_k[bidb]) : p.seqlen_k;
    s.len_k = p.seqused_k ? p.seqused_k[bidb] : len_k_cache;
    s.q_off = s.q_start < 0 ? bidb * p.q_bs : static_cast<int64_t>(s.q_start) * p.q_rs;
    s.o_off = s.q_start < 0 ? bidb * p.o_bs : static_cast<int64_t>(s.q_start) * p.o_rs;
    s.k_off = k_start < 0 ? bidb * p.k_bs : static_cast<int64_t>(k_start) * p.k_rs;
    s.v_off = k_start < 0 ? bidb * p.v_bs : static_cast<int64_t>(k_start) * p.v_rs;
    return s;
}

template <typename T, int DP, int BN>
__global__ void __launch_bounds__(kThreads) attn_tile_kernel(const MhaParams p) {
    constexpr int LDS = DP + 8;  // padded row: ldmatrix rows land in distinct bank groups
    constexpr int CPR = DP / 8;  // 16-byte chunks per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sQ = reinterpret_cast<T *>(smem_raw);
    T *sK = sQ + kBM * LDS;
    T *sV = sK + BN * LDS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int m_block = blockIdx.x, head = blockIdx.y, bidb = blockIdx.z;
    const SeqInfo si = seq_info(p, bidb);
    if (m_block * kBM >= si.len_q) return;
    const int kvh = head / (p.h / p.h_k);  // flash_fwd_kernel.h:593
    const int len_q = si.len_q, len_k = si.len_k;
    const int causal_shift = len_k - len_q;  // mask.h:171-172 (bottom-right aligned)

    const T *qg = static_cast<const T *>(p.q) + si.q_off + head * p.q_hs;
    T *og = static_cast<T *>(p.o) + si.o_off + head * p.o_hs;
    const T *kg = static_cast<const T *>(p.k) + kvh * p.k_hs;
    const T *vg = static_cast<const T *>(p.v) + kvh * p.v_hs;
    const int *bt = p.block_table ? p.block_table + bidb * p.bt_bs : nullptr;

    int n_max = (len_k + BN - 1) / BN;
    if (p.is_causal) {
        const int lim = (m_block + 1) * kBM + causal_shift;
        n_max = min(n_max, lim <= 0 ? 0 : (lim + BN - 1) / BN);
    }

    float lse_total_q = 0;  // only to keep the compiler quiet about unused paths
    (void)lse_total_q;

    auto write_lse = [&](int row, float val) {
        if (!p.lse || row >= len_q) return;
        int64_t idx;
        if (p.unpadded_lse && p.cu_q) {
            const int total_q = p.cu_q[p.b];
            idx = static_cast<int64_t>(head) * total_q + si.q_start + row;
        } else {
            idx = (static_cast<int64_t>(bidb) * p.h + head) * p.seqlen_q + row;
        }
        p.lse[idx] = val;
    };

    if (n_max <= 0) {
        // No visible key for any row of this tile: O = 0, LSE = +inf (flash_fwd_kernel.h:543-582).
        for (int idx = tid; idx < kBM * CPR; idx += kThreads) {
            const int r = idx / CPR, c = idx % CPR, row = m_block * kBM + r;
            if (row < len_q && c * 8 < p.d)
                *reinterpret_cast<uint4 *>(og + static_cast<int64_t>(row) * p.o_rs + c * 8) =
                    make_uint4(0, 0, 0, 0);
        }
        if (tid < kBM) write_lse(m_block * kBM + tid, INFINITY);
        return;
    }

    // ---- tile loaders (cp.async, zero fill for rows >= len and columns >= d) ----
    auto load_q = [&]() {
        for (int idx = tid; idx < kBM * CPR; idx += kThreads) {
            const int r = idx / CPR, c = idx % CPR, row = m_block * kBM + r;
            const bool ok = row < len_q && c * 8 < p.d;
            const T *src = ok ? qg + static_cast<int64_t>(row) * p.q_rs + c * 8 : qg;
            cp_async_16(smem_u32(sQ + r * LDS + c * 8), src, ok ? 16 : 0);
        }
    };
    auto load_kv = [&](T *s, const T *gbase, int64_t seq_off, int64_t bs, int64_t rs, int n_block) {
        for (int idx = tid; idx < BN * CPR; idx += kThreads) {
            const int r = idx / CPR, c = idx % CPR, t = n_block * BN + r;
            const bool ok = t < len_k && c * 8 < p.d;
            const T *src = gbase;
            if (ok) {
                if (bt) {
                    // reference utils.h:296-314; entries past ceil(len_k/page) are never read
                    const int pg = t / p.page;
                    src = gbase + static_cast<int64_t>(bt[pg]) * bs +
                          static_cast<int64_t>(t - pg * p.page) * rs + c * 8;
                } else {
                    src = gbase + seq_off + static_cast<int64_t>(t) * rs + c * 8;
                }
            }
            cp_async_16(smem_u32(s + r * LDS + c * 8), src, ok ? 16 : 0);
        }
    };

    load_q();
    load_kv(sK, kg, si.k_off, p.k_bs, p.k_rs, 0);
    cp_async_commit();

    float o_acc[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    const int row0 = m_block * kBM + warp * 16 + g;
    float slope = 0.f;
    if (p.alibi) slope = p.alibi[bidb * p.alibi_bs + head] * 1.4426950408889634f;  // log2 domain

    for (int n = 0; n < n_max; ++n) {
        cp_async_wait_all();
        __syncthreads();  // K(n) (+Q) visible; every warp is done with V(n-1)
        load_kv(sV, vg, si.v_off, p.v_bs, p.v_rs, n);
        cp_async_commit();

        // ---- S = Q K^T ----
        float s[BN / 8][4];
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk) {
            uint32_t a[4];
            ldmatrix_x4(a, smem_u32(sQ + (warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS +
                                    kk * 16 + (lane >> 4) * 8));
#pragma unroll
            for (int j2 = 0; j2 < BN / 16; ++j2) {
                uint32_t bf[4];
                ldmatrix_x4(bf, smem_u32(sK + (j2 * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS +
                                         kk * 16 + ((lane >> 3) & 1) * 8));
                Mma16816<T>::run(s[2 * j2], a, bf[0], bf[1]);
                Mma16816<T>::run(s[2 * j2 + 1], a, bf[2], bf[3]);
            }
        }

        cp_async_wait_all();
        __syncthreads();  // V(n) visible; every warp is done with K(n)
        if (n + 1 < n_max) {
            load_kv(sK, kg, si.k_off, p.k_bs, p.k_rs, n + 1);
            cp_async_commit();
        }

        // ---- scale, bias, mask (log2 domain) ----
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = n * BN + j * 8 + 2 * t4 + (e & 1);
                const int row = row0 + (e >> 1) * 8;
                float x = s[j][e] * p.scale_log2;
                if (p.alibi) {
                    // mask.h:155-157, 179-186 (slope is applied after the softmax scale)
                    x += p.is_causal ? slope * col : -slope * fabsf(float(row + causal_shift - col));
                }
                const bool masked = col >= len_k || (p.is_causal && col > row + causal_shift);
                x = masked ? -INFINITY : x;
                s[j][e] = x;
                mx[e >> 1] = fmaxf(mx[e >> 1], x);
            }
        }
        float corr[2], m_safe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            m_safe[r] = m_new == -INFINITY ? 0.f : m_new;
            corr[r] = fast_exp2(m_run[r] - m_safe[r]);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pv = fast_exp2(s[j][e] - m_safe[e >> 1]);
                s[j][e] = pv;
                rs[e >> 1] += pv;
            }
        }
        l_run[0] = l_run[0] * corr[0] + rs[0];
        l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
        for (int i = 0; i < DP / 8; ++i) {
            o_acc[i][0] *= corr[0];
            o_acc[i][1] *= corr[0];
            o_acc[i][2] *= corr[1];
            o_acc[i][3] *= corr[1];
        }

        // ---- O += P V (P rounded to the 16-bit type first, flash_fwd_kernel.h:949-954) ----
#pragma unroll
        for (int kk2 = 0; kk2 < BN / 16; ++kk2) {
            uint32_t a[4];
            a[0] = pack2<T>(s[2 * kk2][0], s[2 * kk2][1]);
            a[1] = pack2<T>(s[2 * kk2][2], s[2 * kk2][3]);
            a[2] = pack2<T>(s[2 * kk2 + 1][0], s[2 * kk2 + 1][1]);
            a[3] = pack2<T>(s[2 * kk2 + 1][2], s[2 * kk2 + 1][3]);
#pragma unroll
            for (int dj2 = 0; dj2 < DP / 16; ++dj2) {
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, smem_u32(sV + (kk2 * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS +
                                               dj2 * 16 + (lane >> 4) * 8));
                Mma16816<T>::run(o_acc[2 * dj2], a, vf[0], vf[1]);
                Mma16816<T>::run(o_acc[2 * dj2 + 1], a, vf[2], vf[3]);
            }
        }
    }

    // ---- epilogue: 1/rowsum, stage through this warp's Q rows, 16-byte stores ----
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        l_run[r] = l;
        inv[r] = (l == 0.f || l != l) ? 1.f : 1.f / l;
    }
    __syncwarp();
#pragma unroll
    for (int dj = 0; dj < DP / 8; ++dj) {
        *reinterpret_cast<uint32_t *>(sQ + (warp * 16 + g) * LDS + dj * 8 + 2 * t4) =
            pack2<T>(o_acc[dj][0] * inv[0], o_acc[dj][1] * inv[0]);
        *reinterpret_cast<uint32_t *>(sQ + (warp * 16 + g + 8) * LDS + dj * 8 + 2 * t4) =
            pack2<T>(o_acc[dj][2] * inv[1], o_acc[dj][3] * inv[1]);
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * CPR; idx += 32) {
        const int r = idx / CPR, c = idx % CPR, row = m_block * kBM + warp * 16 + r;
        if (row < len_q && c * 8 < p.d)
            *reinterpret_cast<uint4 *>(og + static_cast<int64_t>(row) * p.o_rs + c * 8) =
                *reinterpret_cast<const uint4 *>(sQ + (warp * 16 + r) * LDS + c * 8);
    }
    if (t4 == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float l = l_run[r];
            const float lse = (l == 0.f || l != l)
                                  ? INFINITY
                                  : m_run[r] * 0.6931471805599453f + __logf(l);
            write_lse(row0 + r * 8, lse);
        }
    }
}

template <typename T, int DP, int BN>
bool launch_tile(const MhaParams &p, cudaStream_t stream) {
    constexpr size_t smem = static_cast<size_t>(kBM + 2 * BN) * (DP + 8) * sizeof(T);
    static thread_local int attr_dev = -1;
    const DeviceInfo &di = current_device_info();
    if (attr_dev != di.ordinal) {
        cudaFuncSetAttribute(attn_tile_kernel<T, DP, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
        attr_dev = di.ordinal;
    }
    dim3 grid((p.seqlen_q + kBM - 1) / kBM, p.h, p.b);
    attn_tile_kernel<T, DP, BN><<<grid, kThreads, smem, stream>>>(p);
    note_launch();
    return check_launch("attn_tile_kernel");
}

template <typename T>
bool launch_tile_t(const MhaParams &p, cudaStream_t stream) {
    if (p.d <= 32) return launch_tile<T, 32, 64>(p, stream);
    if (p.d <= 64) return launch_tile<T, 64, 64>(p, stream);
    if (p.d <= 128) return launch_tile<T, 128, 64>(p, stream);
    return launch_tile<T, 256, 32>(p, stream);
}

}  // namespace

bool launch_attention_tile(const MhaParams &p, bool is_bf16, cudaStream_t stream) {
    if (p.b <= 0 || p.h <= 0 || p.seqlen_q <= 0) return true;
    if (p.h > 65535 || p.b > 65535) {
        set_status(kErrUnsupported, "run_mha: b=%d / h=%d exceed the grid limits (65535)", p.b, p.h);
        return false;
    }
    return is_bf16 ? launch_tile_t<__nv_bfloat16>(p, stream) : launch_tile_t<__half>(p, stream);
}

}  // namespace atoma
