// Argument block of the decode-time ("skinny") GEMM (gemm_skinny.cu) and the entry points the rest of
// the library uses to launch it (flash_api.cu: C ABI; llama_step.cu: the decode-step executor).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "norm_row.cuh"

namespace atoma {

enum : int {
    kEpiPlain = 0,  // y[t, n] = acc
    kEpiRope = 1,   // qkv projection: RoPE on q/k heads, store to y, scatter k/v to the paged cache (f1)
    kEpiSilu = 2,   // gate_up projection: y[t, j] = silu(gate_j) * up_j (f3)
    kEpiPush = 3,   // row-parallel projection under tensor parallelism: push the addend to every rank (f4)
};

struct SkinnyArgs {
    // ---- tiling (filled by launch_skinny_gemm) ----
    int T, t_box, m;               // token rows, TMA box rows (multiple of 8), UMMA M (64 | 128)
    int kb_total, kbs, chunks;     // K / 64, k blocks per stage, stages per tile
    int w_ops, wkb;                // weight ops per stage (1 | 2) and k blocks per weight op
    int stages;
    int np, halves, n_tile;        // weight rows per half per tile, halves (1 | 2), np * halves
    int J, G;                      // rows per half per group, groups
    int tiles_per_group, tiles;
    int w_bytes, x_bytes, stage_bytes;
    int acc_stride, tmem_cols, nacc;   // TMEM columns between accumulators, allocation, independent accumulators per tile
    int nbuf;                          // accumulator buffers (2: a tile's epilogue overlaps the next tile's MMAs; 1)
    int prefetch_w, is_bf16, epi;
    int debug_skip;                // diagnostics: 1 = skip x loads, 2 = skip W loads (garbage results)
    int trigger_late;              // 0: release the dependent grid at kernel start; N: N stages before this CTA's last weight load
    unsigned long long *trace;     // diagnostics: [cta][4] globaltimer ns, or null
    // ---- what the next kernel of the stream streams first (L2 prefetch by the W producer warp) ----
    const unsigned char *pf_ptr;
    long long pf_bytes;
    int pf_spread;                 // 1: a few pieces per stage from the start (this kernel leaves HBM headroom); 0: all after the last load
    // ---- output (plain / rope / silu) ----
    void *y;
    int64_t ldy;
    // ---- rope + cache write ----
    const int64_t *positions;      // [T]
    const void *cos, *sin;         // [max_pos, D/2], row stride rot_stride
    int64_t rot_stride;
    const int64_t *slot_mapping;   // [T] or null
    void *k_cache, *v_cache;       // [NB, block_size, Hkv, D]
    int64_t block_stride, block_size;
    int num_heads, num_kv_heads;
    // ---- tensor-parallel push ----
    unsigned char *push_region[8]; // every rank's peer-mapped region (own rank: local pointer)
    const uint32_t *push_state;    // local region header: [0] = epoch of the last completed all-reduce
    int64_t push_recv_offset;      // byte offset of recv[0][0] inside a region
    int64_t push_slot_stride;      // bytes between recv[0][..] and recv[1][..]
    int64_t push_recv_stride;      // bytes between recv[s][r] and recv[s][r + 1]
    int64_t push_plane_bytes;      // bytes between the two 16-byte planes of a vector
    int push_rank, push_world;
    // ---- fused add + RMSNorm tail (plain / push epilogues of the row-parallel projections o_proj, down_proj):
    // the min(T, grid) CTAs that finish LAST wait until every CTA's tile is stored (or pushed) and then run the
    // residual add + RMSNorm of the token rows (norm_row.cuh) — the separate norm launch disappears ----
    int norm_tail;
    uint32_t *tail_sync;           // [2], zero between launches: CTAs finished, norm CTAs finished
    NormArgs norm;                 // norm.x = this GEMM's y (plain) or null (push: the receive buffers)
};

bool skinny_gemm_supported(int64_t T, int64_t K, int64_t ldx, int64_t ldw);
bool launch_skinny_gemm(SkinnyArgs a, const void *x, const void *w, int T, int64_t K, int64_t ldx, int64_t ldw, int G,
                        int J, int halves, int64_t half_stride_rows, int64_t group_stride_rows, bool is_bf16,
                        cudaStream_t stream);
void set_skinny_trace(void *device_buf);
void set_skinny_tuning(int np, int kbs, int stages);
void set_skinny_debug(int skip, int w_ops);
void set_skinny_budget(int kib);
void set_skinny_chain(int on);
// gate_up + SwiGLU and down_proj chained in ONE kernel (gemm_skinny.cu: gemm_chain2_kernel).  `a1`: the epilogue fields of
// the second projection (epi = kEpiPlain with y / ldy, or kEpiPush with the push_* fields).  0 = shapes do not qualify
// (launch the two kernels), 1 = launched, -1 = error.
int try_launch_mlp_chain(const void *x, const void *w_gate_up, void *act, const void *w_down, SkinnyArgs a1, int T, int64_t inter,
                         int64_t hidden, int64_t ldx, int64_t ldw_gu, int64_t ld_act, int64_t ldw_down, bool is_bf16,
                         cudaStream_t stream);

}  // namespace atoma
