// Host-side operator layer (include/atoma_ops.h): a C++ restatement of the reference's Rust
// wrappers — csrc/src/lib.rs (FlashAttention :31-343, FlashAttentionVarLen :606-1103,
// FlashAttentionKvCache :1523-1855, utils :2107-2199), csrc/src/cache_manager.rs and the operator
// models/src/flash_attention.rs:322-469.  It validates like the reference (same error texts where
// the reference's tests assert them), derives the same run_mha argument tuple, and calls the
// kernel ABI.  No device memory is allocated here.
#include "atoma_ops.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local char t_msg[1024];

const char *fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_msg, sizeof(t_msg), fmt, ap);
    va_end(ap);
    return t_msg;
}

std::string dims_str(const int64_t *d, int n) {  // Rust `{:?}` of a slice / Shape
    std::string s = "[";
    for (int i = 0; i < n; ++i) {
        if (i) s += ", ";
        s += std::to_string(d[i]);
    }
    return s + "]";
}
std::string shape_str(const atoma_tensor *t) { return dims_str(t->shape, t->rank); }
std::string stride_str(const atoma_tensor *t) { return dims_str(t->stride, t->rank); }

const char *dtype_name(int dt) {
    switch (dt) {
        case ATOMA_F16: return "F16";
        case ATOMA_BF16: return "BF16";
        case ATOMA_F32: return "F32";
        case ATOMA_U32: return "U32";
        case ATOMA_I64: return "I64";
        case ATOMA_U8: return "U8";
        case ATOMA_F64: return "F64";
    }
    return "?";
}

int64_t round_multiple(int64_t x, int64_t m) { return (x + m - 1) / m * m; }  // lib.rs:2112-2114

int device_sm_count(int num_sms) {
    if (num_sms > 0) return num_sms;
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// lib.rs:2122-2167
int64_t num_splits_heuristic(int64_t batch_nheads_mblocks, int64_t num_sms, int64_t num_n_blocks,
                             int64_t max_splits) {
    if (static_cast<float>(batch_nheads_mblocks) >= 0.8f * static_cast<float>(num_sms)) return 1;
    max_splits = std::min(std::min(max_splits, num_sms), num_n_blocks);
    float max_eff = 0.f;
    std::vector<float> eff;
    auto ceil_div = [](int64_t a, int64_t b) { return (a + b - 1) / b; };
    auto eligible = [&](int64_t ns) {
        return ns == 1 || ceil_div(num_n_blocks, ns) != ceil_div(num_n_blocks, ns - 1);
    };
    for (int64_t ns = 1; ns <= max_splits; ++ns) {
        if (!eligible(ns)) {
            eff.push_back(0.f);
        } else {
            const float n_waves = static_cast<float>(batch_nheads_mblocks * ns) / static_cast<float>(num_sms);
            const float e = n_waves / std::ceil(n_waves);
            max_eff = std::max(max_eff, e);
            eff.push_back(e);
        }
    }
    for (int64_t ns = 1; ns <= max_splits; ++ns) {
        if (!eligible(ns)) continue;
        if (eff[ns - 1] >= 0.85f * max_eff) return ns;
    }
    return 1;
}

bool is_half(int dt) { return dt == ATOMA_F16 || dt == ATOMA_BF16; }

struct Alibi {
    void *ptr = nullptr;
    int64_t batch_stride = 0;
};

// Shared by the three wrappers (lib.rs:149-186, 865-904, 1679-1718).
const char *check_alibi(const atoma_tensor *alibi, int64_t num_heads, Alibi *out) {
    if (!alibi) return nullptr;
    if (alibi->dtype != ATOMA_F32)
        return fail("DType mismatch alibi_slopes %s, expected F32", dtype_name(alibi->dtype));
    out->batch_stride = alibi->rank == 2 ? alibi->stride[0] : 0;
    // `alibi_slopes_layout.shape().dims1()?` — a rank-2 tensor fails here in the reference too.
    if (alibi->rank != 1)
        return fail("unexpected rank, expected: 1, got: %d (%s)", alibi->rank, shape_str(alibi).c_str());
    if (num_heads != alibi->shape[0])
        return fail("shape mismatch alibi_slopes %s, expected %lld", shape_str(alibi).c_str(),
                    static_cast<long long>(num_heads));
    if (alibi->device < 0) return fail("alibi_slopes must be a cuda tensor");
    out->ptr = alibi->data;
    return nullptr;
}

void clear_plan(atoma_mha_plan *plan) { std::memset(plan, 0, sizeof(*plan)); }

}  // namespace

extern "C" uint32_t atoma_compute_num_splits(int64_t batch_size, int64_t num_heads,
                                             int64_t head_size, int64_t max_seqlen_k,
                                             int64_t max_seqlen_q, int num_sms) {
    // lib.rs:2169-2199
    const int64_t block_n = head_size <= 64 ? 256 : (head_size <= 128 ? 128 : 64);
    const int64_t num_n_blocks = (max_seqlen_k + block_n - 1) / block_n;
    const int64_t num_m_blocks = (max_seqlen_q + 64 - 1) / 64;
    const int64_t sms = device_sm_count(num_sms);
    return static_cast<uint32_t>(
        num_splits_heuristic(batch_size * num_heads * num_m_blocks, sms * 2, num_n_blocks, 128));
}

// ------------------------------------------------------------------------------------------------
// flash_attn (dense)                                                    csrc/src/lib.rs:31-343
// ------------------------------------------------------------------------------------------------
extern "C" const char *atoma_flash_attn_plan(const atoma_tensor *q, const atoma_tensor *k,
                                             const atoma_tensor *v, const atoma_tensor *alibi_slopes,
                                             float softmax_scale_in, int64_t wl_in, int64_t wr_in,
                                             float softcap_in, int num_sms, atoma_mha_plan *plan) {
    clear_plan(plan);
    if (!is_half(q->dtype)) return fail("flash-attn is only supported for f16/bf16 (%s)", dtype_name(q->dtype));
    if (q->dtype != k->dtype) return fail("query and key must have the same dtype");
    if (q->dtype != v->dtype) return fail("query and value must have the same dtype");
    if (q->rank != 4 || k->rank != 4 || v->rank != 4)
        return fail("flash-attn expects input tensors of rank 4 (q: %d, k: %d, v: %d)", q->rank, k->rank, v->rank);
    const int64_t b_sz = q->shape[0], seqlen_q_og = q->shape[1], num_heads_og = q->shape[2],
                  head_size_og = q->shape[3];
    const int64_t seqlen_k = k->shape[1], num_heads_k = k->shape[2];

    const bool swapped = seqlen_q_og == 1 && num_heads_og > num_heads_k && wl_in < 0 && wr_in < 0 &&
                         head_size_og % 8 == 0 && !alibi_slopes;
    int64_t q_stride[4], o_stride[4], out_shape[4], seqlen_q = seqlen_q_og, num_heads = num_heads_og;
    if (swapped) {  // lib.rs:77-90: q re-read as a contiguous (b, ngroups, hk, d)
        const int64_t ngroups = num_heads_og / num_heads_k;
        out_shape[0] = b_sz; out_shape[1] = ngroups; out_shape[2] = num_heads_k; out_shape[3] = head_size_og;
        q_stride[3] = 1; q_stride[2] = head_size_og; q_stride[1] = num_heads_k * head_size_og;
        q_stride[0] = ngroups * num_heads_k * head_size_og;
        seqlen_q = ngroups;
        num_heads = num_heads_k;
    } else {
        for (int i = 0; i < 4; ++i) { out_shape[i] = q->shape[i]; q_stride[i] = q->stride[i]; }
    }
    o_stride[3] = 1; o_stride[2] = out_shape[3]; o_stride[1] = out_shape[2] * out_shape[3];
    o_stride[0] = out_shape[1] * o_stride[1];

    if (q_stride[3] != 1) return fail("the last dim of q must be contiguous %s", dims_str(q_stride, 4).c_str());
    if (k->stride[3] != 1) return fail("the last dim of k must be contiguous %s", stride_str(k).c_str());
    if (v->stride[3] != 1) return fail("the last dim of v must be contiguous %s", stride_str(v).c_str());
    const int64_t expected_kv[4] = {b_sz, seqlen_k, num_heads_k, head_size_og};
    if (std::memcmp(expected_kv, k->shape, sizeof(expected_kv)) != 0)
        return fail("shape mismatch q %s and k %s", shape_str(q).c_str(), shape_str(k).c_str());
    if (std::memcmp(expected_kv, v->shape, sizeof(expected_kv)) != 0)
        return fail("shape mismatch q %s and v %s", shape_str(q).c_str(), shape_str(v).c_str());
    if (head_size_og > 256)
        return fail("only supports head dimension at most 256 (got %lld)", static_cast<long long>(head_size_og));
    if (head_size_og % 8 != 0)
        return fail("only supports head sizes that are a multiple of 8 (got %lld)", static_cast<long long>(head_size_og));
    if (num_heads % num_heads_k != 0)
        return fail("number of k/v heads %lld must divide number of heads in query %lld",
                    static_cast<long long>(num_heads_k), static_cast<long long>(num_heads));
    Alibi al;
    if (const char *e = check_alibi(alibi_slopes, num_heads, &al)) return e;

    // lib.rs:188-229
    int wl = (wl_in >= 0 && wl_in <= seqlen_k) ? static_cast<int>(wl_in) : -1;
    int wr = (wr_in >= 0 && wr_in <= seqlen_k) ? static_cast<int>(wr_in) : -1;
    int is_causal = (wl < 0 && wr == 0) ? 1 : 0;
    if (seqlen_q == 1 && !alibi_slopes) is_causal = 0;
    const int64_t head_size = round_multiple(head_size_og, 8);
    const int64_t head_size_rounded = round_multiple(head_size, 32);
    if (wl < 0 && wr >= 0) wl = static_cast<int>(seqlen_k);
    if (wl >= 0 && wr < 0) wr = static_cast<int>(seqlen_k);
    const uint32_t num_splits =
        atoma_compute_num_splits(b_sz, num_heads, head_size, seqlen_k, seqlen_q, num_sms);
    float softcap = softcap_in, scale = softmax_scale_in;
    if (softcap > 0.f) { softcap = softmax_scale_in / softcap; scale = softcap; } else softcap = 0.f;

    atoma_mha_args &a = plan->args;
    a.q_ptr = q->data; a.k_ptr = k->data; a.v_ptr = v->data;
    a.alibi_slopes_ptr = al.ptr;
    a.is_seqlens_k_cumulative = true;
    a.q_batch_stride = static_cast<uint32_t>(swapped ? q_stride[0] * seqlen_q : q_stride[0]);
    a.o_batch_stride = static_cast<uint32_t>(swapped ? o_stride[0] * seqlen_q : o_stride[0]);
    a.k_batch_stride = static_cast<uint32_t>(k->stride[0]);
    a.v_batch_stride = static_cast<uint32_t>(v->stride[0]);
    a.alibi_slopes_batch_stride = static_cast<uint32_t>(al.batch_stride);
    a.q_row_stride = static_cast<uint32_t>(q_stride[1]); a.k_row_stride = static_cast<uint32_t>(k->stride[1]);
    a.v_row_stride = static_cast<uint32_t>(v->stride[1]); a.o_row_stride = static_cast<uint32_t>(o_stride[1]);
    a.q_head_stride = static_cast<uint32_t>(q_stride[2]); a.k_head_stride = static_cast<uint32_t>(k->stride[2]);
    a.v_head_stride = static_cast<uint32_t>(v->stride[2]); a.o_head_stride = static_cast<uint32_t>(o_stride[2]);
    a.num_splits = num_splits;
    a.b = static_cast<uint32_t>(b_sz); a.h = static_cast<uint32_t>(num_heads);
    a.h_k = static_cast<uint32_t>(num_heads_k); a.d = static_cast<uint32_t>(head_size);
    a.d_rounded = static_cast<uint32_t>(head_size_rounded);
    a.softmax_scale = scale; a.scale_softmax_log2 = scale * 1.4426950408889634f;
    a.block_table = nullptr; a.block_table_batch_stride = 0; a.page_block_size = 0;
    a.seqused_k = nullptr;
    a.seqlen_q = static_cast<uint32_t>(seqlen_q); a.seqlen_k = static_cast<uint32_t>(seqlen_k);
    a.seqlen_q_rounded = static_cast<uint32_t>(round_multiple(seqlen_q, 128));
    a.seqlen_k_rounded = static_cast<uint32_t>(round_multiple(seqlen_k, 128));
    a.is_bf16 = q->dtype == ATOMA_BF16; a.is_causal = is_causal;
    a.window_size_left = wl; a.window_size_right = wr; a.softcap = softcap;
    a.unpadded_lse = true; a.force_split_kernel = false;

    plan->out_rank = 4;
    if (swapped) {  // lib.rs:334-336
        plan->out_shape[0] = b_sz; plan->out_shape[1] = 1; plan->out_shape[2] = num_heads_k * seqlen_q;
        plan->out_shape[3] = head_size_og;
    } else {
        std::memcpy(plan->out_shape, out_shape, sizeof(out_shape));
    }
    plan->lse_elems = b_sz * 128 * num_heads * seqlen_q;  // lib.rs:218-220 (sic)
    if (num_splits > 1) {
        plan->lseaccum_elems = static_cast<int64_t>(num_splits) * b_sz * num_heads * seqlen_q;
        plan->oaccum_elems = plan->lseaccum_elems * head_size_rounded;
    }
    plan->seqlenq_ngroups_swapped = swapped;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// flash_attn_varlen (+ block table)                                     csrc/src/lib.rs:606-1103
// ------------------------------------------------------------------------------------------------
extern "C" const char *atoma_flash_attn_varlen_plan(
    const atoma_tensor *q, const atoma_tensor *k, const atoma_tensor *v,
    const atoma_tensor *alibi_slopes, const atoma_tensor *seqlens_q, const atoma_tensor *seqlens_k,
    const atoma_tensor *seqused_k, const atoma_tensor *block_table, int64_t max_seqlen_q_in,
    int64_t max_seqlen_k, float softmax_scale_in, int64_t wl_in, int64_t wr_in, float softcap_in,
    int num_sms, atoma_mha_plan *plan) {
    clear_plan(plan);
    if (!is_half(q->dtype)) return fail("flash-attn is only supported for f16/bf16 (%s)", dtype_name(q->dtype));
    if (q->dtype != k->dtype) return fail("query and key must have the same dtype");
    if (q->dtype != v->dtype) return fail("query and value must have the same dtype");
    if (seqlens_q->device < 0) return fail("seqlens_q must be a cuda tensor");
    if (seqlens_q->dtype != ATOMA_U32) return fail("unexpected dtype, expected: U32, got: %s", dtype_name(seqlens_q->dtype));
    if (seqlens_q->rank != 1 || seqlens_q->stride[0] != 1) return fail("seqlens_q has to be contiguous");
    if (seqlens_k->device < 0) return fail("seqlens_k must be a cuda tensor");
    if (seqlens_k->dtype != ATOMA_U32) return fail("unexpected dtype, expected: U32, got: %s", dtype_name(seqlens_k->dtype));
    if (seqlens_k->rank != 1 || seqlens_k->stride[0] != 1) return fail("seqlens_k has to be contiguous");

    const int64_t nseqlens_q = seqlens_q->shape[0];
    const int64_t batch_size = nseqlens_q - 1;
    if (q->rank != 3) return fail("unexpected rank, expected: 3, got: %d (%s)", q->rank, shape_str(q).c_str());
    const int64_t total_q = q->shape[0], num_heads_og = q->shape[1], head_size_og = q->shape[2];

    const bool paged = block_table != nullptr;
    if (paged) {
        if (block_table->device < 0) return fail("block_table must be a cuda tensor");
        if (block_table->dtype != ATOMA_U32)
            return fail("unexpected dtype, expected: U32, got: %s", dtype_name(block_table->dtype));
        if (block_table->stride[block_table->rank - 1] != 1) return fail("block_table must be contiguous");
    }
    int64_t num_blocks = 0, total_k, num_heads_k;
    if (paged) {
        if (k->rank != 4) return fail("unexpected rank, expected: 4, got: %d (%s)", k->rank, shape_str(k).c_str());
        num_blocks = k->shape[0]; total_k = k->shape[1]; num_heads_k = k->shape[2];
    } else {
        if (k->rank != 3) return fail("unexpected rank, expected: 3, got: %d (%s)", k->rank, shape_str(k).c_str());
        total_k = k->shape[0]; num_heads_k = k->shape[1];
    }

    const bool swapped = max_seqlen_q_in == 1 && num_heads_og > num_heads_k && wl_in < 0 && wr_in < 0 &&
                         head_size_og % 8 == 0 && !alibi_slopes;
    int64_t q_stride[4] = {0, 0, 0, 0}, o_stride[4] = {0, 0, 0, 0}, out_shape[4] = {0, 0, 0, 0};
    int q_rank, o_rank;
    int64_t max_seqlen_q = max_seqlen_q_in, num_heads = num_heads_og;
    if (swapped) {  // lib.rs:704-717
        const int64_t ngroups = num_heads_og / num_heads_k;
        out_shape[0] = batch_size; out_shape[1] = ngroups; out_shape[2] = num_heads_k; out_shape[3] = head_size_og;
        q_rank = o_rank = 4;
        q_stride[3] = 1; q_stride[2] = head_size_og; q_stride[1] = num_heads_k * head_size_og;
        q_stride[0] = ngroups * num_heads_k * head_size_og;
        std::memcpy(o_stride, q_stride, sizeof(q_stride));
        max_seqlen_q = ngroups;
        num_heads = num_heads_k;
    } else {
        q_rank = o_rank = 3;
        for (int i = 0; i < 3; ++i) { out_shape[i] = q->shape[i]; q_stride[i] = q->stride[i]; }
        o_stride[2] = 1; o_stride[1] = out_shape[2]; o_stride[0] = out_shape[1] * out_shape[2];
    }
    const int k_rank = k->rank, v_rank = v->rank;
    if (!paged && (q_rank != 3 || k_rank != 3 || v_rank != 3))
        return fail("flash-attn-varlen expects input tensors of rank 3 (q: %d, k: %d, v: %d", q_rank, k_rank, v_rank);
    if (paged && (q_rank != 3 || k_rank != 4 || v_rank != 4))
        return fail("flash-attn-varlen expects input tensors of rank 4 (q: %d, k: %d, v: %d", q_rank, k_rank, v_rank);
    if (q_stride[q_rank - 1] != 1) return fail("the last dim of q must be contiguous %s", dims_str(q_stride, q_rank).c_str());
    if (k->stride[k_rank - 1] != 1) return fail("the last dim of k must be contiguous %s", stride_str(k).c_str());
    if (v->stride[v_rank - 1] != 1) return fail("the last dim of v must be contiguous %s", stride_str(v).c_str());

    int64_t max_blocks = 0;
    if (paged) {
        if (block_table->rank != 2)
            return fail("unexpected rank, expected: 2, got: %d (%s)", block_table->rank, shape_str(block_table).c_str());
        max_blocks = block_table->shape[1];
        if (block_table->shape[0] != batch_size)
            return fail("shape mismatch of block_table (got %s) expected (%lld, %lld))", shape_str(block_table).c_str(),
                        static_cast<long long>(batch_size), static_cast<long long>(max_blocks));
    }
    const int64_t page_block_size = paged ? total_k : 1;
    if (paged && page_block_size % 16 != 0)
        return fail("page_block_size must be a multiple of 16, got %lld", static_cast<long long>(page_block_size));
    if (batch_size == 0) return fail("batch_size must be > 0");
    if (head_size_og > 256)
        return fail("only supports head dimension at most 256 (got %lld)", static_cast<long long>(head_size_og));
    if (head_size_og % 8 != 0)
        return fail("only supports head sizes that are a multiple of 8 (got %lld)", static_cast<long long>(head_size_og));
    if (num_heads % num_heads_k != 0)
        return fail("number of k/v heads %lld must divide number of heads in query %lld",
                    static_cast<long long>(num_heads_k), static_cast<long long>(num_heads));
    if (paged) {
        const int64_t exp[4] = {num_blocks, page_block_size, num_heads_k, head_size_og};
        if (std::memcmp(exp, k->shape, sizeof(exp)) != 0)
            return fail("shape mismatch of k (got %s) expected %s)", shape_str(k).c_str(), dims_str(exp, 4).c_str());
        if (v->rank != 4 || std::memcmp(exp, v->shape, sizeof(exp)) != 0)
            return fail("shape mismatch of v (got %s) expected %s)", shape_str(v).c_str(), dims_str(exp, 4).c_str());
    } else {
        const int64_t exp[3] = {total_k, num_heads_k, head_size_og};
        if (std::memcmp(exp, k->shape, sizeof(exp)) != 0)
            return fail("shape mismatch of k (got %s) expected %s)", shape_str(k).c_str(), dims_str(exp, 3).c_str());
        if (v->rank != 3 || std::memcmp(exp, v->shape, sizeof(exp)) != 0)
            return fail("shape mismatch of v (got %s) expected %s)", shape_str(v).c_str(), dims_str(exp, 3).c_str());
    }
    if (seqlens_k->shape[0] != batch_size + 1)
        return fail("shape mismatch of seqlens_k (got %s) expected %lld)", shape_str(seqlens_k).c_str(),
                    static_cast<long long>(batch_size + 1));
    if (nseqlens_q < 2) return fail("seqlens_q should have a len >= 2 %lld", static_cast<long long>(nseqlens_q));

    Alibi al;
    if (const char *e = check_alibi(alibi_slopes, num_heads, &al)) return e;
    int *seqused_ptr = nullptr;
    if (seqused_k) {
        if (seqused_k->device < 0) return fail("block_table must be a cuda tensor");  // sic, lib.rs:919
        if (seqused_k->stride[seqused_k->rank - 1] != 1) return fail("block_table must be contiguous");  // sic, :915
        seqused_ptr = static_cast<int *>(seqused_k->data);
    }

    // lib.rs:926-963
    int wl = (wl_in >= 0 && wl_in <= max_seqlen_k) ? static_cast<int>(wl_in) : -1;
    int wr = (wr_in >= 0 && wr_in <= max_seqlen_k) ? static_cast<int>(wr_in) : -1;
    const int64_t head_size = round_multiple(head_size_og, 8);
    const int64_t head_size_rounded = round_multiple(head_size, 32);
    const int is_causal = (wl < 0 && wr == 0) ? 1 : 0;
    if (wl < 0 && wr >= 0) wl = static_cast<int>(max_seqlen_k);
    if (wl >= 0 && wr < 0) wr = static_cast<int>(max_seqlen_k);
    const uint32_t num_splits =
        swapped ? atoma_compute_num_splits(batch_size, num_heads, head_size, max_seqlen_k, max_seqlen_q, num_sms) : 0;
    float softcap = softcap_in, scale = softmax_scale_in;
    if (softcap > 0.f) { softcap = softmax_scale_in / softcap; scale = softcap; } else softcap = 0.f;

    atoma_mha_args &a = plan->args;
    a.q_ptr = q->data; a.k_ptr = k->data; a.v_ptr = v->data;
    a.alibi_slopes_ptr = al.ptr;
    a.cu_seqlens_q_ptr = swapped ? nullptr : static_cast<int32_t *>(seqlens_q->data);
    a.cu_seqlens_k_ptr = static_cast<int32_t *>(seqlens_k->data);
    a.is_seqlens_k_cumulative = true;
    if (a.cu_seqlens_q_ptr) {
        a.q_batch_stride = a.o_batch_stride = 0;
    } else {  // swapped (lib.rs:1012-1015)
        a.q_batch_stride = static_cast<uint32_t>(q_stride[0] * max_seqlen_q);
        a.o_batch_stride = static_cast<uint32_t>(o_stride[0] * max_seqlen_q);
    }
    a.k_batch_stride = paged ? static_cast<uint32_t>(k->stride[0]) : 0;
    a.v_batch_stride = paged ? static_cast<uint32_t>(v->stride[0]) : 0;
    a.alibi_slopes_batch_stride = static_cast<uint32_t>(al.batch_stride);
    a.q_row_stride = static_cast<uint32_t>(q_stride[q_rank - 3]);
    a.k_row_stride = static_cast<uint32_t>(k->stride[k_rank - 3]);
    a.v_row_stride = static_cast<uint32_t>(v->stride[v_rank - 3]);
    a.o_row_stride = static_cast<uint32_t>(o_stride[o_rank - 3]);
    a.q_head_stride = static_cast<uint32_t>(q_stride[q_rank - 2]);
    a.k_head_stride = static_cast<uint32_t>(k->stride[k_rank - 2]);
    a.v_head_stride = static_cast<uint32_t>(v->stride[v_rank - 2]);
    a.o_head_stride = static_cast<uint32_t>(o_stride[o_rank - 2]);
    a.num_splits = num_splits;
    a.b = static_cast<uint32_t>(batch_size); a.h = static_cast<uint32_t>(num_heads);
    a.h_k = static_cast<uint32_t>(num_heads_k); a.d = static_cast<uint32_t>(head_size);
    a.d_rounded = static_cast<uint32_t>(head_size_rounded);
    a.softmax_scale = scale; a.scale_softmax_log2 = scale * 1.4426950408889634f;
    a.block_table = paged ? static_cast<int *>(block_table->data) : nullptr;
    a.block_table_batch_stride = paged ? static_cast<uint32_t>(block_table->stride[0]) : 0;
    a.page_block_size = static_cast<int>(page_block_size);
    a.seqused_k = seqused_ptr;
    a.seqlen_q = static_cast<uint32_t>(max_seqlen_q_in);  // sic: self.max_seqlen_q (lib.rs:1078)
    a.seqlen_k = static_cast<uint32_t>(max_seqlen_k);
    a.seqlen_q_rounded = static_cast<uint32_t>(round_multiple(max_seqlen_q_in, 128));
    a.seqlen_k_rounded = static_cast<uint32_t>(round_multiple(max_seqlen_k, 128));
    a.is_bf16 = q->dtype == ATOMA_BF16; a.is_causal = is_causal;
    a.window_size_left = wl; a.window_size_right = wr; a.softcap = softcap;
    a.unpadded_lse = true; a.force_split_kernel = paged;

    if (swapped) {  // lib.rs:1094-1096
        plan->out_rank = 4;
        plan->out_shape[0] = batch_size; plan->out_shape[1] = 1;
        plan->out_shape[2] = num_heads_k * max_seqlen_q; plan->out_shape[3] = head_size_og;
    } else {
        plan->out_rank = 3;
        std::memcpy(plan->out_shape, out_shape, sizeof(out_shape));
    }
    plan->lse_elems = total_q * num_heads;
    if (num_splits > 1) {
        plan->lseaccum_elems = static_cast<int64_t>(num_splits) * batch_size * num_heads * max_seqlen_q;
        plan->oaccum_elems = plan->lseaccum_elems * head_size_rounded;
    }
    plan->seqlenq_ngroups_swapped = swapped;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// flash_attn_kv_cache (+ block table)                                   csrc/src/lib.rs:1523-1855
// ------------------------------------------------------------------------------------------------
extern "C" const char *atoma_flash_attn_kv_cache_plan(
    const atoma_tensor *q, const atoma_tensor *kc, const atoma_tensor *vc,
    const atoma_tensor *alibi_slopes, const atoma_tensor *block_table,
    const atoma_tensor *seqlens_k, float softmax_scale, int64_t wl_in, int64_t wr_in, int num_sms,
    atoma_mha_plan *plan) {
    clear_plan(plan);
    if (!is_half(q->dtype)) return fail("flash-attn is only supported for f16/bf16 (%s)", dtype_name(q->dtype));
    if (q->dtype != kc->dtype) return fail("query and key must have the same dtype");
    if (q->dtype != vc->dtype) return fail("query and value must have the same dtype");
    const bool paged = block_table != nullptr;
    if (paged) {
        if (block_table->device < 0) return fail("block_table must be a cuda tensor");
        if (block_table->dtype != ATOMA_U32)
            return fail("unexpected dtype, expected: U32, got: %s", dtype_name(block_table->dtype));
        if (block_table->stride[block_table->rank - 1] != 1) return fail("block_table must be contiguous");
    }
    if (q->rank != 4) return fail("unexpected rank, expected: 4, got: %d (%s)", q->rank, shape_str(q).c_str());
    const int64_t batch_size = q->shape[0], seqlen_q = q->shape[1], num_heads = q->shape[2],
                  head_size_og = q->shape[3];
    int64_t max_blocks = 0;
    if (paged) {
        if (block_table->rank != 2)
            return fail("unexpected rank, expected: 2, got: %d (%s)", block_table->rank, shape_str(block_table).c_str());
        max_blocks = block_table->shape[1];
        if (block_table->shape[0] != batch_size)
            return fail("shape mismatch of block_table (got %s) expected (%lld, %lld))", shape_str(block_table).c_str(),
                        static_cast<long long>(batch_size), static_cast<long long>(max_blocks));
    }
    if (kc->rank != 4) return fail("unexpected rank, expected: 4, got: %d (%s)", kc->rank, shape_str(kc).c_str());
    int64_t page_block_size, seqlen_k, num_heads_k;
    if (paged) {
        page_block_size = kc->shape[1];
        seqlen_k = max_blocks * page_block_size;
        num_heads_k = kc->shape[2];
    } else {
        page_block_size = 1;
        seqlen_k = kc->shape[1];
        num_heads_k = kc->shape[2];
    }
    if (paged && page_block_size % 16 != 0)
        return fail("page_block_size must be a multiple of 16 when block_table is provided");

    // lib.rs:1610-1641
    int wl = wl_in >= 0 ? (wl_in >= seqlen_k ? -1 : static_cast<int>(wl_in)) : -1;
    int wr = wr_in >= 0 ? (wr_in >= seqlen_k ? -1 : static_cast<int>(wr_in)) : -1;
    bool is_causal = wl < 0 && wr == 0;
    if (seqlen_q == 1 && !alibi_slopes) is_causal = false;
    if (is_causal) wr = 0;
    if (wl < 0 && wr >= 0) wl = static_cast<int>(seqlen_k);
    if (wr < 0 && wl >= 0) wr = static_cast<int>(seqlen_k);

    if (vc->rank != 4)
        return fail("flash-attn expects input tensors of rank 4 (q: %d, k: %d, v: %d)", q->rank, kc->rank, vc->rank);
    if (q->stride[3] != 1) return fail("the last dim of q must be contiguous %s", stride_str(q).c_str());
    if (kc->stride[3] != 1) return fail("the last dim of k must be contiguous %s", stride_str(kc).c_str());
    if (vc->stride[3] != 1) return fail("the last dim of v must be contiguous %s", stride_str(vc).c_str());
    Alibi al;
    if (const char *e = check_alibi(alibi_slopes, num_heads, &al)) return e;

    const int64_t head_size = round_multiple(head_size_og, 8);
    const int64_t head_size_rounded = round_multiple(head_size, 32);

    int32_t *cu_k = nullptr;
    if (seqlens_k) {
        if (seqlens_k->rank != 1 || seqlens_k->shape[0] != batch_size)
            return fail("shape mismatch of seqlens_k (got %s) expected [%lld])", shape_str(seqlens_k).c_str(),
                        static_cast<long long>(batch_size));
        if (seqlens_k->dtype != ATOMA_U32)
            return fail("DType mismatch seqlens_k %s, expected U32", dtype_name(seqlens_k->dtype));
        if (seqlens_k->device < 0) return fail("seqlens_k must be a cuda tensor");
        if (seqlens_k->stride[0] != 1)
            return fail("the last dim of seqlens_k must be contiguous %s", stride_str(seqlens_k).c_str());
        cu_k = static_cast<int32_t *>(seqlens_k->data);
    }
    const uint32_t num_splits =
        atoma_compute_num_splits(batch_size, num_heads, head_size, seqlen_k, seqlen_q, num_sms);
    if (num_splits > 128) return fail("num_splits > 128 not supported");

    const int64_t o_stride[4] = {seqlen_q * num_heads * head_size_og, num_heads * head_size_og, head_size_og, 1};
    atoma_mha_args &a = plan->args;
    a.q_ptr = q->data; a.k_ptr = kc->data; a.v_ptr = vc->data;
    a.alibi_slopes_ptr = al.ptr;
    a.cu_seqlens_q_ptr = nullptr;
    a.cu_seqlens_k_ptr = cu_k;
    a.is_seqlens_k_cumulative = seqlens_k == nullptr;
    a.q_batch_stride = static_cast<uint32_t>(q->stride[0]); a.k_batch_stride = static_cast<uint32_t>(kc->stride[0]);
    a.v_batch_stride = static_cast<uint32_t>(vc->stride[0]); a.o_batch_stride = static_cast<uint32_t>(o_stride[0]);
    a.alibi_slopes_batch_stride = static_cast<uint32_t>(al.batch_stride);
    a.q_row_stride = static_cast<uint32_t>(q->stride[1]); a.k_row_stride = static_cast<uint32_t>(kc->stride[1]);
    a.v_row_stride = static_cast<uint32_t>(vc->stride[1]); a.o_row_stride = static_cast<uint32_t>(o_stride[1]);
    a.q_head_stride = static_cast<uint32_t>(q->stride[2]); a.k_head_stride = static_cast<uint32_t>(kc->stride[2]);
    a.v_head_stride = static_cast<uint32_t>(vc->stride[2]); a.o_head_stride = static_cast<uint32_t>(o_stride[2]);
    a.num_splits = num_splits;
    a.b = static_cast<uint32_t>(batch_size); a.h = static_cast<uint32_t>(num_heads);
    a.h_k = static_cast<uint32_t>(num_heads_k); a.d = static_cast<uint32_t>(head_size);
    a.d_rounded = static_cast<uint32_t>(head_size_rounded);
    a.softmax_scale = softmax_scale; a.scale_softmax_log2 = softmax_scale * 1.4426950408889634f;
    a.block_table = paged ? static_cast<int *>(block_table->data) : nullptr;
    a.block_table_batch_stride = paged ? static_cast<uint32_t>(block_table->stride[0]) : 0;
    a.page_block_size = static_cast<int>(page_block_size);
    a.seqused_k = nullptr;
    a.seqlen_q = static_cast<uint32_t>(seqlen_q); a.seqlen_k = static_cast<uint32_t>(seqlen_k);
    a.seqlen_q_rounded = static_cast<uint32_t>(round_multiple(seqlen_q, 128));
    a.seqlen_k_rounded = static_cast<uint32_t>(round_multiple(seqlen_k, 128));
    a.is_bf16 = q->dtype == ATOMA_BF16; a.is_causal = is_causal ? 1 : 0;
    a.window_size_left = wl; a.window_size_right = wr; a.softcap = 0.f;
    a.unpadded_lse = false; a.force_split_kernel = paged;

    plan->out_rank = 4;
    for (int i = 0; i < 4; ++i) plan->out_shape[i] = q->shape[i];
    plan->lse_elems = batch_size * num_heads * seqlen_q;
    if (num_splits > 1) {
        plan->lseaccum_elems = static_cast<int64_t>(num_splits) * batch_size * num_heads * seqlen_q;
        plan->oaccum_elems = plan->lseaccum_elems * head_size_rounded;
    }
    return nullptr;
}

extern "C" const char *atoma_mha_execute(const atoma_mha_plan *plan, void *out, void *softmax_lse,
                                         void *softmax_lseaccum, void *oaccum,
                                         int use_legacy_stream, atoma_stream_t stream) {
    if (!out) return fail("atoma_mha_execute: output buffer is required");
    const atoma_mha_args &a = plan->args;
    if (use_legacy_stream) {
        run_mha(a.q_ptr, a.k_ptr, a.v_ptr, out, softmax_lse, a.alibi_slopes_ptr, a.cu_seqlens_q_ptr,
                a.cu_seqlens_k_ptr, a.is_seqlens_k_cumulative, a.q_batch_stride, a.k_batch_stride,
                a.v_batch_stride, a.o_batch_stride, a.alibi_slopes_batch_stride, a.q_row_stride,
                a.k_row_stride, a.v_row_stride, a.o_row_stride, a.q_head_stride, a.k_head_stride,
                a.v_head_stride, a.o_head_stride, a.num_splits, a.b, a.h, a.h_k, a.d, a.d_rounded,
                a.softmax_scale, a.scale_softmax_log2, a.block_table, a.block_table_batch_stride,
                a.page_block_size, a.seqused_k, a.seqlen_q, a.seqlen_k, a.seqlen_q_rounded,
                a.seqlen_k_rounded, a.is_bf16, a.is_causal, a.window_size_left, a.window_size_right,
                a.softcap, a.unpadded_lse, a.force_split_kernel, softmax_lseaccum, oaccum);
    } else {
        run_mha_stream(a.q_ptr, a.k_ptr, a.v_ptr, out, softmax_lse, a.alibi_slopes_ptr,
                       a.cu_seqlens_q_ptr, a.cu_seqlens_k_ptr, a.is_seqlens_k_cumulative,
                       a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.o_batch_stride,
                       a.alibi_slopes_batch_stride, a.q_row_stride, a.k_row_stride, a.v_row_stride,
                       a.o_row_stride, a.q_head_stride, a.k_head_stride, a.v_head_stride,
                       a.o_head_stride, a.num_splits, a.b, a.h, a.h_k, a.d, a.d_rounded,
                       a.softmax_scale, a.scale_softmax_log2, a.block_table,
                       a.block_table_batch_stride, a.page_block_size, a.seqused_k, a.seqlen_q,
                       a.seqlen_k, a.seqlen_q_rounded, a.seqlen_k_rounded, a.is_bf16, a.is_causal,
                       a.window_size_left, a.window_size_right, a.softcap, a.unpadded_lse,
                       a.force_split_kernel, softmax_lseaccum, oaccum, stream);
    }
    if (atoma_last_status() != 0) return fail("%s", atoma_last_error());
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// cache ops                                                    csrc/src/cache_manager.rs
// ------------------------------------------------------------------------------------------------
extern "C" const char *atoma_reshape_and_cache_flash(const atoma_tensor *key, const atoma_tensor *value,
                                                     const atoma_tensor *key_cache,
                                                     const atoma_tensor *value_cache,
                                                     const atoma_tensor *slot_mapping,
                                                     atoma_stream_t stream) {
    // cache_manager.rs:326-344
    if (key->dtype != value->dtype || key->dtype != key_cache->dtype || key->dtype != value_cache->dtype)
        return fail("Only support f16/bf16 dtypes and key, value, key_cache and value_cache must have same dtype");
    if (!is_half(key->dtype)) return fail("Only support f16/bf16 dtypes must have same dtype");
    if (key->device < 0) return fail("device must be a cuda device");
    if (value->device < 0 || key_cache->device < 0 || value_cache->device < 0 || slot_mapping->device < 0)
        return fail("value, key_cache and value_cache must be on the same device");
    if (value->device != key->device || key_cache->device != key->device ||
        value_cache->device != key->device || slot_mapping->device != key->device)
        return fail("key, value, key_cache, value_cache and slot_mapping must be on the same device");
    if (key->rank < 3 || key_cache->rank < 2) return fail("Only support key and value tensors with rank 3 (got %d and v_rank %d)", key->rank, value->rank);
    const int64_t num_tokens = key->shape[0], num_heads = key->shape[1], head_size = key->shape[2];
    const int64_t num_blocks = key_cache->shape[0], block_size = key_cache->shape[1];
    const int64_t key_stride = key->stride[0], value_stride = value->stride[0];
    const int64_t block_stride = key_cache->stride[0];
    if (block_stride != value_cache->stride[0])
        return fail("Only support block_stride == value_cache.stride[0] (got block_stride %lld and value_cache.stride[0] %lld)",
                    static_cast<long long>(block_stride), static_cast<long long>(value_cache->stride[0]));
    if (key->rank != 3 || value->rank != 3)
        return fail("Only support key and value tensors with rank 3 (got %d and v_rank %d)", key->rank, value->rank);
    if (key_cache->rank != 4) return fail("Only support key_cache tensors with rank 4 (got %d)", key_cache->rank);
    if (value_cache->rank != 4) return fail("Only support value_cache tensors with rank 4 (got %d)", value_cache->rank);
    const int64_t cexp[4] = {num_blocks, block_size, num_heads, head_size};
    if (std::memcmp(cexp, key_cache->shape, sizeof(cexp)) != 0)
        return fail("Only support key_cache with shape %s (got %s)", dims_str(cexp, 4).c_str(), shape_str(key_cache).c_str());
    if (std::memcmp(cexp, value_cache->shape, sizeof(cexp)) != 0)
        return fail("Only support value_cache with shape %s (got %s)", dims_str(cexp, 4).c_str(), shape_str(value_cache).c_str());
    const int64_t vexp[3] = {num_tokens, num_heads, head_size};
    if (std::memcmp(vexp, value->shape, sizeof(vexp)) != 0)
        return fail("Only support value with shape %s (got %s)", dims_str(vexp, 3).c_str(), shape_str(value).c_str());
    if (slot_mapping->rank != 1 || slot_mapping->shape[0] != num_tokens)
        return fail("Only support slot_mapping with shape [%lld] (got %s)", static_cast<long long>(num_tokens),
                    shape_str(slot_mapping).c_str());
    if (slot_mapping->dtype != ATOMA_I64)
        return fail("unexpected dtype, expected: I64, got: %s", dtype_name(slot_mapping->dtype));
    reshape_and_cache_flash(key->data, value->data, key_cache->data, value_cache->data,
                            static_cast<int64_t *>(slot_mapping->data), block_stride, num_tokens,
                            num_heads, head_size, block_size, key_stride, value_stride,
                            key->dtype == ATOMA_BF16 ? 1u : 0u, stream);
    if (atoma_last_status() != 0) return fail("%s", atoma_last_error());
    return nullptr;
}

extern "C" const char *atoma_copy_blocks(const atoma_tensor *key_caches, int64_t num_key_caches,
                                         const atoma_tensor *value_caches, int64_t num_value_caches,
                                         const atoma_tensor *block_mapping, void *ptr_scratch,
                                         atoma_stream_t stream) {
    // cache_manager.rs:148-307
    if (num_key_caches > 0 && num_value_caches > 0) {
        const int kd = key_caches[0].dtype, vd = value_caches[0].dtype;
        if (!((kd == ATOMA_F16 && vd == ATOMA_F16) || (kd == ATOMA_BF16 && vd == ATOMA_BF16)))
            return fail("Only support f16/bf16 dtypes and src and dst must have same dtype");
        if (key_caches[0].device < 0) return fail("device must be a cuda device");
    }
    const int64_t num_layers = num_key_caches;
    if (num_layers != num_value_caches) return fail("key_caches and value_caches must have the same length");
    if (num_layers == 0) return nullptr;
    if (value_caches[0].device < 0) return fail("key_caches and value_caches must be on the same device");
    std::vector<int64_t> ptrs(2 * num_layers);
    for (int64_t l = 0; l < num_layers; ++l) {
        if (key_caches[l].device < 0) return fail("key_caches must be a cuda tensor");
        if (value_caches[l].device < 0) return fail("value_caches must be a cuda tensor");
        ptrs[l] = reinterpret_cast<int64_t>(key_caches[l].data);
        ptrs[num_layers + l] = reinterpret_cast<int64_t>(value_caches[l].data);
    }
    if (block_mapping->rank != 2 || block_mapping->shape[1] != 2)
        return fail("block_mapping must have shape [num_pairs, 2]");
    const int64_t num_pairs = block_mapping->shape[0];
    if (block_mapping->device < 0) return fail("block_mapping must be a cuda tensor");
    if (block_mapping->dtype != ATOMA_I64)
        return fail("unexpected dtype, expected: I64, got: %s", dtype_name(block_mapping->dtype));
    int64_t numel_per_block = 1;
    for (int i = 1; i < key_caches[0].rank; ++i) numel_per_block *= key_caches[0].shape[i];
    if (!ptr_scratch) return fail("copy_blocks: ptr_scratch (2 * num_layers int64 on the device) is required");
    cudaError_t e = cudaMemcpyAsync(ptr_scratch, ptrs.data(), ptrs.size() * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail("copy_blocks: pointer upload: %s", cudaGetErrorString(e));
    int64_t *dptr = static_cast<int64_t *>(ptr_scratch);
    if (key_caches[0].dtype == ATOMA_F16)
        copy_blocks_f16(dptr, dptr + num_layers, block_mapping->data, num_layers, num_pairs, numel_per_block, stream);
    else
        copy_blocks_bf16(dptr, dptr + num_layers, block_mapping->data, num_layers, num_pairs, numel_per_block, stream);
    if (atoma_last_status() != 0) return fail("%s", atoma_last_error());
    return nullptr;
}

extern "C" const char *atoma_swap_blocks(const atoma_tensor *src, const atoma_tensor *dst,
                                         const int64_t *src_blocks, const int64_t *dst_blocks,
                                         int64_t num_pairs, atoma_stream_t stream) {
    // cache_manager.rs:18-128
    int64_t elems = 1;
    for (int i = 1; i < src->rank; ++i) elems *= src->shape[i];
    const int64_t esize = src->dtype == ATOMA_F32 ? 4 : (src->dtype == ATOMA_I64 || src->dtype == ATOMA_F64 ? 8 : (src->dtype == ATOMA_U8 ? 1 : (src->dtype == ATOMA_U32 ? 4 : 2)));
    const int64_t block_bytes = esize * elems;
    int kind;
    if (src->device >= 0 && dst->device >= 0) {
        if (src->device != dst->device)
            return fail("swap_blocks: Both src and dst tensors should be on the same device to swap");
        kind = 0;
    } else if (src->device < 0 && dst->device >= 0) {
        if (!is_half(src->dtype))
            return fail("swap_blocks: Invalid combination of src and dst tensors storage to swap");
        kind = 1;
    } else if (src->device >= 0 && dst->device < 0) {
        if (!is_half(src->dtype))
            return fail("swap_blocks:Invalid dtype for cuda src tensor, expected f16/bf16, got %s", dtype_name(src->dtype));
        kind = 2;
    } else {
        return fail("swap_blocks: Either src and dst are on the same cuda device, or src and dst are on cpu and cuda devices, alternately");
    }
    swap_blocks_batched(dst->data, src->data, src_blocks, dst_blocks, num_pairs, block_bytes, kind, stream);
    if (atoma_last_status() != 0) return fail("%s", atoma_last_error());
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// models::FlashAttention::forward                         models/src/flash_attention.rs:322-469
// ------------------------------------------------------------------------------------------------
namespace {
atoma_tensor narrow0(const atoma_tensor &t, int64_t start, int64_t len) {  // tensor.i(start..start+len)
    atoma_tensor r = t;
    const int64_t esize = (t.dtype == ATOMA_F32 || t.dtype == ATOMA_U32) ? 4 : (t.dtype == ATOMA_I64 || t.dtype == ATOMA_F64 ? 8 : (t.dtype == ATOMA_U8 ? 1 : 2));
    r.data = static_cast<char *>(t.data) + start * t.stride[0] * esize;
    r.shape[0] = len;
    return r;
}
atoma_tensor drop_dim0(const atoma_tensor &t, int64_t index) {  // tensor.i(index)
    atoma_tensor r = narrow0(t, index, 1);
    for (int i = 1; i < t.rank; ++i) { r.shape[i - 1] = t.shape[i]; r.stride[i - 1] = t.stride[i]; }
    r.rank = t.rank - 1;
    return r;
}
}  // namespace

extern "C" const char *atoma_flash_attention_forward(const atoma_attention_config *cfg,
                                                     const atoma_tensor *q, const atoma_tensor *k,
                                                     const atoma_tensor *v, const atoma_tensor *kv_cache,
                                                     const atoma_tensor *alibi_slopes,
                                                     const atoma_attention_metadata *md,
                                                     const atoma_tensor *out, atoma_stream_t stream) {
    if (q->rank != 3 || k->rank != 3 || v->rank != 3)
        return fail("unexpected rank, expected: 3, got: %d", q->rank != 3 ? q->rank : (k->rank != 3 ? k->rank : v->rank));
    const int64_t qT = q->shape[0], qH = q->shape[1], qD = q->shape[2];
    const int64_t kT = k->shape[0], kH = k->shape[1], kD = k->shape[2];
    const int64_t vT = v->shape[0], vH = v->shape[1], vD = v->shape[2];
    if (qT != kT || qT != vT)
        return fail("query, key, and value must have the same number of tokens (got %lld, %lld, %lld)",
                    static_cast<long long>(qT), static_cast<long long>(kT), static_cast<long long>(vT));
    if (kT != vT || kH != vH || kD != vD)
        return fail("key and value must have the same shape (got [%lld, %lld, %lld], [%lld, %lld, %lld])",
                    static_cast<long long>(kT), static_cast<long long>(kH), static_cast<long long>(kD),
                    static_cast<long long>(vT), static_cast<long long>(vH), static_cast<long long>(vD));
    if (qH != cfg->num_heads || qD != cfg->head_dim)
        return fail("query must have [num_head, hidden_dim] = [%lld, %lld] (got [%lld, %lld])",
                    static_cast<long long>(cfg->num_heads), static_cast<long long>(cfg->head_dim),
                    static_cast<long long>(qH), static_cast<long long>(qD));
    if (kH != cfg->num_kv_heads || kD != cfg->head_dim)
        return fail("key must have k_num_heads = %lld and hidden dim %lld (got %lld, %lld)",
                    static_cast<long long>(cfg->num_kv_heads), static_cast<long long>(cfg->head_dim),
                    static_cast<long long>(kH), static_cast<long long>(kD));
    // split_kv_cache (flash_attention.rs:249-278)
    if (kv_cache->rank != 5) return fail("KV cache must have rank 5 (got %d)", kv_cache->rank);
    if (kv_cache->shape[0] != 2) return fail("KV cache must have cache_size 2 (got %lld)", static_cast<long long>(kv_cache->shape[0]));
    if (kv_cache->shape[3] != cfg->num_kv_heads)
        return fail("KV cache must have num_heads %lld (got %lld)", static_cast<long long>(cfg->num_kv_heads),
                    static_cast<long long>(kv_cache->shape[3]));
    if (kv_cache->shape[4] != cfg->head_dim)
        return fail("KV cache must have head_dim %lld (got %lld)", static_cast<long long>(cfg->head_dim),
                    static_cast<long long>(kv_cache->shape[4]));
    const atoma_tensor k_cache = drop_dim0(*kv_cache, 0), v_cache = drop_dim0(*kv_cache, 1);

    if (const char *e = atoma_reshape_and_cache_flash(k, v, &k_cache, &v_cache, md->slot_mapping, stream)) return e;

    const int64_t P = md->num_prefill_tokens, Dn = md->num_decoding_tokens;
    if (kT != P + Dn)
        return fail("query must have number of tokens %lld (got %lld)", static_cast<long long>(P + Dn),
                    static_cast<long long>(qT));
    if (out->rank != 2 || out->shape[0] != qT || out->shape[1] != cfg->num_heads * cfg->head_dim ||
        out->stride[1] != 1 || out->stride[0] != out->shape[1] || out->dtype != q->dtype)
        return fail("output must be a contiguous [%lld, %lld] tensor of the query dtype", static_cast<long long>(qT),
                    static_cast<long long>(cfg->num_heads * cfg->head_dim));
    const int64_t esize = 2;
    char *out_base = static_cast<char *>(out->data);
    atoma_mha_plan plan;

    if (md->has_prefill && P > 0) {
        const atoma_tensor qp = narrow0(*q, 0, P), kp = narrow0(*k, 0, P), vp = narrow0(*v, 0, P);
        const bool no_prefix = !md->prefill_block_tables || md->prefill_block_tables->shape[0] * md->prefill_block_tables->shape[1] == 0;
        if (no_prefix) {
            if (!md->sequence_start_locations)
                return fail("Missing sequence start locations tensor for prefill inference");
            // flash_attn_varlen(q, k, v, ssl, ssl, max, max, scale, causal = q_num_tokens > 1)
            const int64_t wr = qT > 1 ? 0 : -1;
            if (const char *e = atoma_flash_attn_varlen_plan(
                    &qp, &kp, &vp, nullptr, md->sequence_start_locations, md->sequence_start_locations, nullptr,
                    nullptr, md->max_prefill_sequence_length, md->max_prefill_sequence_length, cfg->softmax_scale,
                    -1, wr, 0.f, 0, &plan))
                return e;
        } else {
            if (!md->query_start_locations)
                return fail("Missing query start locations tensor for prefill inference, with prefix enabled attention");
            if (!md->sequence_start_locations)
                return fail("Missing sequence start locations tensor for prefill inference, with prefix enabled attention");
            // flash_attn_varlen_with_block_table(..., sliding_window, None, block_tables)
            if (const char *e = atoma_flash_attn_varlen_plan(
                    &qp, &k_cache, &v_cache, alibi_slopes, md->query_start_locations, md->sequence_start_locations,
                    nullptr, md->prefill_block_tables, md->max_prefill_sequence_length,
                    md->max_prefill_sequence_length_k, cfg->softmax_scale, cfg->sliding_window, -1, 0.f, 0, &plan))
                return e;
        }
        if (plan.seqlenq_ngroups_swapped)
            return fail("prefill with max_seqlen_q == 1 takes the reference's head-swapped layout; not routed here");
        if (const char *e = atoma_mha_execute(&plan, out_base, nullptr, nullptr, nullptr, 0, stream)) return e;
    }
    if (md->has_decode && Dn > 0) {
        // decode_q.unsqueeze(1): [Dn, 1, H, D]
        atoma_tensor dq = narrow0(*q, P, Dn);
        atoma_tensor dq4 = dq;
        dq4.rank = 4;
        dq4.shape[0] = Dn; dq4.shape[1] = 1; dq4.shape[2] = qH; dq4.shape[3] = qD;
        dq4.stride[0] = dq.stride[0]; dq4.stride[1] = dq.stride[0]; dq4.stride[2] = dq.stride[1]; dq4.stride[3] = dq.stride[2];
        if (const char *e = atoma_flash_attn_kv_cache_plan(&dq4, &k_cache, &v_cache, alibi_slopes,
                                                           md->decode_block_tables, md->decode_sequence_lengths,
                                                           cfg->softmax_scale, -1, 0, 0, &plan))
            return e;
        if (const char *e = atoma_mha_execute(&plan, out_base + P * cfg->num_heads * cfg->head_dim * esize, nullptr,
                                              nullptr, nullptr, 0, stream))
            return e;
    }
    return nullptr;
}
