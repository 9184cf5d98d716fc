// Device-side parameter block shared by the attention kernels.  Field meaning follows the
// reference's Flash_fwd_params (csrc/kernels/flash.h:18-134) for the live subset only; strides are
// in elements, widened to 64-bit here.
#pragma once
#include <cstdint>

namespace atoma {

struct MhaParams {
    const void *q;
    const void *k;
    const void *v;
    void *o;
    float *lse;           // may be null
    const float *alibi;   // may be null; [h] or [b,h] (alibi_bs = 0 / row stride)
    const int *cu_q;      // may be null (dense batch layout)
    const int *cu_k;      // may be null
    const int *seqused_k; // may be null
    const int *block_table;  // may be null (contiguous K/V)

    int64_t q_bs, k_bs, v_bs, o_bs;  // batch strides (k_bs/v_bs = page stride when paged)
    int64_t q_rs, k_rs, v_rs, o_rs;  // row (token) strides
    int64_t q_hs, k_hs, v_hs, o_hs;  // head strides
    int64_t alibi_bs;
    int64_t bt_bs;  // block_table row stride

    int b, h, h_k, d;
    int seqlen_q, seqlen_k;
    int page;  // page_block_size (tokens per KV block)
    float scale;
    float scale_log2;
    int is_causal;
    int k_cumulative;  // cu_k holds b+1 cumulative offsets (1) or b per-sequence lengths (0)
    int unpadded_lse;
};

}  // namespace atoma
