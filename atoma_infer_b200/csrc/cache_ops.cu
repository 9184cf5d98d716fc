// KV-cache maintenance kernels for sm_100a: reshape_and_cache_flash, copy_blocks, swap_blocks.
//
// Semantics follow the reference's csrc/kernels/cache_manager.cu (copy_blocks_kernel :16-37,
// reshape_and_cache_flash_kernel :140-170) and csrc/src/cache_manager.rs:18-128 (swap_blocks);
// the implementation is new: whole cache blocks move as cp.async.bulk (TMA 1-D) global->shared->
// global transfers driven by one thread per CTA, token rows move as 16-byte vectors.  All of it is
// pure data movement, bit-exact by construction.
#include "atoma_flash.h"
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace atoma {
namespace {

// ---------------------------------------------------------------------------------------------
// Bulk-async block mover
// ---------------------------------------------------------------------------------------------
constexpr int kPieceBytes = 16 * 1024;  // one bulk transfer
constexpr int kRing = 4;                // ring slots per CTA (64 KiB smem -> 3 CTAs / SM)

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Moves `nspans` spans of `bytes` bytes each (16-byte aligned, bytes % 16 == 0) through a
// kRing-deep shared-memory ring.  Called by ONE thread of the CTA.
struct Span {
    const char *src;
    char *dst;
};

template <typename SpanFn>
__device__ void bulk_move(SpanFn span_of, int nspans, int64_t bytes, char *ring, uint64_t *bars) {
    const int pieces_per_span = static_cast<int>((bytes + kPieceBytes - 1) / kPieceBytes);
    const int total = nspans * pieces_per_span;
    auto piece = [&](int idx, const char *&src, char *&dst, uint32_t &n) {
        const int s = idx / pieces_per_span, p = idx % pieces_per_span;
        const Span sp = span_of(s);
        const int64_t off = static_cast<int64_t>(p) * kPieceBytes;
        n = static_cast<uint32_t>(min(static_cast<int64_t>(kPieceBytes), bytes - off));
        src = sp.src + off;
        dst = sp.dst + off;
    };
    auto issue_load = [&](int idx) {
        const char *src;
        char *dst;
        uint32_t n;
        piece(idx, src, dst, n);
        const int slot = idx % kRing;
        mbar_arrive_expect_tx(&bars[slot], n);
        bulk_g2s(ring + slot * kPieceBytes, src, n, &bars[slot]);
    };
    // Loads run kRing-1 pieces ahead of the stores.  Piece i lives in slot i % kRing and completes
    // phase i / kRing of that slot's mbarrier.
    for (int i = 0; i < total && i < kRing - 1; ++i) issue_load(i);
    for (int idx = 0; idx < total; ++idx) {
        const int ahead = idx + kRing - 1;
        if (ahead < total) {
            // `ahead` reuses the slot piece idx-1 was stored from: that store must have finished
            // READING shared memory (not necessarily writing global memory).
            if (idx > 0) bulk_wait_read<0>();
            issue_load(ahead);
        }
        const int slot = idx % kRing;
        mbar_wait(&bars[slot], (idx / kRing) & 1);
        const char *src;
        char *dst;
        uint32_t n;
        piece(idx, src, dst, n);
        bulk_s2g(dst, ring + slot * kPieceBytes, n);
        bulk_commit();
    }
    bulk_wait_read<0>();
}

__global__ void __launch_bounds__(32)
copy_blocks_bulk_kernel(const int64_t *__restrict__ key_cache_ptrs,
                        const int64_t *__restrict__ value_cache_ptrs,
                        const int64_t *__restrict__ block_mapping, int64_t block_bytes) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t bars[kRing];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < kRing; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    const int layer = blockIdx.x;
    const int64_t pair = blockIdx.y;
    const int64_t src_block = block_mapping[2 * pair];
    const int64_t dst_block = block_mapping[2 * pair + 1];
    char *kbase = reinterpret_cast<char *>(key_cache_ptrs[layer]);
    char *vbase = reinterpret_cast<char *>(value_cache_ptrs[layer]);
    auto span_of = [&](int s) {
        char *base = s == 0 ? kbase : vbase;
        return Span{base + src_block * block_bytes, base + dst_block * block_bytes};
    };
    bulk_move(span_of, 2, block_bytes, ring, bars);
}

__global__ void __launch_bounds__(32)
swap_blocks_bulk_kernel(char *__restrict__ dst_base, const char *__restrict__ src_base,
                        const int64_t *__restrict__ src_blocks,
                        const int64_t *__restrict__ dst_blocks, int64_t block_bytes) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t bars[kRing];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < kRing; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    const int64_t pair = blockIdx.x;
    const int64_t sb = src_blocks[pair], db = dst_blocks[pair];
    auto span_of = [&](int) { return Span{src_base + sb * block_bytes, dst_base + db * block_bytes}; };
    bulk_move(span_of, 1, block_bytes, ring, bars);
}

// Fallback for blocks that are not 16-byte aligned / sized: 2-byte elements, same indexing as the
// reference kernel (cache_manager.cu:31-36).
__global__ void copy_blocks_scalar_kernel(const int64_t *__restrict__ key_cache_ptrs,
                                          const int64_t *__restrict__ value_cache_ptrs,
                                          const int64_t *__restrict__ block_mapping,
                                          int64_t numel_per_block) {
    const int layer = blockIdx.x;
    const int64_t pair = blockIdx.y;
    uint16_t *kc = reinterpret_cast<uint16_t *>(key_cache_ptrs[layer]);
    uint16_t *vc = reinterpret_cast<uint16_t *>(value_cache_ptrs[layer]);
    const int64_t so = block_mapping[2 * pair] * numel_per_block;
    const int64_t dofs = block_mapping[2 * pair + 1] * numel_per_block;
    for (int64_t i = threadIdx.x; i < numel_per_block; i += blockDim.x) {
        kc[dofs + i] = kc[so + i];
        vc[dofs + i] = vc[so + i];
    }
}

__global__ void swap_blocks_scalar_kernel(char *__restrict__ dst_base,
                                          const char *__restrict__ src_base,
                                          const int64_t *__restrict__ src_blocks,
                                          const int64_t *__restrict__ dst_blocks,
                                          int64_t block_bytes) {
    const int64_t pair = blockIdx.x;
    const char *s = src_base + src_blocks[pair] * block_bytes;
    char *d = dst_base + dst_blocks[pair] * block_bytes;
    for (int64_t i = threadIdx.x; i < block_bytes; i += blockDim.x) d[i] = s[i];
}

// ---------------------------------------------------------------------------------------------
// reshape_and_cache_flash
// ---------------------------------------------------------------------------------------------
// One "chunk" = 16 bytes of one token's [num_heads*head_size] row.  A thread moves the same chunk
// of K and of V for kUnroll tokens per iteration (all loads issued before the stores).
constexpr int kRcThreads = 256;
constexpr int kRcUnroll = 4;

__global__ void __launch_bounds__(kRcThreads)
reshape_and_cache_flash_vec_kernel(const char *__restrict__ key, const char *__restrict__ value,
                                   char *__restrict__ k_cache, char *__restrict__ v_cache,
                                   const int64_t *__restrict__ slot_mapping, int64_t num_tokens,
                                   int chunks_per_token, int64_t key_stride_b,
                                   int64_t value_stride_b, int64_t block_stride_b,
                                   int64_t row_bytes, int64_t block_size) {
    pdl_launch_dependents();
    pdl_wait();  // key / value come from the kernel before this one in the stream
    const int64_t total = num_tokens * chunks_per_token;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kRcThreads;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * kRcThreads + threadIdx.x; base < total;
         base += stride * kRcUnroll) {
        uint4 kv[kRcUnroll], vv[kRcUnroll];
        int64_t dst_off[kRcUnroll];
#pragma unroll
        for (int u = 0; u < kRcUnroll; ++u) {
            const int64_t idx = base + u * stride;
            dst_off[u] = -1;
            if (idx < total) {
                const int64_t token = idx / chunks_per_token;
                const int chunk = static_cast<int>(idx - token * chunks_per_token);
                const int64_t slot = slot_mapping[token];
                if (slot >= 0) {  // slot < 0: padded token, skipped (cache_manager.cu:151-155)
                    const int64_t blk = slot / block_size, off = slot % block_size;
                    dst_off[u] = blk * block_stride_b + off * row_bytes + chunk * 16;
                    kv[u] = ld_nc_v4(key + token * key_stride_b + chunk * 16);
                    vv[u] = ld_nc_v4(value + token * value_stride_b + chunk * 16);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kRcUnroll; ++u) {
            if (dst_off[u] >= 0) {
                *reinterpret_cast<uint4 *>(k_cache + dst_off[u]) = kv[u];
                *reinterpret_cast<uint4 *>(v_cache + dst_off[u]) = vv[u];
            }
        }
    }
}

__global__ void reshape_and_cache_flash_scalar_kernel(
    const uint16_t *__restrict__ key, const uint16_t *__restrict__ value,
    uint16_t *__restrict__ k_cache, uint16_t *__restrict__ v_cache,
    const int64_t *__restrict__ slot_mapping, int64_t block_stride, int64_t key_stride,
    int64_t value_stride, int64_t n, int64_t block_size) {
    const int64_t token = blockIdx.x;
    const int64_t slot = slot_mapping[token];
    if (slot < 0) return;
    const int64_t blk = slot / block_size, off = slot % block_size;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t tgt = blk * block_stride + off * n + i;
        k_cache[tgt] = key[token * key_stride + i];
        v_cache[tgt] = value[token * value_stride + i];
    }
}


// ---------------------------------------------------------------------------------------------
// Fused rotary embedding + cache write
// ---------------------------------------------------------------------------------------------
// What the reference does with two index_select, two rope kernels, four transpose+contiguous copies
// and reshape_and_cache_flash (models/src/llama.rs:218-251, 269-304; flash_attention.rs:361) in one
// pass: q and k are rotated in place (non-interleaved "rotate half": out[:d/2] = x1 cos - x2 sin,
// out[d/2:] = x1 sin + x2 cos, cos/sin rows selected by the token's position), the rotated k and
// the untouched v go to their cache slot.  fp32 arithmetic, one rounding to the io type.
// One CTA per token; a work item = 8 elements of each half of one head (two 16-byte accesses).
template <typename T>
__global__ void __launch_bounds__(256)
rope_reshape_and_cache_kernel(T *__restrict__ q, T *__restrict__ k, const T *__restrict__ v,
                              T *__restrict__ k_cache, T *__restrict__ v_cache,
                              const int64_t *__restrict__ positions, const T *__restrict__ cos_t,
                              const T *__restrict__ sin_t, const int64_t *__restrict__ slot_mapping,
                              int num_heads, int num_kv_heads, int head_size, int64_t q_stride,
                              int64_t k_stride, int64_t v_stride, int64_t rot_stride, int64_t block_stride,
                              int64_t block_size) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t token = blockIdx.x;
    const int half = head_size / 2, cph = half / 8;  // 16-byte chunks per half head
    const int64_t pos = positions[token];
    const int64_t slot = slot_mapping ? slot_mapping[token] : -1;
    const bool to_cache = slot >= 0;  // slot < 0: padded token (cache_manager.cu:151-155)
    const int64_t cache_row = to_cache ? (slot / block_size) * block_stride + (slot % block_size) * num_kv_heads * head_size : 0;
    const T *cr = cos_t + pos * rot_stride, *sr = sin_t + pos * rot_stride;
    const int q_items = num_heads * cph, k_items = num_kv_heads * cph, v_items = num_kv_heads * head_size / 8;
    for (int it = threadIdx.x; it < q_items + k_items + v_items; it += blockDim.x) {
        if (it < q_items + k_items) {
            const bool is_k = it >= q_items;
            const int j = is_k ? it - q_items : it;
            const int h = j / cph, c = j % cph;
            T *x = (is_k ? k + token * k_stride : q + token * q_stride) + h * head_size + c * 8;
            const uint4 a = *reinterpret_cast<const uint4 *>(x), b = *reinterpret_cast<const uint4 *>(x + half);
            const uint4 cw = *reinterpret_cast<const uint4 *>(cr + c * 8), sw = *reinterpret_cast<const uint4 *>(sr + c * 8);
            const T *x1 = reinterpret_cast<const T *>(&a), *x2 = reinterpret_cast<const T *>(&b);
            const T *cc = reinterpret_cast<const T *>(&cw), *ss = reinterpret_cast<const T *>(&sw);
            uint4 o1, o2;
            T *y1 = reinterpret_cast<T *>(&o1), *y2 = reinterpret_cast<T *>(&o2);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float f1 = static_cast<float>(x1[e]), f2 = static_cast<float>(x2[e]);
                const float fc = static_cast<float>(cc[e]), fs = static_cast<float>(ss[e]);
                // explicit contraction: the same bits as the fused GEMM epilogue (gemm_skinny.cu, kEpiRope)
                y1[e] = static_cast<T>(__fmaf_rn(f1, fc, -__fmul_rn(f2, fs)));
                y2[e] = static_cast<T>(__fmaf_rn(f1, fs, __fmul_rn(f2, fc)));
            }
            *reinterpret_cast<uint4 *>(x) = o1;
            *reinterpret_cast<uint4 *>(x + half) = o2;
            if (is_k && to_cache) {
                T *dst = k_cache + cache_row + h * head_size + c * 8;
                *reinterpret_cast<uint4 *>(dst) = o1;
                *reinterpret_cast<uint4 *>(dst + half) = o2;
            }
        } else if (to_cache) {
            const int j = it - q_items - k_items;
            *reinterpret_cast<uint4 *>(v_cache + cache_row + j * 8) =
                ld_nc_v4(v + token * v_stride + j * 8);
        }
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void copy_blocks_impl(void *key_cache_ptrs, void *value_cache_ptrs, const void *block_mapping,
                      int64_t num_layers, int64_t num_pairs, int64_t numel_per_block,
                      cudaStream_t stream) {
    clear_status();
    if (num_layers <= 0 || num_pairs <= 0 || numel_per_block <= 0) return;
    mark_cache_dirty(stream);   // whole pages are rewritten: the next decode on this stream must not run ahead
    if (num_pairs > 65535) {
        // gridDim.y limit; issue in slabs.
        const int64_t *bm = static_cast<const int64_t *>(block_mapping);
        for (int64_t p0 = 0; p0 < num_pairs; p0 += 65535) {
            copy_blocks_impl(key_cache_ptrs, value_cache_ptrs, bm + 2 * p0, num_layers,
                             std::min<int64_t>(65535, num_pairs - p0), numel_per_block, stream);
        }
        return;
    }
    const int64_t block_bytes = numel_per_block * 2;
    dim3 grid(static_cast<unsigned>(num_layers), static_cast<unsigned>(num_pairs));
    // The base pointers live in device memory; blocks of a 16-bit cache allocated by any CUDA
    // allocator are 256-byte aligned, so 16-byte alignment reduces to block_bytes % 16 == 0.  A
    // misaligned base would fault the bulk copy loudly rather than corrupt data.
    if (block_bytes % 16 == 0) {
        static thread_local int attr_dev = -1;
        const DeviceInfo &di = current_device_info();
        if (attr_dev != di.ordinal) {
            cudaFuncSetAttribute(copy_blocks_bulk_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kRing * kPieceBytes);
            cudaFuncSetAttribute(swap_blocks_bulk_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kRing * kPieceBytes);
            attr_dev = di.ordinal;
        }
        copy_blocks_bulk_kernel<<<grid, 32, kRing * kPieceBytes, stream>>>(
            static_cast<const int64_t *>(key_cache_ptrs),
            static_cast<const int64_t *>(value_cache_ptrs),
            static_cast<const int64_t *>(block_mapping), block_bytes);
    } else {
        const int threads = static_cast<int>(std::min<int64_t>(1024, numel_per_block));
        copy_blocks_scalar_kernel<<<grid, threads, 0, stream>>>(
            static_cast<const int64_t *>(key_cache_ptrs),
            static_cast<const int64_t *>(value_cache_ptrs),
            static_cast<const int64_t *>(block_mapping), numel_per_block);
    }
    note_launch();
    check_launch("copy_blocks");
}

}  // namespace
}  // namespace atoma

using namespace atoma;

extern "C" void copy_blocks_f16(void *key_cache_ptrs, void *value_cache_ptrs,
                                const void *block_mapping, int64_t num_layers, int64_t num_pairs,
                                int64_t numel_per_block, atoma_stream_t stream) {
    copy_blocks_impl(key_cache_ptrs, value_cache_ptrs, block_mapping, num_layers, num_pairs,
                     numel_per_block, static_cast<cudaStream_t>(stream));
}

extern "C" void copy_blocks_bf16(void *key_cache_ptrs, void *value_cache_ptrs,
                                 const void *block_mapping, int64_t num_layers, int64_t num_pairs,
                                 int64_t numel_per_block, atoma_stream_t stream) {
    copy_blocks_impl(key_cache_ptrs, value_cache_ptrs, block_mapping, num_layers, num_pairs,
                     numel_per_block, static_cast<cudaStream_t>(stream));
}

extern "C" void reshape_and_cache_flash(void *key, void *value, void *key_cache, void *value_cache,
                                        int64_t *slot_mapping, int64_t block_stride,
                                        int64_t num_tokens, int64_t num_heads, int64_t head_size,
                                        int64_t block_size, int64_t key_stride,
                                        int64_t value_stride, uint32_t dtype,
                                        atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype > 1) {  // the reference silently does nothing for other dtypes (cache_manager.cu:237-241)
        set_status(kErrUnsupported, "reshape_and_cache_flash: dtype %u (0 = f16, 1 = bf16)", dtype);
        return;
    }
    if (num_tokens <= 0 || num_heads <= 0 || head_size <= 0) return;
    const int64_t n = num_heads * head_size;
    const int64_t row_bytes = n * 2;
    const bool vec_ok = row_bytes % 16 == 0 && aligned16(key) && aligned16(value) &&
                        aligned16(key_cache) && aligned16(value_cache) &&
                        (key_stride * 2) % 16 == 0 && (value_stride * 2) % 16 == 0 &&
                        (block_stride * 2) % 16 == 0;
    if (vec_ok) {
        const int chunks = static_cast<int>(row_bytes / 16);
        const int64_t total = num_tokens * chunks;
        const DeviceInfo &di = current_device_info();
        const int64_t want = (total + kRcThreads * kRcUnroll - 1) / (kRcThreads * kRcUnroll);
        const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, di.sm_count * 8)));
        cudaError_t e = launch_ex(reshape_and_cache_flash_vec_kernel, dim3(grid), dim3(kRcThreads), 0, stream,
                                  static_cast<const char *>(key), static_cast<const char *>(value),
                                  static_cast<char *>(key_cache), static_cast<char *>(value_cache),
                                  static_cast<const int64_t *>(slot_mapping), num_tokens, chunks, key_stride * 2,
                                  value_stride * 2, block_stride * 2, row_bytes, block_size);
        if (e != cudaSuccess) set_status(static_cast<int>(e), "reshape_and_cache_flash: %s", cudaGetErrorString(e));
    } else {
        const int threads = static_cast<int>(std::min<int64_t>(n, 512));
        reshape_and_cache_flash_scalar_kernel<<<static_cast<unsigned>(num_tokens), threads, 0, stream>>>(
            static_cast<const uint16_t *>(key), static_cast<const uint16_t *>(value),
            static_cast<uint16_t *>(key_cache), static_cast<uint16_t *>(value_cache), slot_mapping,
            block_stride, key_stride, value_stride, n, block_size);
    }
    note_launch();
    check_launch("reshape_and_cache_flash");
}

extern "C" void atoma_rope_reshape_and_cache(void *q, void *k, void *v, void *key_cache, void *value_cache,
                                             const int64_t *positions, const void *cos, const void *sin,
                                             const int64_t *slot_mapping, int64_t num_tokens,
                                             int64_t num_heads, int64_t num_kv_heads, int64_t head_size,
                                             int64_t q_stride, int64_t k_stride, int64_t v_stride,
                                             int64_t rot_stride, int64_t block_stride, int64_t block_size,
                                             uint32_t dtype, atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype > 1) {
        set_status(kErrUnsupported, "atoma_rope_reshape_and_cache: dtype %u (0 = f16, 1 = bf16)", dtype);
        return;
    }
    if (num_tokens <= 0) return;
    const bool ok = head_size > 0 && head_size % 16 == 0 && aligned16(q) && aligned16(k) && aligned16(v) &&
                    aligned16(key_cache) && aligned16(value_cache) && aligned16(cos) && aligned16(sin) &&
                    q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && rot_stride % 8 == 0 &&
                    block_stride % 8 == 0 && num_heads > 0 && num_kv_heads > 0 && positions != nullptr;
    if (!ok) {
        set_status(kErrUnsupported,
                   "atoma_rope_reshape_and_cache: head_size must be a multiple of 16 and all pointers / strides "
                   "16-byte aligned");
        return;
    }
    cudaError_t e;
    if (dtype == 1)
        e = launch_ex(rope_reshape_and_cache_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(num_tokens)), dim3(256), 0,
                      stream, static_cast<__nv_bfloat16 *>(q), static_cast<__nv_bfloat16 *>(k),
                      static_cast<const __nv_bfloat16 *>(v), static_cast<__nv_bfloat16 *>(key_cache),
                      static_cast<__nv_bfloat16 *>(value_cache), positions, static_cast<const __nv_bfloat16 *>(cos),
                      static_cast<const __nv_bfloat16 *>(sin), slot_mapping, static_cast<int>(num_heads),
                      static_cast<int>(num_kv_heads), static_cast<int>(head_size), q_stride, k_stride, v_stride,
                      rot_stride, block_stride, block_size);
    else
        e = launch_ex(rope_reshape_and_cache_kernel<__half>, dim3(static_cast<unsigned>(num_tokens)), dim3(256), 0, stream,
                      static_cast<__half *>(q), static_cast<__half *>(k), static_cast<const __half *>(v),
                      static_cast<__half *>(key_cache), static_cast<__half *>(value_cache), positions,
                      static_cast<const __half *>(cos), static_cast<const __half *>(sin), slot_mapping,
                      static_cast<int>(num_heads), static_cast<int>(num_kv_heads), static_cast<int>(head_size), q_stride,
                      k_stride, v_stride, rot_stride, block_stride, block_size);
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "atoma_rope_reshape_and_cache: %s", cudaGetErrorString(e));
        return;
    }
    check_launch("atoma_rope_reshape_and_cache");
}

extern "C" void swap_blocks_batched(void *dst_base, const void *src_base,
                                    const int64_t *src_blocks, const int64_t *dst_blocks,
                                    int64_t num_pairs, int64_t block_bytes, int kind,
                                    atoma_stream_t stream_) {
    clear_status();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_pairs <= 0 || block_bytes <= 0) return;
    if (kind == 0 || kind == 1) mark_cache_dirty(stream);   // whole pages of a device cache are rewritten
    if (kind == 0) {
        // Device -> device: upload the mapping, one launch for all pairs.
        int64_t *dmap = static_cast<int64_t *>(
            get_workspace(stream, 2 * num_pairs * sizeof(int64_t), kWorkspaceSwapMap));
        if (!dmap) return;
        cudaMemcpyAsync(dmap, src_blocks, num_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(dmap + num_pairs, dst_blocks, num_pairs * sizeof(int64_t),
                        cudaMemcpyHostToDevice, stream);
        if (block_bytes % 16 == 0 && aligned16(dst_base) && aligned16(src_base)) {
            static thread_local int attr_dev = -1;
            const DeviceInfo &di = current_device_info();
            if (attr_dev != di.ordinal) {
                cudaFuncSetAttribute(swap_blocks_bulk_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kRing * kPieceBytes);
                attr_dev = di.ordinal;
            }
            swap_blocks_bulk_kernel<<<static_cast<unsigned>(num_pairs), 32, kRing * kPieceBytes, stream>>>(
                static_cast<char *>(dst_base), static_cast<const char *>(src_base), dmap,
                dmap + num_pairs, block_bytes);
        } else {
            swap_blocks_scalar_kernel<<<static_cast<unsigned>(num_pairs), 256, 0, stream>>>(
                static_cast<char *>(dst_base), static_cast<const char *>(src_base), dmap,
                dmap + num_pairs, block_bytes);
        }
        note_launch();
        check_launch("swap_blocks_batched");
        return;
    }
    if (kind != 1 && kind != 2) {
        set_status(kErrInvalidArg, "swap_blocks_batched: kind %d (0 d2d, 1 h2d, 2 d2h)", kind);
        return;
    }
    // Host <-> device.  Pinned (or registered) host memory is mapped into the device address space:
    // the same one-launch bulk mover then reads / writes it over the host link, scattered blocks
    // included, instead of one memcpy per block (the reference: csrc/src/ops.rs:158-216).
    {
        const void *hptr = kind == 1 ? src_base : dst_base;
        cudaPointerAttributes pa;
        std::memset(&pa, 0, sizeof(pa));
        const bool mapped = cudaPointerGetAttributes(&pa, hptr) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                            pa.devicePointer != nullptr;
        cudaGetLastError();  // an unregistered pointer is not an error here
        if (mapped && block_bytes % 16 == 0 && aligned16(dst_base) && aligned16(src_base)) {
            int64_t *dmap = static_cast<int64_t *>(
                get_workspace(stream, 2 * num_pairs * sizeof(int64_t), kWorkspaceSwapMap));
            if (!dmap) return;
            cudaMemcpyAsync(dmap, src_blocks, num_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(dmap + num_pairs, dst_blocks, num_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
            static thread_local int attr_dev = -1;
            const DeviceInfo &di = current_device_info();
            if (attr_dev != di.ordinal) {
                cudaFuncSetAttribute(swap_blocks_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kRing * kPieceBytes);
                attr_dev = di.ordinal;
            }
            char *d = static_cast<char *>(kind == 2 ? pa.devicePointer : dst_base);
            const char *sp = static_cast<const char *>(kind == 1 ? pa.devicePointer : src_base);
            swap_blocks_bulk_kernel<<<static_cast<unsigned>(num_pairs), 32, kRing * kPieceBytes, stream>>>(
                d, sp, dmap, dmap + num_pairs, block_bytes);
            note_launch();
            check_launch("swap_blocks_batched (mapped host memory)");
            return;
        }
    }
    // Pageable host memory: merge runs of consecutive (src, dst) blocks into single async copies.
    std::vector<std::pair<int64_t, int64_t>> pairs(num_pairs);
    for (int64_t i = 0; i < num_pairs; ++i) pairs[i] = {src_blocks[i], dst_blocks[i]};
    std::sort(pairs.begin(), pairs.end());
    const cudaMemcpyKind mk = kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    int64_t i = 0;
    while (i < num_pairs) {
        int64_t j = i + 1;
        while (j < num_pairs && pairs[j].first == pairs[j - 1].first + 1 &&
               pairs[j].second == pairs[j - 1].second + 1)
            ++j;
        cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dst_base) + pairs[i].second * block_bytes,
                                        static_cast<const char *>(src_base) + pairs[i].first * block_bytes,
                                        static_cast<size_t>((j - i) * block_bytes), mk, stream);
        if (e != cudaSuccess) {
            set_status(static_cast<int>(e), "swap_blocks_batched: cudaMemcpyAsync: %s",
                       cudaGetErrorString(e));
            return;
        }
        i = j;
    }
}
