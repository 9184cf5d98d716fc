// Shared device/host helpers for the sm_100a kernels: PTX wrappers (mbarrier, TMA, ldmatrix,
// mma.sync), status plumbing, per-device attribute cache.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace atoma {

// ---------------------------------------------------------------------------------------------
// Host-side status (thread local; the C ABI returns void like the reference's).
// ---------------------------------------------------------------------------------------------
enum : int {
    kOk = 0,
    kErrUnsupported = -1,
    kErrInvalidArg = -2,
    kErrDriver = -3,
};

void set_status(int code, const char *fmt, ...);
void clear_status();
void note_launch(int n = 1);
void set_mha_path(const char *name);
void note_decode_mode(int mode);   // bit 0: early KV streaming, bit 1: waited before releasing dependents
bool check_launch(const char *what);  // cudaGetLastError -> status; true if ok

struct DeviceInfo {
    int ordinal = -1;
    int sm_count = 0;
    int cc_major = 0;
    int cc_minor = 0;
    int max_smem_optin = 0;
};
// Cached per device (the reference re-queries cudaGetDeviceProperties on every call:
// csrc/src/lib.rs:2218-2221).
const DeviceInfo &current_device_info();

// Library-owned scratch, one per (device, stream, tag); grows monotonically, zero-initialised
// when (re)allocated.  Used for the decode kernel's split partials + arrival counters and for the
// device copy of a swap mapping.
enum : int { kWorkspaceDecode = 0, kWorkspaceSwapMap = 1, kWorkspaceGemm = 2, kWorkspaceGemmCounters = 3, kWorkspaceSkinnyTail = 4, kWorkspaceTags = 5 };
void *get_workspace(cudaStream_t stream, size_t bytes, int tag);

// Launch mode (process-wide, atoma_set_launch_mode): programmatic dependent launch between this
// library's kernels, "weights are immutable" (the GEMM may prefetch W before its grid dependency
// resolves) and "only the final page of a sequence is being written" (the decode kernel may stream
// the older pages before its grid dependency resolves).
enum : int { kLaunchPdl = 1, kLaunchConstWeights = 2, kLaunchStableCache = 4 };
int launch_flags();
void set_next_decode_stable_cache(int v);   // per-call form of kLaunchStableCache (thread-local, one-shot)
int take_next_decode_stable_cache();
// "What the NEXT kernel of the stream will stream first" (the decode-step executor knows: the next GEMM's
// weights).  One-shot, thread-local, consumed by the next skinny-GEMM / paged-decode launch on this thread: its
// producer warps issue L2 prefetches of [ptr, ptr + bytes) once their own last load is in flight, so that HBM
// keeps streaming through the kernel's tail, the kernel boundary and the next kernel's ramp-up.
struct NextPrefetch {
    const void *ptr;
    long long bytes;
};
// Step timeline (diagnostics, atoma_debug_step_trace): while a trace buffer is set, every kernel launched through
// the library gets its own slice of it — kStepTraceU64 words = [cta][8] globaltimer stamps — in launch order.
constexpr int kStepTraceCtas = 160;
constexpr int kStepTraceU64 = kStepTraceCtas * 8;
unsigned long long *next_step_trace_slot(const char *kernel_name);   // null when tracing is off or the buffer is full
void set_next_kernel_prefetch(const void *ptr, long long bytes);
void set_next_kernel_prefetch_off(int off);   // diagnostics: ignore the hints (A/B measurements)
NextPrefetch take_next_kernel_prefetch();
void mark_cache_dirty(cudaStream_t stream);   // a cache-maintenance op was enqueued on this stream
bool take_cache_dirty(cudaStream_t stream);   // ... since the last paged decode launch on it?

#if defined(__CUDACC__)
// Kernel launch with the programmatic-stream-serialization attribute when PDL is enabled: the
// kernel may start while its predecessor in the stream drains; it must execute pdl_wait() before
// touching anything the predecessor produces (every kernel launched through here does).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                      Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (launch_flags() & kLaunchPdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch (no-ops when the kernel was launched without the attribute) ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// L2 prefetch of `bytes` (multiple of 16) at a 16-byte aligned global address: no destination, no completion.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// [base, base + bytes) is prefetched in 16 KiB pieces; pieces [i0, i1) by the caller.
constexpr long long kL2PrefetchPiece = 16 * 1024;
__device__ __forceinline__ long long l2_prefetch_pieces(long long bytes) {
    return (bytes + kL2PrefetchPiece - 1) / kL2PrefetchPiece;
}
__device__ __forceinline__ void l2_prefetch_range(const unsigned char *base, long long bytes, long long i0, long long i1) {
    for (long long i = i0; i < i1; ++i) {
        const long long off = i * kL2PrefetchPiece;
        const long long n = min(kL2PrefetchPiece, bytes - off) & ~15ll;
        if (n > 0) l2_prefetch_bulk(base + off, static_cast<uint32_t>(n));
    }
}

// Publish: every write that happens-before this (the caller's own and, through a preceding CTA / warp barrier, its
// peers') becomes visible at GPU scope before the counter moves.  Lighter than __threadfence() + atomicAdd
// (MEMBAR.SC.GPU + L1 invalidate on the publishing side, where nothing is about to be read).
__device__ __forceinline__ void red_release_gpu_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One lane of the (converged) warp: true in exactly one thread.  Code under `if (elect_one_sync())`
// keeps warp-uniform operands in uniform registers, which is what UTMALDG / UTCHMMA consume.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(bytes)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA (cp.async.bulk.tensor), 4-D tiled load with an L2 cache-policy hint ----
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, int c3, int c4, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

// ---- thread-block clusters: rank, split-phase barrier, distributed shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// Every thread of every CTA of the cluster executes arrive then wait, once per phase, warp-converged.
__device__ __forceinline__ void cluster_arrive_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

// ---- cp.async (16-byte, zero-fill when src_bytes < 16) ----
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- named barriers ----
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- ldmatrix / mma.sync (decode + generic attention: M=16 tiles, fp32 accumulate) ----
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

template <typename T>
struct Mma16816;
template <>
struct Mma16816<__nv_bfloat16> {
    __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <>
struct Mma16816<__half> {
    __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

// Pack two fp32 into one 32-bit register of 16-bit floats (lo = first argument).
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- tcgen05 (5th-gen tensor cores, accumulators in TMEM) ----
// One warp allocates `ncols` (power of two >= 32) TMEM columns; the base address lands in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with the
// 128-byte swizzle (what a TMA box {64 x 16-bit, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes):
// 8-row swizzle atoms of 1024 bytes stacked along M/N (stride byte offset 1024), descriptor
// version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  Bits: [0,14) start address >> 4,
// [16,30) leading byte offset >> 4 (unused for swizzled K-major), [32,46) stride byte offset >> 4,
// [46,48) version, [61,64) layout type.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor, kind::f16: D = fp32, A/B = bf16 (1) or f16 (0), both K-major, dense.
// Bits: [4,6) D format, [7,10) A format, [10,13) B format, 15/16 A/B major, [17,23) N >> 3,
// [24,29) M >> 4.
__device__ __host__ __forceinline__ uint32_t umma_idesc_f16(int m, int n, bool is_bf16,
                                                            bool b_mn_major = false) {
    const uint32_t fmt = is_bf16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// Shared-memory descriptor for an MN-major operand (the MMA's N index is contiguous in memory, its
// K index runs over rows): rows of 128 bytes = 64 N-elements, 128-byte swizzle, 8-row groups of
// 1024 bytes along K (stride byte offset), further 64-element blocks of N `lbo_bytes` apart
// (leading byte offset).  This is a [K rows][64] TMA box; V of attention is consumed this way.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is read from tensor memory (16-bit elements packed two per
// 32-bit column, lane = row).
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all tcgen05 ops issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread = lane of its warp's
// TMEM quadrant).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16-byte global accesses with streaming hints (data touched once).
__device__ __forceinline__ uint4 ld_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

#endif  // __CUDACC__

}  // namespace atoma
