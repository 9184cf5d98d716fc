// One-shot all-reduce(sum) over NVLink peer memory for the tensor-parallel path.
//
// Replaces, for the [T, hidden] 16-bit payload that follows o_proj (and down_proj), the reference's
// `comm.all_reduce(src, &mut dst, &ReduceOp::Sum)` through cudarc/NCCL (models/src/multi_gpu.rs:
// 141-179).  The payload is 256 KiB at batch 32: a latency-bound collective (NVLink 5 moves it in
// 0.3 us) for which ncclAllReduce measures 21.3 us on 2 x B200 inside a CUDA graph
// (scripts/nccl_diag.py).  A flag handshake followed by remote loads (a first version of this file)
// costs three NVLink flights, 11.7 us.  This kernel needs ONE flight: data and flag travel in the
// same store.
//
// Every rank keeps one device allocation that all peers map (cudaIpcOpenMemHandle):
//
//     [ header: epoch, arrival counter, error word | staging slot 0 | slot 1 | recv[slot][source rank] ]
//
// The producer kernel (o_proj GEMM) writes the rank's addend into a staging slot.  Then, per 16-byte
// vector i (one thread):
//   1. PUSH: the vector's four 32-bit words are stored into recv[slot][rank] of EVERY peer as two
//      16-byte stores {w0, e, w1, e}, {w2, e, w3, e} (two planes, so that a warp's store covers 512
//      contiguous bytes): each 8-byte unit carries its own flag e, the
//      launch's epoch (device-resident counter, so the launch replays inside a CUDA graph; flags of
//      earlier epochs never equal e).  No fence, no separate flag write.
//   2. POLL the local recv[slot][r][i] of every peer r until both units show e (volatile loads of
//      local memory; 8-byte units are written atomically).
//   3. Add the W addends in fp32 in RANK ORDER (bit-identical on every rank, run to run; NCCL's ring
//      rounds to the 16-bit type at every hop) and write the local output.
// Receive buffers alternate between consecutive all-reduces — by the parity of the epoch, decided on the
// DEVICE, so a captured graph with an odd number of all-reduces, or graphs replayed in any order, keep
// alternating: recv[parity] is overwritten by a peer's all-reduce k+2 only after that peer finished k+1,
// which needed this rank's push of k+1, issued after this rank's kernel k completed — so nothing is
// overwritten before it was read.  (The staging slot the producer writes is the host's choice: it is local
// and stream-ordered.)  The fused send/receive pair of gemm_skinny.cu + llama_ops.cu shares the epoch
// counter and the receive buffers of this file.
#include "common.cuh"

#include "atoma_flash.h"

#include <algorithm>
#include <cstring>

namespace atoma {
namespace {

constexpr int kArMaxWorld = 8;
constexpr int kArMaxCtas = 128;
constexpr int kArThreads = 256;
constexpr size_t kArHeaderBytes = 4096;   // u32 [0] epoch, [1] CTA arrival counter, [2] error word

struct ArArgs {
    unsigned char *region[kArMaxWorld];  // every rank's region base (own rank: local pointer)
    uint32_t *state;                     // local header
    void *out;
    int64_t stage_offset;                // byte offset of the staging slot inside a region
    int64_t recv_offset;                 // byte offset of recv[0][0] inside a region
    int64_t recv_slot_stride;            // bytes between recv[0][..] and recv[1][..]; the receive slot of a launch
                                         // is the parity of its epoch (device side: graphs replay in any order)
    int64_t recv_stride;                 // bytes between recv[slot][r] and recv[slot][r + 1]
    int64_t nvec;                        // 16-byte vectors
    int rank, world;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_relaxed_sys_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ar_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

template <typename T>
__device__ __forceinline__ void add_vec(float (&acc)[8], const uint4 &v);
template <>
__device__ __forceinline__ void add_vec<__nv_bfloat16>(float (&acc)[8], const uint4 &v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <>
__device__ __forceinline__ void add_vec<__half>(float (&acc)[8], const uint4 &v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
    }
}
template <>
__device__ __forceinline__ void add_vec<float>(float (&acc)[8], const uint4 &v) {
    acc[0] += __uint_as_float(v.x);
    acc[1] += __uint_as_float(v.y);
    acc[2] += __uint_as_float(v.z);
    acc[3] += __uint_as_float(v.w);
}
template <typename T>
__device__ __forceinline__ uint4 pack_vec(const float (&acc)[8]) {
    uint4 o;
    o.x = pack2<T>(acc[0], acc[1]);
    o.y = pack2<T>(acc[2], acc[3]);
    o.z = pack2<T>(acc[4], acc[5]);
    o.w = pack2<T>(acc[6], acc[7]);
    return o;
}
template <>
__device__ __forceinline__ uint4 pack_vec<float>(const float (&acc)[8]) {
    return make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]),
                      __float_as_uint(acc[3]));
}

template <typename T, int W>
__global__ void __launch_bounds__(kArThreads) allreduce_p2p_kernel(const ArArgs a) {
    const int tid = threadIdx.x;
    // The staging slot was written by the kernel before this one in the stream.
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t e = a.state[0] + 1;   // this launch's epoch; advanced by the last CTA to finish
    if (a.state[2] != 0) return;          // a peer failed to arrive earlier: do not wait for it again
    const int64_t recv_offset = a.recv_offset + (e & 1u) * a.recv_slot_stride;
    bool failed = false;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kArThreads + tid; i < a.nvec;
         i += static_cast<int64_t>(gridDim.x) * kArThreads) {
        // ---- 1. push ----
        const uint4 mine = *reinterpret_cast<const uint4 *>(a.region[a.rank] + a.stage_offset + i * 16);
#pragma unroll
        for (int d = 1; d < W; ++d) {
            const int r = (a.rank + d) & (W - 1);   // start with the next rank: spreads the links
            // two planes of nvec 16-byte units: every warp-wide store covers 512 contiguous bytes
            unsigned char *dst = a.region[r] + recv_offset + a.rank * a.recv_stride + i * 16;
            st_relaxed_sys_v4(dst, mine.x, e, mine.y, e);
            st_relaxed_sys_v4(dst + a.nvec * 16, mine.z, e, mine.w, e);
        }
        // ---- 2. poll what the peers pushed, 3. add in rank order ----
        uint4 lo[W], hi[W];
        bool done[W];
#pragma unroll
        for (int r = 0; r < W; ++r) done[r] = (r == a.rank);
        const unsigned char *src = a.region[a.rank] + recv_offset + i * 16;
        const unsigned long long t0 = ar_timer_ns();
        for (int spin = 0;; ++spin) {
            bool all = true;
#pragma unroll
            for (int r = 0; r < W; ++r) {
                if (!done[r]) {
                    lo[r] = ld_volatile_v4(src + r * a.recv_stride);
                    hi[r] = ld_volatile_v4(src + r * a.recv_stride + a.nvec * 16);
                    done[r] = lo[r].y == e && lo[r].w == e && hi[r].y == e && hi[r].w == e;
                    all = all && done[r];
                }
            }
            if (all) break;
            if ((spin & 63) == 63 && ar_timer_ns() - t0 > a.timeout_ns) {
                failed = true;
                break;
            }
        }
        if (failed) break;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < W; ++r) {
            const uint4 v = (r == a.rank) ? mine : make_uint4(lo[r].x, lo[r].z, hi[r].x, hi[r].z);
            add_vec<T>(acc, v);
        }
        reinterpret_cast<uint4 *>(a.out)[i] = pack_vec<T>(acc);
    }
    if (failed) atomicExch(&a.state[2], 1u);   // error word: a peer never arrived
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&a.state[1], 1u) == gridDim.x - 1) {   // last CTA: next launch sees the next epoch
            a.state[1] = 0;
            a.state[0] = e;
        }
    }
}

// Two-shot variant for larger worlds: rank o owns the o-th W-th of the vectors.  Flight 1: every rank
// pushes its addend of a vector to the vector's owner; the owner adds the W addends in rank order,
// writes its own output and (flight 2) pushes the finished vector to every peer, which copies it
// out.  Each rank sends 2 (W-1)/W of the payload instead of (W-1) x: at 8 ranks a quarter of the
// one-shot traffic for one more NVLink flight.  Results stay bit-identical on every rank (each vector
// is summed once, by its owner).  recv[slot][r] lower half: flight-1 data from rank r; upper half of
// recv[slot][o]: flight-2 data of owner o.
template <typename T, int W>
__global__ void __launch_bounds__(kArThreads) allreduce_p2p_twoshot_kernel(const ArArgs a) {
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t e = a.state[0] + 1;
    if (a.state[2] != 0) return;
    const int64_t recv_offset = a.recv_offset + (e & 1u) * a.recv_slot_stride;
    const int64_t chunk = (a.nvec + W - 1) / W;          // vectors per owner
    const int64_t half = a.recv_stride / 2;              // byte offset of the flight-2 area
    bool failed = false;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kArThreads + tid; i < a.nvec;
         i += static_cast<int64_t>(gridDim.x) * kArThreads) {
        const int owner = static_cast<int>(i / chunk);
        const int64_t il = i - owner * chunk;             // index inside the owner's chunk
        const uint4 mine = *reinterpret_cast<const uint4 *>(a.region[a.rank] + a.stage_offset + i * 16);
        const unsigned long long t0 = ar_timer_ns();
        if (owner != a.rank) {
            // flight 1: my addend to the owner
            unsigned char *dst = a.region[owner] + recv_offset + a.rank * a.recv_stride + il * 16;
            st_relaxed_sys_v4(dst, mine.x, e, mine.y, e);
            st_relaxed_sys_v4(dst + chunk * 16, mine.z, e, mine.w, e);
            // flight 2: the finished vector from the owner
            const unsigned char *src = a.region[a.rank] + recv_offset + owner * a.recv_stride + half + il * 16;
            uint4 lo, hi;
            for (int spin = 0;; ++spin) {
                lo = ld_volatile_v4(src);
                hi = ld_volatile_v4(src + chunk * 16);
                if (lo.y == e && lo.w == e && hi.y == e && hi.w == e) break;
                if ((spin & 63) == 63 && ar_timer_ns() - t0 > a.timeout_ns) {
                    failed = true;
                    break;
                }
            }
            if (failed) break;
            reinterpret_cast<uint4 *>(a.out)[i] = make_uint4(lo.x, lo.z, hi.x, hi.z);
        } else {
            uint4 lo[W], hi[W];
            bool done[W];
#pragma unroll
            for (int r = 0; r < W; ++r) done[r] = (r == a.rank);
            const unsigned char *src = a.region[a.rank] + recv_offset + il * 16;
            for (int spin = 0;; ++spin) {
                bool all = true;
#pragma unroll
                for (int r = 0; r < W; ++r) {
                    if (!done[r]) {
                        lo[r] = ld_volatile_v4(src + r * a.recv_stride);
                        hi[r] = ld_volatile_v4(src + r * a.recv_stride + chunk * 16);
                        done[r] = lo[r].y == e && lo[r].w == e && hi[r].y == e && hi[r].w == e;
                        all = all && done[r];
                    }
                }
                if (all) break;
                if ((spin & 63) == 63 && ar_timer_ns() - t0 > a.timeout_ns) {
                    failed = true;
                    break;
                }
            }
            if (failed) break;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int r = 0; r < W; ++r) {
                const uint4 v = (r == a.rank) ? mine : make_uint4(lo[r].x, lo[r].z, hi[r].x, hi[r].z);
                add_vec<T>(acc, v);
            }
            const uint4 res = pack_vec<T>(acc);
            reinterpret_cast<uint4 *>(a.out)[i] = res;
#pragma unroll
            for (int d = 1; d < W; ++d) {
                const int r = (a.rank + d) & (W - 1);
                unsigned char *dst = a.region[r] + recv_offset + a.rank * a.recv_stride + half + il * 16;
                st_relaxed_sys_v4(dst, res.x, e, res.y, e);
                st_relaxed_sys_v4(dst + chunk * 16, res.z, e, res.w, e);
            }
        }
    }
    if (failed) atomicExch(&a.state[2], 1u);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&a.state[1], 1u) == gridDim.x - 1) {
            a.state[1] = 0;
            a.state[0] = e;
        }
    }
}

// 0 = library's choice, 1 = one-shot, 2 = two-shot (atoma_debug_allreduce_algo).  The choice is one-shot
// at every world size: measured at 4 ranks it is the faster one (6.6 vs 7.2 us per 256 KiB all-reduce,
// step 1.881 vs 1.895 ms); the 8-rank crossover has not been measured yet.
int g_ar_algo = 0;

template <typename T>
cudaError_t launch_ar(const ArArgs &a, int grid, cudaStream_t stream) {
    const bool two = g_ar_algo == 2;
    if (two) {
        switch (a.world) {
            case 2: return launch_ex(allreduce_p2p_twoshot_kernel<T, 2>, dim3(grid), dim3(kArThreads), 0, stream, a);
            case 4: return launch_ex(allreduce_p2p_twoshot_kernel<T, 4>, dim3(grid), dim3(kArThreads), 0, stream, a);
            case 8: return launch_ex(allreduce_p2p_twoshot_kernel<T, 8>, dim3(grid), dim3(kArThreads), 0, stream, a);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (a.world) {
        case 2: return launch_ex(allreduce_p2p_kernel<T, 2>, dim3(grid), dim3(kArThreads), 0, stream, a);
        case 4: return launch_ex(allreduce_p2p_kernel<T, 4>, dim3(grid), dim3(kArThreads), 0, stream, a);
        case 8: return launch_ex(allreduce_p2p_kernel<T, 8>, dim3(grid), dim3(kArThreads), 0, stream, a);
        default: return cudaErrorInvalidValue;
    }
}

inline int64_t slot_round(int64_t slot_bytes) { return (slot_bytes + 255) / 256 * 256; }

}  // namespace
}  // namespace atoma

using namespace atoma;

extern "C" void atoma_debug_allreduce_algo(int algo) { g_ar_algo = algo; }

extern "C" int64_t atoma_p2p_region_bytes(int64_t slot_bytes) {
    clear_status();
    const int64_t s = slot_round(slot_bytes);
    // header | 2 staging slots | recv[2 slots][8 source ranks], each 2 x slot (8-byte units = 4 data + 4 flag)
    return static_cast<int64_t>(kArHeaderBytes) + 2 * s + 2 * kArMaxWorld * 2 * s;
}

extern "C" int atoma_p2p_alloc(void **ptr, int64_t bytes) {
    clear_status();
    if (!ptr || bytes <= 0) {
        set_status(kErrInvalidArg, "atoma_p2p_alloc: bad arguments");
        return kErrInvalidArg;
    }
    cudaError_t e = cudaMalloc(ptr, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, static_cast<size_t>(bytes));
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "atoma_p2p_alloc: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return 0;
}

extern "C" int atoma_p2p_free(void *ptr) {
    clear_status();
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) set_status(static_cast<int>(e), "atoma_p2p_free: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
}

extern "C" int atoma_p2p_export(void *ptr, void *handle64) {
    clear_status();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int atoma_p2p_import(const void *handle64, void **peer_ptr) {
    clear_status();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return 0;
}

extern "C" int atoma_p2p_close(void *peer_ptr) {
    clear_status();
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    if (e != cudaSuccess) set_status(static_cast<int>(e), "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
}

extern "C" void *atoma_p2p_slot(void *region, int64_t slot_bytes, int slot) {
    return static_cast<unsigned char *>(region) + kArHeaderBytes + (slot & 1) * slot_round(slot_bytes);
}

extern "C" int atoma_p2p_error(void *region, atoma_stream_t stream) {
    uint32_t v = 0;
    const unsigned char *st = static_cast<const unsigned char *>(region);
    cudaError_t e = cudaMemcpyAsync(&v, st + 2 * sizeof(uint32_t), sizeof(v), cudaMemcpyDeviceToHost,
                                    static_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
    return static_cast<int>(v);
}

extern "C" void atoma_allreduce_sum_p2p(void *const *regions, int rank, int world, int64_t slot_bytes, int slot,
                                        void *out, int64_t count, uint32_t dtype, atoma_stream_t stream) {
    clear_status();
    if (!regions || !out || world < 2 || world > kArMaxWorld || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_status(kErrInvalidArg, "atoma_allreduce_sum_p2p: world must be 2, 4 or 8 and 0 <= rank < world (rank %d, world %d)",
                   rank, world);
        return;
    }
    const int64_t esz = dtype == 2 ? 4 : 2;
    if (dtype > 2 || count <= 0 || (count * esz) % 16 != 0 || count * esz > slot_bytes ||
        (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_status(kErrInvalidArg,
                   "atoma_allreduce_sum_p2p: dtype 0=f16 1=bf16 2=f32, count*size a multiple of 16 bytes within the slot, "
                   "16-byte aligned output (count %lld, dtype %u, slot %lld bytes)",
                   static_cast<long long>(count), dtype, static_cast<long long>(slot_bytes));
        return;
    }
    ArArgs a;
    for (int r = 0; r < kArMaxWorld; ++r) a.region[r] = static_cast<unsigned char *>(regions[r < world ? r : rank]);
    a.state = reinterpret_cast<uint32_t *>(regions[rank]);
    a.out = out;
    const int64_t s = slot_round(slot_bytes);
    a.stage_offset = static_cast<int64_t>(kArHeaderBytes) + (slot & 1) * s;
    a.recv_stride = 2 * s;
    a.recv_slot_stride = kArMaxWorld * a.recv_stride;
    a.recv_offset = static_cast<int64_t>(kArHeaderBytes) + 2 * s;
    a.nvec = count * esz / 16;
    a.rank = rank;
    a.world = world;
    // A peer that never arrives must not hang the GPU; ranks of one job may legitimately reach their
    // first all-reduce seconds apart (library initialisation), so the limit is generous and is paid
    // once: later launches return at once when the error word is set.
    a.timeout_ns = 30000000000ull;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(kArMaxCtas, (a.nvec + kArThreads - 1) / kArThreads)));
    cudaError_t e;
    if (dtype == 1) e = launch_ar<__nv_bfloat16>(a, grid, static_cast<cudaStream_t>(stream));
    else if (dtype == 0) e = launch_ar<__half>(a, grid, static_cast<cudaStream_t>(stream));
    else e = launch_ar<float>(a, grid, static_cast<cudaStream_t>(stream));
    note_launch();
    if (e != cudaSuccess) {
        set_status(static_cast<int>(e), "allreduce_p2p_kernel: %s", cudaGetErrorString(e));
        return;
    }
    check_launch("allreduce_p2p_kernel");
}
