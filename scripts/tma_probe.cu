// Diagnostics (not part of the library): how fast can one SM stream a paged KV cache with TMA tensor
// loads as a function of the box size and of the number of issuing warps?  Mirrors the decode
// kernel's producer: 4-D map {d, token-in-page, kv head, page}, box {64, 16, hkt, 1}, 128B swizzle,
// 3 stages of 64 KiB, random pages.  Mode "5d" merges the two 64-wide halves of D=128 into one op
// (5-D map {64, 2, 16, hk, page}).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_probe scripts/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma5(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

constexpr int kStages = 3, kStageBytes = 64 * 1024;

// P producer warps (warps 0..P-1) + 1 consumer warp (warp P).
__global__ void __launch_bounds__(160, 1)
probe(const __grid_constant__ CUtensorMap map, int five_d, int hkt, int hg_count, int nb, int stages_per_cta, int P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], P); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int box_bytes = (five_d ? 2 : 1) * 128 * 16 * hkt;  // [64 d](x2) x 16 tokens x hkt heads
    const int ops = kStageBytes / box_bytes;                  // ops per stage
    int stage = 0; uint32_t phase = 0;
    if (warp < P) {
        if (lane == 0) {
            uint32_t rng = blockIdx.x * 2654435761u + 12345u;
            for (int s = 0; s < stages_per_cta; ++s) {
                mbar_wait(&empty[stage], phase ^ 1);
                int mine = 0;
                for (int o = warp; o < ops; o += P) ++mine;
                mbar_arrive_expect_tx(&full[stage], mine * box_bytes);
                for (int o = 0; o < ops; ++o) {
                    // every warp walks the same sequence; in 4-D mode ops come in pairs (the two halves of a page)
                    if (five_d || !(o & 1)) rng = rng * 1664525u + 1013904223u;
                    if (o % P != warp) continue;
                    const int page = (rng >> 8) % nb, hg = (rng >> 4) % hg_count;
                    unsigned char *dst = tiles + stage * kStageBytes + o * box_bytes;
                    if (five_d) tma5(dst, &map, &full[stage], 0, 0, 0, hg * hkt, page);
                    else tma4(dst, &map, &full[stage], (o & 1) * 64, 0, hg * hkt, page);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == P) {
        for (int s = 0; s < stages_per_cta; ++s) {
            mbar_wait(&full[stage], phase);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    const int D = 128, page = 16;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes + 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("mode hk hkt box_KiB warps stages/cta  us   GB/s  GB/s/SM ops/us/SM\n");
    for (int hk : {8, 4, 2, 1}) {
        const int64_t nb = (int64_t)(1u << 30) / (page * hk * D * 2);   // 1 GiB of K
        void *buf;
        CK(cudaMalloc(&buf, (size_t)nb * page * hk * D * 2));
        CK(cudaMemset(buf, 1, (size_t)nb * page * hk * D * 2));
        for (int five = 0; five < 2; ++five) {
            const int hkt = hk;
            CUtensorMap m;
            CUresult r;
            if (!five) {
                const cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)page, (cuuint64_t)hk, (cuuint64_t)nb};
                const cuuint64_t strides[3] = {(cuuint64_t)hk * D * 2, (cuuint64_t)D * 2, (cuuint64_t)page * hk * D * 2};
                const cuuint32_t box[4] = {64, 16, (cuuint32_t)hkt, 1}, es[4] = {1, 1, 1, 1};
                r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else {
                const cuuint64_t dims[5] = {64, 2, (cuuint64_t)page, (cuuint64_t)hk, (cuuint64_t)nb};
                const cuuint64_t strides[4] = {128, (cuuint64_t)hk * D * 2, (cuuint64_t)D * 2, (cuuint64_t)page * hk * D * 2};
                const cuuint32_t box[5] = {64, 2, 16, (cuuint32_t)hkt, 1}, es[5] = {1, 1, 1, 1, 1};
                r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            if (r != CUDA_SUCCESS) { printf("encode failed %d (five=%d hk=%d)\n", (int)r, five, hk); continue; }
            for (int P : {1, 2, 4}) {
                for (int spc : {7, 28, 56}) {
                    for (int it = 0; it < 3; ++it) probe<<<148, 160, kStages * kStageBytes + 1024>>>(m, five, hkt, hk / hkt, (int)nb, spc, P);
                    CK(cudaDeviceSynchronize());
                    const int reps = 10;
                    CK(cudaEventRecord(e0));
                    for (int it = 0; it < reps; ++it) probe<<<148, 160, kStages * kStageBytes + 1024>>>(m, five, hkt, hk / hkt, (int)nb, spc, P);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double us = ms * 1e3 / reps, bytes = 148.0 * spc * kStageBytes;
                    const int box_bytes = (five ? 2 : 1) * 128 * 16 * hkt;
                    printf("%s   %d  %d   %5.1f     %d     %3d      %7.2f %7.1f %6.1f %6.2f\n", five ? "5d" : "4d", hk, hkt, box_bytes / 1024.0, P, spc, us,
                           bytes / us / 1e3, bytes / us / 1e3 / 148, (double)spc * (kStageBytes / box_bytes) / us);
                }
            }
        }
        CK(cudaFree(buf));
    }
    return 0;
}
