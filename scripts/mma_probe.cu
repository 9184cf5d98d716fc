// Standalone probe of the tcgen05.mma issue path (kind::f16, SS operands in 128B-swizzled shared
// memory, no memory traffic: operands are whatever shared memory holds).  What does one MMA cost as a
// function of N, and what does a commit group cost?  All loops run warp-converged with the MMAs under
// elect.sync (the structure gemm_skinny.cu uses).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I atoma_infer_b200/csrc -I include \
//        scripts/mma_probe.cu -o /tmp/mma_probe && /tmp/mma_probe
#include "common.cuh"

#include <cstdio>
#include <cstdlib>

using namespace atoma;

// mode 0: `groups` groups of `per` MMAs, one commit per group, ONE wait at the very end
// mode 1: same, but the thread waits for every commit before issuing the next group
// mode 2: like 0, and before each group the warp waits on an mbarrier that another warp arrives on
//         (a producer that is always ahead): the cost of the wait/wake path itself
// mode 3: like 0 but the groups alternate between two accumulators
template <int M>
__global__ void __launch_bounds__(128, 1) probe(int n, int per, int groups, int mode, long long *out) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar[8], fullb[8];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) {
            mbar_init(&bar[i], 1);
            mbar_init(&fullb[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_f16(M, n, true);
        const uint32_t base = smem_u32(ring);
        const uint64_t da = umma_desc_k_sw128(base), db = umma_desc_k_sw128(base + 64 * 1024);
        const long long t0 = clock64();
        int slot = 0;
        uint32_t phase = 0;
        for (int g = 0; g < groups; ++g) {
            if (mode == 2) {
                mbar_wait(&fullb[slot], phase);
                tc_fence_after_sync();
            }
            if (elect_one_sync()) {
                const uint32_t d = tmem_base + ((mode == 3 && (g & 1)) ? 256u : 0u);
                for (int j = 0; j < per; ++j) umma_f16_ss(d, da + 2 * (j & 3), db + 2 * (j & 3), idesc, (g | j) != 0 ? 1u : 0u);
                umma_commit(&bar[slot]);
            }
            __syncwarp();
            if (mode == 1) mbar_wait(&bar[slot], phase);
            if (++slot == 8) {
                slot = 0;
                phase ^= 1;
            }
        }
        const long long t_issue = clock64();
        // wait for the last commit
        const int last = (groups - 1) & 7;
        mbar_wait(&bar[last], ((groups - 1) >> 3) & 1);
        const long long t1 = clock64();
        if (tid == 0) {
            out[blockIdx.x * 2] = t_issue - t0;
            out[blockIdx.x * 2 + 1] = t1 - t0;
        }
    } else if (warp == 2 && mode == 2) {
        // producer: arrives on fullb[slot] as soon as the MMA warp released the slot 8 groups ago
        int slot = 0;
        uint32_t phase = 0;
        for (int g = 0; g < groups; ++g) {
            mbar_wait(&bar[slot], phase ^ 1);
            if (elect_one_sync()) mbar_arrive(&fullb[slot]);
            __syncwarp();
            if (++slot == 8) {
                slot = 0;
                phase ^= 1;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

int main() {
    long long *out;
    cudaMalloc(&out, 1024 * sizeof(long long));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("# tcgen05.mma kind::f16 bf16 K=16 SS; 148 CTAs; cycles per MMA: issue-side | until the last commit completes\n");
    printf("#  M   N per grp mode   issue/mma  total/mma   total/group\n");
    const int ns[] = {32, 48, 64, 128, 208, 256};
    for (int m : {64, 128})
        for (int n : ns)
            for (int per : {4, 16, 32})
                for (int mode : {0, 1, 2, 3}) {
                    const int groups = 1024 / per;
                    if (m == 64)
                        probe<64><<<148, 128, smem>>>(n, per, groups, mode, out);
                    else
                        probe<128><<<148, 128, smem>>>(n, per, groups, mode, out);
                    long long h[1024];
                    cudaError_t e = cudaMemcpy(h, out, 2 * 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) {
                        printf("error %s\n", cudaGetErrorString(e));
                        return 1;
                    }
                    printf("%4d %3d %3d %3d %d   %8.1f %8.1f %10.1f\n", m, n, per, groups, mode, h[148] / 1024.0, h[149] / 1024.0,
                           static_cast<double>(h[149]) / groups);
                }
    return 0;
}
