// Dense bf16 GEMM y = x w^T (placeholder until the tcgen05 kernel lands in this file).
#include "common.cuh"
namespace atoma {
bool launch_gemm_bf16_tn(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t,
                         int64_t, int64_t, cudaStream_t) {
    set_status(kErrUnsupported, "atoma_gemm_bf16_tn: not built yet");
    return false;
}
}  // namespace atoma
