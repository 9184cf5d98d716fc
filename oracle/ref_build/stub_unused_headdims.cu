// Test infrastructure (oracle build only; never linked into the product).
//
// The reference's flash_api.cu dispatches over 8 head dims x 2 dtypes x 2 causal x 2 variants
// (static_switch.h:96-123, flash_api.cu:8-20).  The oracle library only builds the instantiation
// TUs for head dims 32/64/128; this TU supplies the remaining specialisations as aborting stubs so
// the reference's own flash_api.cu links unmodified.
#include <cstdio>
#include <cstdlib>
#include <cutlass/numeric_types.h>
#include "flash.h"

#define ATOMA_ORACLE_STUB(T, HD, CAUSAL)                                                          \
    template <> void run_mha_fwd_<T, HD, CAUSAL>(Flash_fwd_params &, cudaStream_t) {              \
        std::fprintf(stderr, "oracle/_ref: head dim %d not built into the oracle library\n", HD); \
        std::abort();                                                                             \
    }                                                                                             \
    template <> void run_mha_fwd_splitkv_dispatch<T, HD, CAUSAL>(Flash_fwd_params &,              \
                                                                 cudaStream_t) {                  \
        std::fprintf(stderr, "oracle/_ref: head dim %d not built into the oracle library\n", HD); \
        std::abort();                                                                             \
    }

#define ATOMA_ORACLE_STUB_HD(HD)                       \
    ATOMA_ORACLE_STUB(cutlass::half_t, HD, false)      \
    ATOMA_ORACLE_STUB(cutlass::half_t, HD, true)       \
    ATOMA_ORACLE_STUB(cutlass::bfloat16_t, HD, false)  \
    ATOMA_ORACLE_STUB(cutlass::bfloat16_t, HD, true)

ATOMA_ORACLE_STUB_HD(96)
ATOMA_ORACLE_STUB_HD(160)
ATOMA_ORACLE_STUB_HD(192)
ATOMA_ORACLE_STUB_HD(224)
ATOMA_ORACLE_STUB_HD(256)
