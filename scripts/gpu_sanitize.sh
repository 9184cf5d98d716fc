#!/usr/bin/env bash
# compute-sanitizer over the kernels whose correctness rests on hand-written synchronisation (the
# reference has no race detection; the decode kernel's cross-CTA arrival counters, the GEMM's mbarrier ring and
# the decode step's PDL chain are what memcheck / racecheck / synccheck are for).  One GPU.  Logs: gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
SEL_STEP='tests/test_llama_ops_gpu.py::test_decode_step_matches_the_oracle_and_the_composition_of_its_operators tests/test_llama_ops_gpu.py::test_decode_step_pdl_modes_are_bit_identical'
SEL_OPS='tests/test_llama_ops_gpu.py::test_qkv_rope_cache_fused_is_bit_identical_to_unfused_and_matches_oracle tests/test_llama_ops_gpu.py::test_gate_up_silu tests/test_llama_ops_gpu.py::test_add_rmsnorm tests/test_gemm_gpu.py::test_gemm_exact_small_integers'
SEL_ATT='tests/test_attention_gpu.py::test_peaky_softmax_range tests/test_cache_ops_gpu.py'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 \
      python -m pytest $SEL_STEP $SEL_OPS $SEL_ATT -x -q -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -4
done
