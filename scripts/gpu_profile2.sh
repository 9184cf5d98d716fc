#!/usr/bin/env bash
# ncu evidence: prefill kernel (tensor pipe), GEMM kernel, reference prefill kernel durations; cache-op bandwidths.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_cache_ops.py > gpurun_out/bench_cache_ops.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill_attn -s 3 -c 1 -f -o gpurun_out/prefill_full \
    python scripts/microbench_prefill.py --no-ref > gpurun_out/ncu_prefill.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flash_fwd_kernel|prefill_attn" -c 40 --csv --log-file gpurun_out/launches_prefill.csv \
    python scripts/microbench_prefill.py > gpurun_out/ncu_prefill2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_stream -s 10 -c 1 -f -o gpurun_out/gemm_full \
    python scripts/microbench_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -20 gpurun_out/bench_cache_ops.log
