#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/microbench_step.py --layers 8 --variants default,dbuf,default,dbuf --out gpurun_out/step_tmem_ab.json > gpurun_out/step_tmem_ab.log 2>&1; tail -5 gpurun_out/step_tmem_ab.log
timeout 600 python scripts/trace_step.py --layers 3 > gpurun_out/trace_step2.log 2>&1; tail -8 gpurun_out/trace_step2.log | cut -c1-420
timeout 300 python -m pytest tests/test_llama_ops_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
