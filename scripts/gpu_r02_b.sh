#!/usr/bin/env bash
# Half-footprint GEMM experiment: parity tests of the GEMMs under both budgets, per-kernel timing, step A/B.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_llama_ops_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_b.log
timeout 900 python scripts/microbench_step.py --layers 8 --variants default,b100,b80,b64,nopdl,default > gpurun_out/step_ab.log 2>&1
timeout 900 python scripts/microbench_skinny.py --budgets > gpurun_out/skinny_budgets.log 2>&1
cat gpurun_out/pytest_b.log; cat gpurun_out/step_ab.log | tail -8; grep -E "budget=|np=" gpurun_out/skinny_budgets.log
