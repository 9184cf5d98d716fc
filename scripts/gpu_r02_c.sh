#!/usr/bin/env bash
# Cross-kernel L2 weight prefetch: parity tests, then step A/B over the prefetch edges and caps.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_llama_ops_gpu.py tests/test_pipeline_gpu.py tests/test_attention_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_c.log
timeout 900 python scripts/microbench_step.py --layers 8 --variants pf0,pf15,pf15_16,pf15_64,pf30,pf1,pf2,pf4,pf8,pf14,pf0,pf15 --out gpurun_out/step_prefetch_ab.json > gpurun_out/step_prefetch_ab.log 2>&1
cat gpurun_out/pytest_c.log; tail -14 gpurun_out/step_prefetch_ab.log
