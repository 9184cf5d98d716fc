#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llama_ops_gpu.py tests/test_gemm_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_h.log; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_h.log
cat gpurun_out/pytest_h.log
timeout 600 python scripts/microbench_step.py --layers 8 --variants fuse0,fuse2,fuse0,fuse2 --out gpurun_out/step_chain_ab.json > gpurun_out/step_chain_ab.log 2>&1; tail -5 gpurun_out/step_chain_ab.log
timeout 600 python scripts/microbench_step.py --layers 8 --tp 8 --variants fuse0,fuse2,fuse0,fuse2 --out gpurun_out/step_chain_ab_tp8.json > gpurun_out/step_chain_ab_tp8.log 2>&1; tail -5 gpurun_out/step_chain_ab_tp8.log
timeout 600 python scripts/trace_step.py --layers 3 > gpurun_out/trace_step4.log 2>&1; tail -12 gpurun_out/trace_step4.log | cut -c1-520
