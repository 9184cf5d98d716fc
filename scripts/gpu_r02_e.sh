#!/usr/bin/env bash
# Fused add+RMSNorm tail: parity tests (wrapped in a timeout: a protocol bug would spin), A/B of the step, timeline.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llama_ops_gpu.py tests/test_gemm_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_e.log; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_e.log
cat gpurun_out/pytest_e.log
timeout 600 python scripts/microbench_step.py --layers 8 --variants fuse1,fuse0,fuse1,fuse0 --out gpurun_out/step_fusion_ab.json > gpurun_out/step_fusion_ab.log 2>&1; tail -5 gpurun_out/step_fusion_ab.log
timeout 600 python scripts/microbench_step.py --layers 8 --tp 8 --variants fuse1,fuse0,fuse1,fuse0 --out gpurun_out/step_fusion_ab_tp8.json > gpurun_out/step_fusion_ab_tp8.log 2>&1; tail -5 gpurun_out/step_fusion_ab_tp8.log
timeout 600 python scripts/trace_step.py --layers 4 > gpurun_out/trace_step.log 2>&1; tail -24 gpurun_out/trace_step.log | cut -c1-500
