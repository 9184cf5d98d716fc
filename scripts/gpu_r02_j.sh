#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llama_ops_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 600 python scripts/microbench_step.py --layers 8 --variants fuse0,fuse2,fuse0,fuse2 --out gpurun_out/step_epi_ab.json 2>&1 | tail -4
timeout 300 python scripts/microbench_skinny.py --quick 2>&1 | grep -E '"note": "pdl' | cut -c1-200
