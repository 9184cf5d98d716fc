#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_llama_ops_gpu.py -x -q -k "mlp_chain or pdl_modes or composition" 2>&1 | tail -3
timeout 600 python scripts/microbench_step.py --layers 8 --variants fuse0,fuse2,fuse0,fuse2 --out gpurun_out/step_chain_ab2.json 2>&1 | tail -4
timeout 600 python scripts/microbench_step.py --layers 8 --tp 8 --variants fuse0,fuse2,fuse0,fuse2 --out gpurun_out/step_chain_ab2_tp8.json 2>&1 | tail -4
