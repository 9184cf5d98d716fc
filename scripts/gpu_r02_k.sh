#!/usr/bin/env bash
# Prefill softmax: fraction of exponentials on the FMA pipe (poly_exp2) — 0 / 12.5 / 25 (default build) / 50 %
set -u
mkdir -p gpurun_out
for v in pf0 pf4 default pf1; do
  if [ $v = default ]; then unset ATOMA_FLASH_LIB; else export ATOMA_FLASH_LIB=$PWD/scratch_$v/libflashattention.so; fi
  echo "== $v"; timeout 300 python scripts/microbench_prefill.py --no-ref 2>&1 | grep -o '"lens": \[[0-9, ]*\].*"ours_TFLOPs": [0-9.]*' | sed 's/"flops.*ours_TFLOPs"/TFLOPs/' 
done
unset ATOMA_FLASH_LIB
timeout 600 python -m pytest tests/test_prefill_gpu.py tests/test_attention_gpu.py tests/test_parity_reference_gpu.py -x -q 2>&1 | tail -4
