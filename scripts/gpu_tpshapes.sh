#!/usr/bin/env bash
# Decode kernel on tensor-parallel shard shapes (Hkv = 8, 4, 2, 1 per rank), merged K+V ops vs two maps.
set -u
mkdir -p gpurun_out
for hk in 8 4 2 1; do
  ATOMA_MB_HK=$hk timeout 120 python scripts/microbench_decode.py 4096 --only-ours 2>&1 | tail -1
  ATOMA_MB_HK=$hk ATOMA_MB_NO_MERGE=1 timeout 120 python scripts/microbench_decode.py 4096 --only-ours 2>&1 | tail -1
done
