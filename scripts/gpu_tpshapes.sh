#!/usr/bin/env bash
# Decode kernel on tensor-parallel shard shapes (Hkv = 8, 4, 2, 1 per rank) + per-CTA timelines.
set -u
mkdir -p gpurun_out
for hk in 8 4 2 1; do
  ATOMA_MB_HK=$hk timeout 120 python scripts/microbench_decode.py 4096 --only-ours 2>&1 | tail -1
done
for hk in 4 1; do
  echo "== trace hk=$hk"; ATOMA_MB_HK=$hk timeout 120 python scripts/trace_decode.py 2>&1 | head -12
done
