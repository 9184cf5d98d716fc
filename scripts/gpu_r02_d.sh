#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/microbench_skinny.py --half-ring --quick > gpurun_out/skinny_half_ring.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_d.json 2> gpurun_out/bench_ours_d.err
grep -E "no pdl" gpurun_out/skinny_half_ring.log; cat gpurun_out/bench_ours_d.json | cut -c1-200
