#!/usr/bin/env bash
# ncu --set full of the cache-op kernels at their LARGE configurations (the first launches of bench_cache_ops.py are the
# one-pair cases): copy_blocks with 1024 pairs x 32 layers, reshape_and_cache_flash at 8192 tokens, swap (device to device).
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copy_blocks_bulk -s 56 -c 1 -f -o gpurun_out/r02_copy_blocks_full \
    python scripts/bench_cache_ops.py > gpurun_out/ncu_copy_blocks.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reshape_and_cache_flash_vec -s 50 -c 1 -f -o gpurun_out/r02_reshape_full \
    python scripts/bench_cache_ops.py > gpurun_out/ncu_reshape.log 2>&1
ls -la gpurun_out/r02_copy_blocks_full.ncu-rep gpurun_out/r02_reshape_full.ncu-rep
