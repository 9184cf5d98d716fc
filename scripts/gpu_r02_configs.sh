#!/usr/bin/env bash
# The non-headline configurations of BASELINE.json on ONE GPU: config 1 (Llama-3.2-1B, single stream) both arms,
# config 3 (continuous-batching mix through FlashAttention.forward), config 4 as one rank's shard (--emulate-tp 8).
set -u
mkdir -p gpurun_out
timeout 600 python scripts/trace_step.py --layers 4 > gpurun_out/trace_step.log 2>&1
timeout 600 python bench.py --config 1 --steps 50 --warmup 10 > gpurun_out/bench_c1_ours.json 2> gpurun_out/bench_c1_ours.err
timeout 600 python bench.py --config 1 --impl reference --steps 20 --warmup 5 > gpurun_out/bench_c1_ref.json 2> gpurun_out/bench_c1_ref.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/bench_c3_ours.json 2> gpurun_out/bench_c3_ours.err
timeout 600 python bench.py --config 4 --emulate-tp 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_rank_shard.json 2> gpurun_out/bench_c4_rank_shard.err
timeout 600 python bench.py --emulate-tp 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_tp8_rank_shard.json 2> gpurun_out/bench_c2_tp8_rank_shard.err
tail -30 gpurun_out/trace_step.log | cut -c1-420
for f in c1_ours c1_ref c3_ours c4_rank_shard c2_tp8_rank_shard; do echo "== $f"; tail -2 gpurun_out/bench_$f.err | cut -c1-300; cut -c1-600 gpurun_out/bench_$f.json; done
