#!/usr/bin/env bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the decode kernel.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:decode_paged -s 20 -c 2 -f -o gpurun_out/decode_full \
    python scripts/microbench_decode.py 4096 --no-ref > gpurun_out/ncu_decode.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_bench.csv; tail -3 gpurun_out/ncu_decode.log
