#!/usr/bin/env bash
# Round-2 evidence on ONE GPU (gpurun -- bash scripts/gpu_r02_profile.sh): GPU parity tests, smoke, both bench arms,
# the ncu launch list of the bench command, and `ncu --set full` captures of every kernel of the decode step.
# Outputs under gpurun_out/; the summaries that are judged get copied to profiles/r02_* afterwards.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench exit $?" >> gpurun_out/bench_ref.err
# launch list: every kernel of the bench command with its device time (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
# full captures: the four projection GEMMs of one layer (+ lm_head), the decode kernel, add+RMSNorm
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -c 5 -f -o gpurun_out/r02_skinny_full \
    python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_skinny.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_paged|add_rmsnorm|argmax|embedding" -c 6 -f -o gpurun_out/r02_step_others_full \
    python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_others.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err; wc -l gpurun_out/r02_launches_bench.csv; tail -3 gpurun_out/ncu_skinny.log; tail -3 gpurun_out/ncu_others.log
ls -la gpurun_out/*.ncu-rep
