#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, both bench arms (+ our arm with the hand-written GEMM).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?" >> gpurun_out/bench_ours.err
timeout 400 python bench.py --steps 20 --warmup 5 --gemm atoma --no-cpu-baseline > gpurun_out/bench_ours_atoma_gemm.json 2> gpurun_out/bench_ours_atoma_gemm.err; echo "bench exit $?" >> gpurun_out/bench_ours_atoma_gemm.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench exit $?" >> gpurun_out/bench_ref.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours_atoma_gemm.json; tail -3 gpurun_out/bench_ours_atoma_gemm.err; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
