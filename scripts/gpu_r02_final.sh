#!/usr/bin/env bash
# Round-2 evidence on ONE GPU: parity tests, smoke, both bench arms (headline + configs 1/3/4-shard), microbenchmarks,
# the ncu launch list of the bench command and `ncu --set full` captures of every kernel named in DESIGN.md.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench exit $?" >> gpurun_out/bench_ref.err
timeout 600 python bench.py --config 1 --steps 50 --warmup 10 > gpurun_out/bench_c1_ours.json 2> gpurun_out/bench_c1_ours.err
timeout 600 python bench.py --config 1 --impl reference --steps 20 --warmup 5 > gpurun_out/bench_c1_ref.json 2> gpurun_out/bench_c1_ref.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/bench_c3_ours.json 2> gpurun_out/bench_c3_ours.err
timeout 600 python bench.py --config 4 --emulate-tp 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_rank_shard.json 2> gpurun_out/bench_c4_rank_shard.err
timeout 900 python scripts/microbench_skinny.py > gpurun_out/microbench_skinny.log 2>&1
timeout 600 python scripts/microbench_prefill.py > gpurun_out/microbench_prefill.log 2>&1
timeout 600 python scripts/microbench_gemm.py > gpurun_out/microbench_gemm.log 2>&1
timeout 600 python scripts/bench_cache_ops.py > gpurun_out/bench_cache_ops.log 2>&1
for hk in 8 4 2 1; do ATOMA_MB_HK=$hk timeout 300 python scripts/microbench_decode.py 4096 --no-ref > gpurun_out/microbench_decode_hk$hk.log 2>&1; done
# launch list of the bench command (the step's kernels only), cold-cache and serialised: shares, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_skinny|add_rmsnorm|decode_paged|embedding|argmax" -c 700 \
    --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|decode_paged|add_rmsnorm|argmax|embedding" -c 12 -f -o gpurun_out/r02_step_full \
    python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn -s 3 -c 1 -f -o gpurun_out/r02_prefill_full \
    python scripts/microbench_prefill.py --no-ref > gpurun_out/ncu_prefill.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"copy_blocks_bulk|reshape_and_cache_flash_vec|swap_blocks_bulk" -c 4 -f -o gpurun_out/r02_cache_ops_full \
    python scripts/bench_cache_ops.py > gpurun_out/ncu_cache_ops.log 2>&1
ATOMA_MB_HK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_paged -s 10 -c 1 -f -o gpurun_out/r02_decode_hk1_full \
    python scripts/microbench_decode.py 4096 --no-ref > gpurun_out/ncu_decode_hk1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rope_reshape_and_cache -c 1 -f -o gpurun_out/r02_rope_full \
    python -m pytest tests/test_rope_gpu.py -x -q > gpurun_out/ncu_rope.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-250 gpurun_out/bench_ours.json; cut -c1-250 gpurun_out/bench_ref.json
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/r02_launches_bench.csv
