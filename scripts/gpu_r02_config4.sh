#!/bin/bash
# BASELINE.json configs[3] for real: Llama-3.1-70B, TP = 8 over NVLink, batch 16, ctx 4096 (gpurun --gpus 8).
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --config 4 --gpus 8 --steps 20 --warmup 5 --no-tp-parity --verbose > gpurun_out/bench_c4_tp8_ours.json 2> gpurun_out/bench_c4_tp8_ours.err
tail -3 gpurun_out/bench_c4_tp8_ours.err; cut -c1-400 gpurun_out/bench_c4_tp8_ours.json
