#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --config 4 --impl reference --gpus 8 --steps 10 --warmup 3 --verbose > gpurun_out/bench_c4_tp8_ref.json 2> gpurun_out/bench_c4_tp8_ref.err
tail -3 gpurun_out/bench_c4_tp8_ref.err; cut -c1-300 gpurun_out/bench_c4_tp8_ref.json
