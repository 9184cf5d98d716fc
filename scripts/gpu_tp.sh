#!/usr/bin/env bash
# Multi-GPU call (gpurun --gpus N): NCCL binding test + tensor-parallel bench with both collective paths.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nccl_binding.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_nccl.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=WARN timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --verbose --no-cpu-baseline > gpurun_out/bench_tp${N}_nccl.json 2> gpurun_out/bench_tp${N}_nccl.err; echo "exit $?" >> gpurun_out/bench_tp${N}_nccl.err
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --tp-comm torch --no-cpu-baseline > gpurun_out/bench_tp${N}_torch.json 2> gpurun_out/bench_tp${N}_torch.err; echo "exit $?" >> gpurun_out/bench_tp${N}_torch.err
timeout 400 $RUN bench.py --gpus $N --steps 10 --warmup 3 --impl reference > gpurun_out/bench_tp${N}_ref.json 2> gpurun_out/bench_tp${N}_ref.err; echo "exit $?" >> gpurun_out/bench_tp${N}_ref.err
cat gpurun_out/pytest_nccl.log; for f in nccl torch ref; do cut -c1-330 gpurun_out/bench_tp${N}_$f.json; tail -6 gpurun_out/bench_tp${N}_$f.err | cut -c1-300; done
