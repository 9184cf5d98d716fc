#!/bin/bash
# Tensor-parallel evidence on N GPUs (gpurun --gpus N -- bash scripts/gpu_tp_r02.sh N [noref]): collective tests, then the
# bench at N ranks for both arms (same-N reference arm), torchrun launch exactly like the driver's.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_step_gpu.py tests/test_p2p_allreduce_gpu.py tests/test_nccl_binding.py -q -x > gpurun_out/pytest_tp${N}.log 2>&1
tail -5 gpurun_out/pytest_tp${N}.log
PORT=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 20 --warmup 5 --verbose > gpurun_out/bench_tp${N}_ours.json 2> gpurun_out/bench_tp${N}_ours.err
tail -3 gpurun_out/bench_tp${N}_ours.err; cat gpurun_out/bench_tp${N}_ours.json
if [ "${2:-}" != "noref" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --impl reference --gpus $N --steps 10 --warmup 3 --verbose > gpurun_out/bench_tp${N}_ref.json 2> gpurun_out/bench_tp${N}_ref.err
tail -3 gpurun_out/bench_tp${N}_ref.err; cat gpurun_out/bench_tp${N}_ref.json
fi
