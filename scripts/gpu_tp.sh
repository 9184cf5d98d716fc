#!/usr/bin/env bash
# Multi-GPU call (gpurun --gpus N): collective tests + tensor-parallel bench with each all-reduce path.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nccl_binding.py tests/test_p2p_allreduce_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_tp.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for comm in ${COMMS:-p2p nccl torch}; do
  timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 --tp-comm $comm --no-cpu-baseline > gpurun_out/bench_tp${N}_$comm.json 2> gpurun_out/bench_tp${N}_$comm.err; echo "exit $?" >> gpurun_out/bench_tp${N}_$comm.err
done
cat gpurun_out/pytest_tp.log; for f in ${COMMS:-p2p nccl torch}; do cut -c1-330 gpurun_out/bench_tp${N}_$f.json; grep -v "^$\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_tp${N}_$f.err | tail -4 | cut -c1-300; done
