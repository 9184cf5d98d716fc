#!/usr/bin/env bash
set -u
N=${1:-4}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_p2p_allreduce_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/nccl_diag.py 2>&1 | grep "^\[r0\]" | cut -c1-200
for algo in 1 2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 \
      bench.py --gpus $N --steps 20 --warmup 5 --ar-algo $algo --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('algo $algo', d['config']['parallelism'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
done
