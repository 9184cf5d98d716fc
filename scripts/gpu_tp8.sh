#!/usr/bin/env bash
# 8-GPU validation: P2P all-reduce parity at world 4, bench at N=8 (p2p and nccl) and N=4.
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_p2p_allreduce_gpu.py -m gpu -x -q 2>&1 | tail -4
for cfg in "8 p2p" "8 nccl" "4 p2p"; do
  set -- $cfg
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29540 \
      bench.py --gpus $1 --steps 20 --warmup 5 --tp-comm $2 --no-cpu-baseline > gpurun_out/bench_tp$1_$2.json 2> gpurun_out/bench_tp$1_$2.err
  echo "N=$1 $2 exit $?"; python -c "
import sys,json
try:
    d=json.loads(open('gpurun_out/bench_tp$1_$2.json').read()); print(d['config']['parallelism'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['config']['all_reduce'], d['roofline']['us_per_launch'])
except Exception as e: print('no json', e)"
  grep -v "^$\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_tp$1_$2.err | tail -3 | cut -c1-300
done
