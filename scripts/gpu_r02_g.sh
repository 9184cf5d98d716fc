#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/microbench_step.py --layers 8 --variants default,late1,late2,late4,default,late1 --out gpurun_out/step_trigger_ab.json > gpurun_out/step_trigger_ab.log 2>&1; tail -7 gpurun_out/step_trigger_ab.log
