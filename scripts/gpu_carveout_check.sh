#!/usr/bin/env bash
# Which shared-memory configuration does each kernel of the step actually get? (ncu launch metrics, one layer, eager)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics launch__shared_mem_config_size,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,gpu__time_duration.sum \
    --clock-control none -k regex:"gemm_skinny|add_rmsnorm|decode_paged|embedding|argmax" -c 12 --csv --log-file gpurun_out/carveout_check.csv \
    python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/carveout_check.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn = h.index("Kernel Name"); mn = h.index("Metric Name"); mv = h.index("Metric Value"); idc = h.index("ID")
cur = {}
for r in rows[hi + 1:]:
    if len(r) > mv:
        cur.setdefault((int(r[idc]), r[kn].split("(")[0][-50:]), {})[r[mn].replace("launch__", "")] = r[mv]
for k, v in cur.items():
    print(k, v)
PY
