#!/bin/bash
# Per-kernel device time of the swap phase (plan + application) at ladder shapes of the scaling runs (development aid).
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pt_swap_plan|pt_apply|pt_publish" -c 400 --csv \
    --log-file gpurun_out/${1:-swap}_phase_launches.csv env PLAN_BENCH_PRODUCT_ONLY=1 python scripts/plan_bench.py > gpurun_out/${1:-swap}_phase.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${1:-swap}_phase_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[(r[4].split("(")[0], r[8])].append(float(r[-1]))
for k, v in agg.items():
    print(k, len(v), "launches", [round(x / 1e3, 1) for x in v[:8]], "us")
PY
