#!/bin/bash
# Produce the round's measurement artefacts on the GPU box into gpurun_out/<tag>_* (copied into profiles/ by hand):
#   bench lines (c4 with cpu_baseline, reference arm, c5, c2), the ncu launch list of the bench command,
#   one `ncu --set full` capture of the likelihood kernel (raw metrics exported as CSV), compute-sanitizer.
# usage: scripts/capture_profiles.sh <tag>
tag=${1:-r01}
o=gpurun_out
mkdir -p $o
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench_c4.json 2> $o/${tag}_bench_c4.err
python bench.py --workload c5 --steps 4 --warmup 3 --burn 20 --no-cpu-baseline > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err
python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err
# launch list: same command as the bench, short (burn-in 30 so that the listed steps are of a burnt-in chain)
ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 120 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 --burn 30 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
# full capture of the dominant kernel: a launch of the timed region (burn 30 -> launch index ~70)
ncu --set full --import-source on --clock-control none -k regex:logl_rv -s 66 -c 1 -o $o/${tag}_logl -f \
    python bench.py --steps 1 --warmup 3 --burn 30 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
ncu -i $o/${tag}_logl.ncu-rep --page raw --csv > $o/${tag}_logl_raw.csv 2>/dev/null
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_sanitizer.log 2>&1
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" >> $o/${tag}_sanitizer.log 2>&1
grep "ERROR SUMMARY" $o/${tag}_sanitizer.log
tail -3 $o/${tag}_sanitizer.log
for f in reference c4 c5 c2; do python - <<EOF
import json
try:
    d = json.loads(open("$o/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("$f", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"],
          "kernel ms %.3f frac %.3f" % (r.get("avg_launch_ms", 0), r.get("frac", 0)) if r else "", d.get("clocks"))
except Exception as e:
    print("$f", "FAILED", e)
EOF
done
