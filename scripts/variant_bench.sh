#!/bin/bash
# Time the likelihood kernel of every library under variants/ through bench.py (burnt-in C4 ensemble).
# usage: scripts/variant_bench.sh [workload]
wl=${1:-c4}
for so in variants/libemp_*.so; do
  EMP_B200_LIB=$PWD/$so python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --legs none 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$so', 'kernel %.3f ms/launch' % r['avg_launch_ms'], 'ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'evalfrac %.3f' % r['evaluated_fraction'])"
done
