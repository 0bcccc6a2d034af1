#!/bin/bash
# same-box A/B of two builds of libudt_b200.so on the sustained (power-capped) request benchmark:
#   bash scripts/ab_bench.sh <a.so> <b.so> [rounds]      (paths relative to udifftext_b200/)
A=$1; B=$2; R=${3:-2}
for r in $(seq 1 $R); do
  for L in $A $B; do
    cp udifftext_b200/$L udifftext_b200/libudt_b200.so
    python bench.py --steps 6 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', round(d['value'],3), 'img/s', round(d['ms_per_step'],2), 'ms/request', d['clocks']['sm_mhz'], 'MHz')"
  done
done
