#!/bin/bash
# A/B: alternate two builds of the library (flappie_b200/csrc/libA.so, libB.so) under the same conditions
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for v in A B; do
    cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
    else: print(l.rstrip()[-200:])
"
  done
done
cp flappie_b200/csrc/libB.so flappie_b200/csrc/libflappie_b200.so
