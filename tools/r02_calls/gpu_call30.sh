#!/bin/bash
# streamed GEMM tickets: adaptive batch (four while behind the recurrence, one at the frontier) vs one ticket per tile, ragged + uniform
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
cp $L/libTBA.so $L/libflappie_b200.so
timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py tests/test_gpu_configs.py -m gpu -q -k "streamed or repeat_bitwise or cfg" 2>&1 | tail -2
for v in TB1 TBA TB1 TBA; do
  cp $L/lib$v.so $L/libflappie_b200.so
  echo "== $v"
  timeout -s KILL 300 python tools/mixed_bench.py 3072 1000 50000 2>&1 | tail -2 | cut -c1-200
  timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --model r941_native 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lstm384 value ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
cp $L/libTBA.so $L/libflappie_b200.so
