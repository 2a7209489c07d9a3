#!/bin/bash
# K = 512 W-stationary GEMM (lo plane of the panel in shared memory, K-split accumulators): numerics + r103_native timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 100 python -m pytest tests/test_gpu_tensor.py -m gpu -q -k "gemm_tc_matches" 2>&1 | tail -3 | cut -c1-300
timeout -s KILL 100 python -m pytest tests/test_gpu_hardening.py -m gpu -q -s -k "s512" 2>&1 | grep -E "parity\]|passed|failed" | cut -c1-300
timeout -s KILL 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "512" 2>&1 | tail -1
for e in 0 1; do
  if [ $e = 1 ]; then export FFB_GEMM_NO_WS512=1; fi
  timeout -s KILL 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r103_native 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('r103 NO_WS512=$e value ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})"
done
