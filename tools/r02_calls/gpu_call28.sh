#!/bin/bash
# streamed GEMM: tickets / dependency checks / fences per batch of 4 (or 8) tiles instead of per tile
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
cp $L/libTB4.so $L/libflappie_b200.so
timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py -m gpu -q -k "streamed or repeat_bitwise" 2>&1 | tail -2
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -2
for v in TB1 TB4 TB8 TB1 TB4; do
  cp $L/lib$v.so $L/libflappie_b200.so
  for m in r941_native r941_native_gru; do
    timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --model $m > gpurun_out/c28_bench_${m}_$v.txt 2>&1
    python - "$v" "$m" <<'P'
import json,sys
v,m=sys.argv[1],sys.argv[2]
for l in open(f'gpurun_out/c28_bench_{m}_{v}.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(v, m, 'value ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['roofline']['step_breakdown_ms'].items()})
P
  done
done
for v in TB1 TB4; do
  cp $L/lib$v.so $L/libflappie_b200.so
  timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r941_5mC --reads 4096 > gpurun_out/c28_bench_5mC_$v.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --model r941_rna002 > gpurun_out/c28_bench_lstm256_$v.txt 2>&1
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
for m in ('5mC','lstm256'):
    for l in open(f'gpurun_out/c28_bench_{m}_{v}.txt'):
        if l.startswith('{'):
            d=json.loads(l); print(v, m, 'value ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'])
P
done
cp $L/libTB4P.so $L/libflappie_b200.so
timeout -s KILL 200 python tools/step_timeline.py r941_native 1024 2>&1 | tee gpurun_out/c28_timeline_lstm384_TB4.txt | grep -E "launch [01234] |^#" | cut -c1-220
cp $L/libTB4.so $L/libflappie_b200.so
