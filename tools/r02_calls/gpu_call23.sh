#!/bin/bash
# head gate position: before the trimming kernels (default) or only before the network (FFB_HEAD_GATE=conv)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2 3; do
  timeout -s KILL 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c23_bench_gateprep_$rep.txt 2>&1
  FFB_HEAD_GATE=conv timeout -s KILL 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c23_bench_gateconv_$rep.txt 2>&1
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c23_bench_*.txt')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True; print(f, 'value ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e M/s', round(d['e2e']['value']/1e6,1), d['clocks']['sm_mhz'])
    if not ok: print(f, 'NO JSON', open(f).read()[-400:])
P
FFB_HEAD_GATE=conv timeout -s KILL 400 python tools/cli_bench.py 32768 > gpurun_out/c23_cli_gateconv.txt 2>&1; grep -E "run|identical" gpurun_out/c23_cli_gateconv.txt | cut -c1-200
