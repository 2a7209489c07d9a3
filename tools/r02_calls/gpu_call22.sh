#!/bin/bash
# head gate (the next batch's signal prep / conv / first GEMM start when the batch in flight enters its last recurrent layer): e2e A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2 3; do
  FFB_NO_HEAD_GATE=1 timeout -s KILL 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c22_bench_nogate_$rep.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c22_bench_gate_$rep.txt 2>&1
done
for m in r941_native r941_rna002; do
  FFB_NO_HEAD_GATE=1 timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --model $m > gpurun_out/c22_bench_${m}_nogate.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --model $m > gpurun_out/c22_bench_${m}_gate.txt 2>&1
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c22_bench_*.txt')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True; print(f, 'value ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e M/s', round(d['e2e']['value']/1e6,1), d['clocks']['sm_mhz'])
    if not ok: print(f, 'NO JSON', open(f).read()[-400:])
P
timeout -s KILL 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_signal.py -m gpu -x -q 2>&1 | tail -2
FFB_NO_HEAD_GATE=1 timeout -s KILL 400 python tools/cli_bench.py 32768 > gpurun_out/c22_cli_nogate.txt 2>&1; grep -E "run|identical" gpurun_out/c22_cli_nogate.txt | cut -c1-200
timeout -s KILL 400 python tools/cli_bench.py 32768 > gpurun_out/c22_cli_gate.txt 2>&1; grep -E "run|identical" gpurun_out/c22_cli_gate.txt | cut -c1-200
