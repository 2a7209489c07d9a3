#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.txt
tail -3 gpurun_out/c3_pytest.txt
for v in A B C D; do
  cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
  timeout 300 python tools/report_parity.py gru > gpurun_out/c3_parity_gru_$v.txt 2>&1
  timeout 300 python tools/report_parity.py lstm > gpurun_out/c3_parity_lstm_$v.txt 2>&1
  echo "== $v"; grep tensor gpurun_out/c3_parity_gru_$v.txt gpurun_out/c3_parity_lstm_$v.txt
done
for rep in 1 2; do
  for v in A B C D; do
    cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_${v}_$rep.txt 2>&1
  done
done
cp flappie_b200/csrc/libA.so flappie_b200/csrc/libflappie_b200.so
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c3_bench_*.txt')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
P
