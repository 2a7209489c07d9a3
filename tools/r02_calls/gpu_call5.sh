#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp flappie_b200/csrc/libflappie_b200.so flappie_b200/csrc/libKEEP.so
for v in G0 G1 G2; do
  cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
  timeout 300 python tools/report_parity.py gru > gpurun_out/c5_parity_gru_$v.txt 2>&1
  timeout 300 python tools/report_parity.py lstm > gpurun_out/c5_parity_lstm_$v.txt 2>&1
  echo "== $v"; grep tensor gpurun_out/c5_parity_gru_$v.txt gpurun_out/c5_parity_lstm_$v.txt
  timeout 600 python tools/parity_report.py 32 > gpurun_out/c5_report_$v.txt 2>&1; grep -v "^#" gpurun_out/c5_report_$v.txt
done
for rep in 1 2; do
  for v in G0 G1 G2; do
    cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c5_bench_${v}_$rep.txt 2>&1
  done
done
cp flappie_b200/csrc/libKEEP.so flappie_b200/csrc/libflappie_b200.so
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5_bench_*.txt')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True
            d=json.loads(l); print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
    if not ok: print(f, open(f).read()[-1500:])
P
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c5_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR" gpurun_out/c5_pytest.txt | tail -30
