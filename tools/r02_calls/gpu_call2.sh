#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.txt
timeout 300 python tools/report_parity.py gru > gpurun_out/c2_parity_gru.txt 2>&1
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_fuse_$rep.txt 2>&1
  FFB_NO_FUSE_Z=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_nofuse_$rep.txt 2>&1
done
cp flappie_b200/csrc/libP.so flappie_b200/csrc/libflappie_b200.so
timeout 300 python tools/rnn_phase_profile.py r941_native_gru 1024 > gpurun_out/c2_phase_gru256.txt 2>&1
cp flappie_b200/csrc/libN.so flappie_b200/csrc/libflappie_b200.so
tail -3 gpurun_out/c2_pytest.txt; cat gpurun_out/c2_parity_gru.txt
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c2_bench_*.txt')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
P
cat gpurun_out/c2_phase_gru256.txt
