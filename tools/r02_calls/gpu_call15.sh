#!/bin/bash
# Xin prefetch issued a whole step ahead (PE) vs at the top of the step (BASE = previous commit)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
cp $L/libPE.so $L/libflappie_b200.so
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c15_pytest_parity.txt 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest_parity.txt; tail -3 gpurun_out/c15_pytest_parity.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py -m gpu -q -k "streamed or repeat_bitwise" > gpurun_out/c15_pytest_hard.txt 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest_hard.txt; grep -E "parity\]|passed|failed" gpurun_out/c15_pytest_hard.txt | tail -5
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/c15_pytest_cfg.txt 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest_cfg.txt; tail -3 gpurun_out/c15_pytest_cfg.txt
for rep in 1 2 3; do
  for v in BASE PE; do
    cp $L/lib$v.so $L/libflappie_b200.so
    timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c15_bench_${v}_$rep.txt 2>&1
  done
done
for v in BASE PE; do
  cp $L/lib$v.so $L/libflappie_b200.so
  timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --model r941_native > gpurun_out/c15_bench_lstm384_${v}.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r941_5mC --reads 4096 > gpurun_out/c15_bench_5mC_${v}.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --model r941_rna002 > gpurun_out/c15_bench_lstm256_${v}.txt 2>&1
done
cp $L/libPEP.so $L/libflappie_b200.so
timeout -s KILL 300 python tools/rnn_phase_profile.py r941_native_gru 1024 > gpurun_out/c15_phase_1024_PE.txt 2>&1; head -13 gpurun_out/c15_phase_1024_PE.txt
cp $L/libPE.so $L/libflappie_b200.so
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c15_bench_*.txt')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok = True
            print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4), {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
    if not ok: print(f, 'NO JSON', open(f).read()[-300:])
P
