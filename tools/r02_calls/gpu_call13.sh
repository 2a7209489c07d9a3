#!/bin/bash
# in-order MMA issuer (one issuer warp per CTA) vs the per-slot issuers of the previous commit (libBASE.so)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
cp $L/libNEW.so $L/libflappie_b200.so
timeout -s KILL 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c13_pytest_parity.txt 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest_parity.txt; tail -3 gpurun_out/c13_pytest_parity.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py -m gpu -x -q -k "streamed or repeat_bitwise or s512" > gpurun_out/c13_pytest_hard.txt 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest_hard.txt; tail -3 gpurun_out/c13_pytest_hard.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/c13_pytest_cfg.txt 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest_cfg.txt; tail -3 gpurun_out/c13_pytest_cfg.txt
for rep in 1 2; do
  for v in BASE NEW NEWG3; do
    cp $L/lib$v.so $L/libflappie_b200.so
    timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/c13_bench_${v}_$rep.txt 2>&1
  done
done
for v in BASE NEW; do
  cp $L/lib$v.so $L/libflappie_b200.so
  timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --model r941_native > gpurun_out/c13_bench_lstm384_${v}.txt 2>&1
  timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r941_5mC --reads 4096 > gpurun_out/c13_bench_5mC_${v}.txt 2>&1
done
cp $L/libNEWP.so $L/libflappie_b200.so
timeout -s KILL 300 python tools/rnn_phase_profile.py r941_native_gru 1024 > gpurun_out/c13_phase_1024.txt 2>&1; head -14 gpurun_out/c13_phase_1024.txt
cp $L/libNEWG3.so $L/libflappie_b200.so
timeout -s KILL 300 python tools/report_parity.py gru > gpurun_out/c13_parity_gru_G3.txt 2>&1; grep tensor gpurun_out/c13_parity_gru_G3.txt
cp $L/libNEW.so $L/libflappie_b200.so
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c13_bench_*.txt')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok = True
            print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4), {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
    if not ok: print(f, 'NO JSON', open(f).read()[-300:])
P
