#!/bin/bash
# S = 512: input projection with hi*hi split over two accumulators (K-split) -- parity of lstm512 against the oracle
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tensor.py -m gpu -q > gpurun_out/c21_pytest_tensor.txt 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest_tensor.txt; tail -3 gpurun_out/c21_pytest_tensor.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py -m gpu -q -s -k "s512" > gpurun_out/c21_pytest_s512.txt 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest_s512.txt; grep -E "parity\]|passed|failed" gpurun_out/c21_pytest_s512.txt | tail -5
FFB_GEMM_NO_KSPLIT=1 timeout -s KILL 600 python -m pytest tests/test_gpu_hardening.py -m gpu -q -s -k "s512" > gpurun_out/c21_pytest_s512_nosplit.txt 2>&1; grep -E "parity\]|passed|failed" gpurun_out/c21_pytest_s512_nosplit.txt | tail -5
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "512" > gpurun_out/c21_pytest_parity512.txt 2>&1; tail -2 gpurun_out/c21_pytest_parity512.txt
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r103_native > gpurun_out/c21_bench_r103.txt 2>&1
FFB_GEMM_NO_KSPLIT=1 timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --model r103_native > gpurun_out/c21_bench_r103_nosplit.txt 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c21_bench_*.txt')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['step_breakdown_ms'].items()})
P
