#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/c4_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.txt
grep -E "\[parity\]|passed|failed|Error|error" gpurun_out/c4_pytest.txt | tail -30
timeout 600 python tools/parity_report.py 64 > gpurun_out/c4_parity_report.txt 2>&1
cat gpurun_out/c4_parity_report.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.txt 2>&1
tail -1 gpurun_out/c4_bench.txt | cut -c1-1500
