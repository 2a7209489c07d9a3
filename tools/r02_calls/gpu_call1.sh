#!/bin/bash
# round-2 call 1: where are we on today's box?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_smi.txt 2>&1
tools/microbench/cluster_probe > gpurun_out/c1_cluster_probe.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.txt 2>&1
cp flappie_b200/csrc/libflappie_b200.so flappie_b200/csrc/libN.so
cp flappie_b200/csrc/libP.so flappie_b200/csrc/libflappie_b200.so
timeout 300 python tools/rnn_phase_profile.py r941_native_gru 1024 > gpurun_out/c1_phase_gru256.txt 2>&1
timeout 300 python tools/rnn_phase_profile.py r941_5mC 4096 > gpurun_out/c1_phase_5mC_4096.txt 2>&1
cp flappie_b200/csrc/libN.so flappie_b200/csrc/libflappie_b200.so
tools/sanitize.sh gpurun_out
tail -3 gpurun_out/c1_pytest.txt; cat gpurun_out/c1_cluster_probe.txt | head -40; tail -2 gpurun_out/c1_bench.txt | cut -c1-600
