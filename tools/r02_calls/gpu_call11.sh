#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp flappie_b200/csrc/libflappie_b200.so flappie_b200/csrc/libKEEP.so
cp flappie_b200/csrc/libP.so flappie_b200/csrc/libflappie_b200.so
timeout 300 python tools/rnn_phase_profile.py r941_native_gru 1024 > gpurun_out/c11_phase_1024.txt 2>&1; tail -14 gpurun_out/c11_phase_1024.txt
timeout 300 python tools/rnn_phase_profile.py r941_5mC 4096 > gpurun_out/c11_phase_4096.txt 2>&1; tail -14 gpurun_out/c11_phase_4096.txt
FFB_TC_SLOTS=5 FFB_TC_CLUSTERS=13 timeout 300 python tools/rnn_phase_profile.py r941_5mC 4096 > gpurun_out/c11_phase_4096_5x13.txt 2>&1; tail -14 gpurun_out/c11_phase_4096_5x13.txt
cp flappie_b200/csrc/libKEEP.so flappie_b200/csrc/libflappie_b200.so
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c11_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR" gpurun_out/c11_pytest.txt | tail -30
