#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp flappie_b200/csrc/libflappie_b200.so flappie_b200/csrc/libKEEP.so
for v in K0 K4 K8 K12; do
  cp flappie_b200/csrc/lib$v.so flappie_b200/csrc/libflappie_b200.so
  echo "== $v"
  timeout 600 python tools/satur_report.py > gpurun_out/c9_satur_$v.txt 2>&1; cat gpurun_out/c9_satur_$v.txt | tail -6
  timeout 300 python tools/report_parity.py gru > gpurun_out/c9_parity_gru_$v.txt 2>&1
  timeout 300 python tools/report_parity.py lstm > gpurun_out/c9_parity_lstm_$v.txt 2>&1
  grep tensor gpurun_out/c9_parity_gru_$v.txt gpurun_out/c9_parity_lstm_$v.txt
  timeout 600 python tools/parity_report.py 32 > gpurun_out/c9_report_$v.txt 2>&1; grep -v "^#" gpurun_out/c9_report_$v.txt
done
cp flappie_b200/csrc/libKEEP.so flappie_b200/csrc/libflappie_b200.so
timeout 600 python tools/slots_ab.py > gpurun_out/r02_slots_ab3.txt 2>&1; grep -E "auto" gpurun_out/r02_slots_ab3.txt
