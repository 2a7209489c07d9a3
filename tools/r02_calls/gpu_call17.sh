#!/bin/bash
# the ablations of call 16 again, in CLOCK CYCLES (the boards are power-capped: removing work raises the clock): -DFFB_RNN_PROFILE builds
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
for m in ${MASKS:-0 1 2 4 16 32 0}; do
  cp $L/libAP$m.so $L/libflappie_b200.so
  echo "== ablate mask $m" | tee -a gpurun_out/c17_ablate_cycles.txt
  timeout -s KILL 200 python tools/rnn_phase_profile.py r941_native_gru 1024 2>&1 | head -12 | tee -a gpurun_out/c17_ablate_cycles.txt
done
cp $L/libFINAL.so $L/libflappie_b200.so
