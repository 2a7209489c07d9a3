#!/bin/bash
# timing-only ablations of the recurrent kernel (FFB_RNN_ABLATE bit mask, see csrc/rnn_tc.cu): what bounds the step?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
for m in 0 1 2 4 8 16 32 5 37 63 0; do
  cp $L/libA$m.so $L/libflappie_b200.so
  timeout -s KILL 200 python tools/ablate_timing.py A$m 2>&1 | grep -v "^$" | tee -a gpurun_out/c16_ablate.txt
done
cp $L/libA0.so $L/libflappie_b200.so
