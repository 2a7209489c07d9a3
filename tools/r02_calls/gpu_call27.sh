#!/bin/bash
# step timeline from in-kernel globaltimer stamps (profile build), streamed schedule, one context
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=flappie_b200/csrc
cp $L/libPROF.so $L/libflappie_b200.so
timeout -s KILL 200 python tools/step_timeline.py r941_native_gru 1024 2>&1 | tee gpurun_out/c27_timeline_gru.txt | cut -c1-200
timeout -s KILL 200 python tools/step_timeline.py r941_native 1024 2>&1 | tee gpurun_out/c27_timeline_lstm384.txt | cut -c1-200
cp $L/libFINAL.so $L/libflappie_b200.so
