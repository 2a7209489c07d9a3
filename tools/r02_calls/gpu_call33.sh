#!/bin/bash
# compute-sanitizer on the K = 512 W-stationary GEMM (r103_native, one ragged 32-read batch): memcheck + racecheck
cd "$(dirname "$0")/../.."
MODELS="r103_native" TOOLS="memcheck racecheck" bash tools/sanitize.sh gpurun_out/sanitize_ws512
head -3 gpurun_out/sanitize_ws512/*.log
