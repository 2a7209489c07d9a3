#!/bin/bash
# compute-sanitizer over one small batch per recurrent-kernel instance; logs -> $1 (default gpurun_out)
out=${1:-gpurun_out}; mkdir -p "$out"
cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
for model in ${MODELS:-r941_native_gru r941_rna002 r941_native r103_native}; do
  for tool in ${TOOLS:-memcheck synccheck racecheck}; do
    log="$out/sanitizer_${tool}_${model}.log"
    timeout 420 $CS --tool $tool --print-limit 20 --log-file "$log.raw" python tools/sanitize_run.py $model 32 700 > "$log.stdout" 2>&1
    echo "exit $? tool=$tool model=$model" >> "$log.stdout"
    { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" "$log.raw" | sort | uniq -c | sort -rn | head -20; tail -3 "$log.stdout"; } > "$log"
    rm -f "$log.stdout"; head -c 200000 "$log.raw" > "$log.head"; rm -f "$log.raw"
  done
done
