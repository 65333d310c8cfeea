#!/bin/bash
# compute-sanitizer over one small launch of every kernel family (tools/sanitize_kernels.py).
# Usage (GPU box): bash tools/run_sanitizers.sh [out_dir]   -> <out_dir>/sanitize_{memcheck,racecheck,synccheck}.log
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
python tools/sanitize_kernels.py > "$OUT/sanitize_plain.log" 2>&1 || { echo "plain run failed"; tail -5 "$OUT/sanitize_plain.log"; exit 1; }
for tool in memcheck synccheck racecheck; do
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    timeout ${SAN_TIMEOUT:-900} $CS --tool $tool $extra --print-limit 30 --target-processes all \
        python tools/sanitize_kernels.py > "$OUT/sanitize_$tool.log" 2>&1
    echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_$tool.log" | tail -1)"
done
