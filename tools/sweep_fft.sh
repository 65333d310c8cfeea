#!/bin/bash
# A/B sweep of the overlap-save FIR kernel's tuning knobs on C2 (run under gpurun); prints value per setting.
out=gpurun_out/sweep_fft.txt
: > $out
run() {  # dtype warps pack stage prefetch
  v=$(LLZ_FFT_WARPS=$2 LLZ_FFT_PACK=$3 LLZ_FFT_STAGE=$4 LLZ_FFT_PREFETCH=$5 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --dtype $1 2>&1 | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.readline()); print(round(d["value"]), round(d["ms_per_step"],4), d["roofline"]["frac"])
except Exception as e: print("failed", e)')
  echo "$1 warps=$2 pack=$3 stage=$4 prefetch=$5 -> $v" >> $out
}
run f64 8 0 0 1
run f64 8 0 1 0
run f64 8 1 0 1
run f64 12 0 0 0
run f32 16 0 0 1
run f32 16 0 1 0
run f32 16 1 0 1
run f32 12 1 1 0
run f32 20 1 0 1
cat $out
