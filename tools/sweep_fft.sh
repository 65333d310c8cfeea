#!/bin/bash
# A/B sweep of the overlap-save FIR kernel's tuning knobs on C2 (run under gpurun); prints value per setting.
out=gpurun_out/sweep_fft.txt
: > $out
run() {  # dtype warps pack prefetch
  v=$(LLZ_FFT_WARPS=$2 LLZ_FFT_PACK=$3 LLZ_FFT_PREFETCH=$4 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --dtype $1 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["value"]), d["ms_per_step"], d["clocks"]["sm_mhz"])')
  echo "$1 warps=$2 pack=$3 prefetch=$4 -> $v" >> $out
}
for w in 8 10 12; do for p in 0 1; do for f in 0 1; do run f64 $w $p $f; done; done; done
for w in 16 20 24; do for p in 0 1; do for f in 0 1; do run f32 $w $p $f; done; done; done
cat $out
