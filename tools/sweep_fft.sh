#!/bin/bash
# A/B sweep of the overlap-save FIR kernel's tuning knobs on C2 (run under gpurun); prints value per setting.
out=gpurun_out/sweep_fft.txt
: > $out
run() {  # dtype env...
  d=$1; shift
  v=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --dtype $d 2>&1 | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.readline()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("failed", e)')
  echo "$d $* -> $v" >> $out
}
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=8 LLZ_FFT_PACK=0 LLZ_FFT_STAGE=1
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=8 LLZ_FFT_PACK=1 LLZ_FFT_STAGE=0
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=10 LLZ_FFT_PACK=0 LLZ_FFT_STAGE=0
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=10 LLZ_FFT_PACK=1 LLZ_FFT_STAGE=0
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=12 LLZ_FFT_PACK=0 LLZ_FFT_STAGE=0
run f32 LLZ_FFT_F32X2=1 LLZ_FFT_WARPS=12 LLZ_FFT_PACK=1 LLZ_FFT_STAGE=0
run f32 LLZ_FFT_F32X2=0 LLZ_FFT_WARPS=16 LLZ_FFT_PACK=1 LLZ_FFT_STAGE=0
cat $out
