#!/bin/bash
# round-2 measurement pass on one B200: the default bench line, its ncu launch list, per-step DRAM traffic of the
# resampler workloads, one full ncu capture of the tcgen05 kernel on config C4 (run under gpurun from the repo root)
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.log 2> gpurun_out/r02_bench_default.err || exit 1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_short_ncu.log 2>&1
for wl in "c4 f64" "c4 f32" "c3 f64" "c3 f32" "interp4 f64"; do
    set -- $wl
    python tools/profile_step.py $1 $2 > gpurun_out/r02_step_$1_$2.log 2>&1 &&
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/r02_step_$1_$2.csv python tools/profile_step.py $1 $2 > /dev/null 2>&1
done
python tools/profile_step.py c4 f64 > /dev/null 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:poly_bank_umma -c 1 \
    -o gpurun_out/r02_c4_exact_umma python tools/profile_step.py c4 f64 > gpurun_out/r02_c4_exact_umma_ncu.log 2>&1
python tools/profile_step.py c4 f32 > /dev/null 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:poly_bank_umma -c 1 \
    -o gpurun_out/r02_c4_fast_umma python tools/profile_step.py c4 f32 > gpurun_out/r02_c4_fast_umma_ncu.log 2>&1
ls -la gpurun_out | tail -20
