#!/bin/bash
# round-2 measurement pass for the 16384-point overlap-save kernel (C5) on one B200: GPU test suite, default bench line, its
# ncu launch list, per-step DRAM traffic of C5 (f64 / f32), one full ncu capture of the kernel (run under gpurun from the
# repo root; each ncu pass only after the same command has exited 0 without ncu)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -2 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_default.log 2> gpurun_out/r02_bench_default.err || exit 1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_short_ncu.log 2>&1
for dt in f64 f32; do
    python tools/profile_step.py c5 $dt > gpurun_out/r02_step_c5_$dt.log 2>&1 &&
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/r02_step_c5_$dt.csv python tools/profile_step.py c5 $dt > /dev/null 2>&1
done
python tools/fir_one.py 4095 f64 16384 > gpurun_out/r02_fir_one_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fir_fft16k_kernel --launch-skip 4 -c 1 -f \
    -o gpurun_out/r02_c5_f64_fft16k python tools/fir_one.py 4095 f64 16384 > gpurun_out/r02_c5_f64_fft16k_ncu.log 2>&1
python tools/fir_one.py 4095 f32 16384 >> gpurun_out/r02_fir_one_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fir_fft16k_kernel --launch-skip 4 -c 1 -f \
    -o gpurun_out/r02_c5_f32_fft16k python tools/fir_one.py 4095 f32 16384 > gpurun_out/r02_c5_f32_fft16k_ncu.log 2>&1
cat gpurun_out/r02_fir_one_c5.log
tail -c 300 gpurun_out/r02_bench_default.log
