set -x
mkdir -p gpurun_out/bench
python bench.py --workload c1 > gpurun_out/bench/bench_c1.json 2> gpurun_out/bench/c1.err
for w in c2 c3 c4 c5; do for dt in f64 f32; do
  python bench.py --workload $w --dtype $dt > gpurun_out/bench/bench_${w}_${dt}.json 2> gpurun_out/bench/${w}_${dt}.err
done; done
python bench.py --impl reference > gpurun_out/bench/bench_ref_c2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_c2_f64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fir_fft8k_kernel -s 2 -c 1 -f -o gpurun_out/prof_c5_f64 python bench.py --workload c5 --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:poly_slide_kernel -s 2 -c 1 -f -o gpurun_out/prof_c3_f64 python bench.py --workload c3 --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:poly_bank_dmma -s 2 -c 1 -f -o gpurun_out/prof_c4_f64 python bench.py --workload c4 --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
tail -c 600 gpurun_out/bench/bench_c5_f64.json
