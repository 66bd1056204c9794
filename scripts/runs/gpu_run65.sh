set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r65_bench_c2.json 2> gpurun_out/r65_c2.err; tail -n 2 gpurun_out/r65_c2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r65_bench_ref.json 2> gpurun_out/r65_ref.err
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 > gpurun_out/r65_bench_c5.json 2> gpurun_out/r65_c5.err; tail -n 2 gpurun_out/r65_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_count_kernel' -c 3 -o gpurun_out/prof_r01_c2 -f python scripts/profile_kernels.py C2 1000000 1 > gpurun_out/prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gscan' -c 3 -o gpurun_out/prof_r01_c5 -f python scripts/profile_kernels.py C5 500000 1 > gpurun_out/prof_c5.log 2>&1
cat gpurun_out/r65_bench_c2.json gpurun_out/r65_bench_ref.json gpurun_out/r65_bench_c5.json | cut -c1-260
