set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_count_kernel' -c 3 -o gpurun_out/prof_r01_c2 -f python scripts/profile_kernels.py C2 1000000 1 > gpurun_out/prof_c2.log 2>&1
tail -n 2 gpurun_out/prof_c2.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r23_bench_c2.json 2> gpurun_out/r23_bench_c2.err; tail -n 3 gpurun_out/r23_bench_c2.err; cat gpurun_out/r23_bench_c2.json
