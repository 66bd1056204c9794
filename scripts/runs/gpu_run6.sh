timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_v5.json 2> gpurun_out/bench_c2_v5.err; tail -3 gpurun_out/bench_c2_v5.err; cat gpurun_out/bench_c2_v5.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_kernel' -s 2 -c 2 -o gpurun_out/prof_r1_c2_v3 -f python scripts/profile_kernels.py C2 1000000 > gpurun_out/prof_c2_v3.log 2>&1
tail -2 gpurun_out/prof_c2_v3.log
