set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_kernel' -s 2 -c 2 -o gpurun_out/prof_r1_c2 -f python scripts/profile_kernels.py C2 262144 > gpurun_out/prof_c2.log 2>&1
tail -3 gpurun_out/prof_c2.log
ncu --set full --clock-control none --import-source on -k regex:'pops_kernel' -s 1 -c 1 -o gpurun_out/prof_r1_c3 -f python scripts/profile_kernels.py C3 131072 > gpurun_out/prof_c3.log 2>&1
tail -3 gpurun_out/prof_c3.log
