set -x
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 > gpurun_out/r73_bench_c5.json 2> gpurun_out/r73_c5.err; tail -n 2 gpurun_out/r73_c5.err
ncu --set full --clock-control none --import-source on -k regex:'gscan' -c 3 -o gpurun_out/prof_r01_c5 -f python scripts/profile_kernels.py C5 500000 1 > gpurun_out/prof_c5.log 2>&1
cat gpurun_out/r73_bench_c5.json | cut -c1-300
