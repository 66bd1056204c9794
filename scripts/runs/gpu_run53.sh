set -x
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -n 3
ncu --set full --clock-control none --import-source on -k regex:'gscan' -c 3 -o gpurun_out/prof_r01_c5 -f python scripts/profile_kernels.py C5 500000 1 > gpurun_out/prof_c5.log 2>&1
tail -n 2 gpurun_out/prof_c5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_c5.csv python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_under_ncu.json 2> gpurun_out/bench_c5_under_ncu.err
