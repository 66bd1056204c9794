set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_v2.json 2> gpurun_out/bench_c2_v2.err; tail -3 gpurun_out/bench_c2_v2.err; cat gpurun_out/bench_c2_v2.json
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_kernel' -s 2 -c 2 -o gpurun_out/prof_r1_c2_v2 -f python scripts/profile_kernels.py C2 262144 > gpurun_out/prof_c2_v2.log 2>&1
tail -2 gpurun_out/prof_c2_v2.log
