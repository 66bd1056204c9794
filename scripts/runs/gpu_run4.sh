python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_v3.json 2> gpurun_out/bench_c2_v3.err; tail -3 gpurun_out/bench_c2_v3.err; cat gpurun_out/bench_c2_v3.json
python scripts/profile_kernels.py C2 262144
python scripts/profile_kernels.py C3 131072
