timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/profile_kernels.py C2 262144
timeout 300 python scripts/profile_kernels.py C3 131072
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_v4.json 2> gpurun_out/bench_c2_v4.err; tail -3 gpurun_out/bench_c2_v4.err; cat gpurun_out/bench_c2_v4.json
