set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r20_bench_c2.json 2> gpurun_out/r20_bench_c2.err; tail -3 gpurun_out/r20_bench_c2.err; cat gpurun_out/r20_bench_c2.json
