set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r52_bench_c2.json 2> gpurun_out/r52_bench_c2.err; tail -n 3 gpurun_out/r52_bench_c2.err; cat gpurun_out/r52_bench_c2.json
