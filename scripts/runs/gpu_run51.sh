set -x
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 > gpurun_out/r51_bench_c5.json 2> gpurun_out/r51_bench_c5.err; tail -n 3 gpurun_out/r51_bench_c5.err; cat gpurun_out/r51_bench_c5.json
