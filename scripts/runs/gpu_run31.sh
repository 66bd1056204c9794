set -x
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r31_bench_c2.json 2> gpurun_out/r31_bench_c2.err; tail -n 3 gpurun_out/r31_bench_c2.err; cat gpurun_out/r31_bench_c2.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count_mode_one or kat_small or golden" 2>&1 | tail -n 8
