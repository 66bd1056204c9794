set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 > gpurun_out/r74_bench_c5.json 2> gpurun_out/r74_c5.err; tail -n 2 gpurun_out/r74_c5.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r74_bench_c2.json 2> gpurun_out/r74_c2.err; tail -n 2 gpurun_out/r74_c2.err
cat gpurun_out/r74_bench_c5.json gpurun_out/r74_bench_c2.json | cut -c1-200
