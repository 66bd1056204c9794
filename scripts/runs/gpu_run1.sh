set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -5 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
nproc
