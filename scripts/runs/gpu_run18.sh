set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in C2 C1 C3; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r18_prof.jsonl 2>&1; done
cat gpurun_out/r18_prof.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r18_bench_c2.json 2> gpurun_out/r18_bench_c2.err; tail -3 gpurun_out/r18_bench_c2.err; cat gpurun_out/r18_bench_c2.json
