set -x
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r8_bench_c2.json 2> gpurun_out/r8_bench_c2.err; tail -3 gpurun_out/r8_bench_c2.err; cat gpurun_out/r8_bench_c2.json
timeout 300 python scripts/profile_kernels.py C1 > gpurun_out/r8_prof.jsonl 2>&1
timeout 300 python scripts/profile_kernels.py C3 >> gpurun_out/r8_prof.jsonl 2>&1
timeout 400 python scripts/profile_kernels.py C4 >> gpurun_out/r8_prof.jsonl 2>&1
timeout 400 python scripts/profile_kernels.py C5 100000 1 >> gpurun_out/r8_prof.jsonl 2>&1
cat gpurun_out/r8_prof.jsonl
