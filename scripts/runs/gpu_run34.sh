set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r34_bench_c2.json 2> gpurun_out/r34_bench_c2.err; tail -n 3 gpurun_out/r34_bench_c2.err; cat gpurun_out/r34_bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r34_bench_ref.json 2> gpurun_out/r34_bench_ref.err; cat gpurun_out/r34_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_count_kernel' -c 3 -o gpurun_out/prof_r01_c2 -f python scripts/profile_kernels.py C2 1000000 1 > gpurun_out/prof_c2.log 2>&1
tail -n 2 gpurun_out/prof_c2.log
for w in C1 C2 C3 C4; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r34_prof.jsonl 2>&1; done
cat gpurun_out/r34_prof.jsonl
