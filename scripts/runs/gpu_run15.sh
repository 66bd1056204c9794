set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in C2 C4 C1 C3; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r15_prof.jsonl 2>&1; done
cat gpurun_out/r15_prof.jsonl
ncu --set full --clock-control none --import-source on -k regex:'nn_kernel|pops_count_kernel' -c 3 -o gpurun_out/prof_r15_c2 -f python scripts/profile_kernels.py C2 1000000 1 > gpurun_out/prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pops_count_kernel' -c 1 -o gpurun_out/prof_r15_c4 -f python scripts/profile_kernels.py C4 5000000 1 > gpurun_out/prof_c4.log 2>&1
tail -2 gpurun_out/prof_c2.log gpurun_out/prof_c4.log
