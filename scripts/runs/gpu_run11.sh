set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in C2 C4 C1 C3; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r11_prof.jsonl 2>&1; done
timeout 300 python scripts/profile_kernels.py C5 100000 1 >> gpurun_out/r11_prof.jsonl 2>&1
cat gpurun_out/r11_prof.jsonl
