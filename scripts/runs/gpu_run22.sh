set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in C1 C3 C2; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r22_prof.jsonl 2>&1; done
cat gpurun_out/r22_prof.jsonl
