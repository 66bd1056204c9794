set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in C2 C4 C1; do timeout 300 python scripts/profile_kernels.py $w >> gpurun_out/r19_prof.jsonl 2>&1; done
cat gpurun_out/r19_prof.jsonl
