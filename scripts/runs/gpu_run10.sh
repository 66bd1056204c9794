set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/profile_kernels.py C2 > gpurun_out/r10_prof.jsonl 2>&1
timeout 300 python scripts/profile_kernels.py C4 >> gpurun_out/r10_prof.jsonl 2>&1
cat gpurun_out/r10_prof.jsonl
DCB200_TRACE=1 timeout 300 python scripts/e2e_breakdown.py C2 2>&1 | tail -22
