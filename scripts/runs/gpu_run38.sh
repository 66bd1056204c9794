set -x
DCB200_GEMM_PROF=1 timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 16 | cut -c1-200
