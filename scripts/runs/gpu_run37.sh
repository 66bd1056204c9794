set -x
timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 1 | cut -c1-330
DCB200_GEMM_SPIN=1 timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 1 | cut -c1-330
