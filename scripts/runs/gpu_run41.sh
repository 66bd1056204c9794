set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 5
DCB200_GEMM_RA=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
DCB200_GEMM_PROF=1 timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 14 | cut -c1-400
timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
