set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
DCB200_GEMM_CB=2 timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
DCB200_GEMM_RA=1 timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
DCB200_GEMM_RA=1 DCB200_GEMM_CB=1 timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
