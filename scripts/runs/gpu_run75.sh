set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
DCB200_GEMM_RA=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
