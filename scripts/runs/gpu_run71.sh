set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-600
timeout 600 python scripts/time_dims.py 300000 17 24 32 48 2>&1 | tail -n 4
