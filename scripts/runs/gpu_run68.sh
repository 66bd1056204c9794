set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -n 3
timeout 600 python scripts/time_dims.py 300000 17 20 24 28 32 48 2>&1 | tail -n 8
