set -x
timeout 120 python scripts/gemm_smoke.py 2>&1 | tail -n 40
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s 2>&1 | tail -n 25
