timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s -k "active_and_bounded" 2>&1 | tail -n 4
