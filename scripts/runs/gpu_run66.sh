set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k full_size 2>&1 | tail -n 8
