set -x
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_gemm.py -x -q -k "dims and (64 or 33 or 256)" 2>&1 | tail -n 6
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_gemm.py -x -q -k "dims and 40" 2>&1 | tail -n 6
