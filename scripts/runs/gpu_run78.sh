set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_gemm.py -x -q -k "not matches_ffma" 2>&1 | tail -n 5
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "count_mode_one_and_two or kat_small or golden" 2>&1 | tail -n 5
