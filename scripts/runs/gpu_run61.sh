set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k edge 2>&1 | tail -n 12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
