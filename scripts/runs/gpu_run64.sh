set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -n 3
for w in C1 C2 C3 C4; do timeout 300 python scripts/profile_kernels.py $w 2>&1 | tail -n 1 | cut -c1-420; done
