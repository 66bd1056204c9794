#!/bin/bash
# full GPU parity suite + C3 timing with / without the separating-axis pruning
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== C3 proj on"; timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1 | cut -c1-420
echo "== C3 proj off"; DCB200_BIN_PROJ=2 timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1 | cut -c1-420
echo "== C2"; timeout 300 python scripts/profile_kernels.py C2 1000000 2 2>&1 | tail -n 1 | cut -c1-420
echo "== C1"; timeout 300 python scripts/profile_kernels.py C1 100000 3 2>&1 | tail -n 1 | cut -c1-420
