set -x
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
DCB200_TRACE=1 timeout 900 python scripts/screening_timing.py C4 5000000 8 2>&1 | grep -v "^\[dcb200\] \(pop\|near\)" | tail -n 40
