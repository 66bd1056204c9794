set -x
DCB200_TRACE=1 timeout 900 python scripts/screening_timing.py C4 5000000 6 2>&1 | grep -v "^\[dcb200\] \(pop\|near\)" | tail -n 45
