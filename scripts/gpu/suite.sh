#!/bin/bash
# the GPU test suite, then the scan times of every FFMA workload (best of three launches each)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in C3 C2 C4 C1; do timeout 300 python scripts/sweep_knobs.py $w 2>&1 | grep "^{" | tee -a gpurun_out/scan_times.jsonl; done
