#!/bin/bash
# round 2, session 2, call 5: two-level neighbour filter: full GPU suite, then scan times of every FFMA workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in C3 C2 C4 C1; do timeout 300 python scripts/sweep_knobs.py $w 2>&1 | grep "^{" | tee -a gpurun_out/r2b_5_scans.jsonl; done
