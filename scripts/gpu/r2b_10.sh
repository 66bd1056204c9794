#!/bin/bash
# round 2, session 2, call 10: coarse (super-tile) level of the producers' pruning: full GPU suite, then A/B on every FFMA workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in C4 C3 C2 C1; do timeout 300 python scripts/sweep_knobs.py $w DCB200_SUPER_PRUNE=0 2>&1 | grep "^{" | tee -a gpurun_out/r2b_10_scans.jsonl; done
