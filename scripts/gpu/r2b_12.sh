#!/bin/bash
# round 2, session 2, call 12: cell-table address as one LOP3 (size-aligned table): parity incl. the new knob test, C3 timing, racecheck with all hazards listed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/sweep_knobs.py C3 2>&1 | grep "^{" | tee -a gpurun_out/r2b_12_scans.jsonl
timeout 300 python scripts/sweep_knobs.py C1 2>&1 | grep "^{" | tee -a gpurun_out/r2b_12_scans.jsonl
timeout 420 compute-sanitizer --tool racecheck --print-limit 100 python scripts/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1
grep -E "^ok|RACECHECK SUMMARY" gpurun_out/sanitize_racecheck.log; grep -oE "in [a-z_]+\.cuh?:[0-9]+" gpurun_out/sanitize_racecheck.log | sort | uniq -c
