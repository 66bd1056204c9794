#!/bin/bash
# round 2, session 2, call 3: unit takeover in the table-driven population kernel: parity, then A/B against the previous build and knob sweeps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/sweep_knobs.py C3 DCB200_BIN_STEAL=0 DCB200_BIN_DENSE_LANES=4 DCB200_BIN_DENSE_LANES=12 DCB200_BIN_DENSE_LANES=16 DCB200_ITEMS_PER_CTA=96 DCB200_ITEMS_PER_CTA=192 DCB200_BIN_PROJ=0 2>&1 | grep "^{" | tee -a gpurun_out/r2b_3_sweep.jsonl
DCB200_LIB=$PWD/clustering_b200/libdcb200_base.so timeout 300 python scripts/sweep_knobs.py C3 2>&1 | grep "^{" | tee -a gpurun_out/r2b_3_sweep.jsonl
