#!/bin/bash
# wide (2 rows x 8 columns) step shape of the table-driven kernel at D >= 9: parity, then A/B against the 4 x 4 shape and the 8-column-in-flight variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
for v in "" _narrow _w8; do DCB200_LIB=$PWD/clustering_b200/libdcb200$v.so timeout 300 python scripts/sweep_knobs.py C3 2>&1 | grep "^{" | cut -c1-150 | tee -a gpurun_out/wide_ab.jsonl; done
