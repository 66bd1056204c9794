#!/bin/bash
# wide step shape of the neighbour kernel at D >= 9: parity, then A/B against the 4 x 4 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in "" _nnnarrow; do DCB200_LIB=$PWD/clustering_b200/libdcb200$v.so timeout 300 python scripts/sweep_knobs.py C3 2>&1 | grep "^{" | cut -c1-150 | tee -a gpurun_out/wide_ab.jsonl; done
