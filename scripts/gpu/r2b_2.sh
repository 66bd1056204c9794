#!/bin/bash
# round 2, session 2, call 2: named-barrier stage hand-back + FFMA2 inner product (product build) against the builds
# without FFMA2 and with an 8-stage ring; parity first
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
line() { python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$1', j['workload'], 'pops', round(j['pops_ms'],2), 'tflops', round(j['pops_exec_tflops'],2), '| nn', round(j['nn_ms'],2), 'tflops', round(j['nn_exec_tflops'],2), '| exact', j['pops_exact'], j['nn_exact'])"; }
for v in "" _noffma2 _s8; do
  for w in C3 C2; do
    DCB200_LIB=$PWD/clustering_b200/libdcb200$v.so timeout 300 python scripts/profile_kernels.py $w 2>&1 | tail -n 1 | tee -a gpurun_out/r2b_2_scans.jsonl | line "lib$v"
  done
done
