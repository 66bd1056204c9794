#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
run() { echo "== $1"; shift; env "$@" timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('pops', round(j['pops_ms'],2), 'eval', round(j['pops_eval_frac'],4), 'tflops', round(j['pops_exec_tflops'],2), '| nn', round(j['nn_ms'],2), 'eval', round(j['nn_eval_frac'],4), 'layout', round(j['layout_ms'],2))"; }
run "product (cols 2, stages 6)" X=1
run "cols 1" DCB200_LIB=clustering_b200/libdcb200_c1.so
run "cols 4" DCB200_LIB=clustering_b200/libdcb200_c4.so
run "stages 8" DCB200_LIB=clustering_b200/libdcb200_s8.so
run "order dims 10" DCB200_ORDER_DIMS=10
run "order dims 8" DCB200_ORDER_DIMS=8
run "order dims 4" DCB200_ORDER_DIMS=4
