#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
run() { echo "== $1 $2"; w=$2; shift; shift; env "$@" timeout 300 python scripts/profile_kernels.py $w 2>&1 | tail -n 1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('pops', round(j['pops_ms'],2), 'eval', round(j['pops_eval_frac'],4), 'tflops', round(j['pops_exec_tflops'],2), '| nn', round(j['nn_ms'],2), 'eval', round(j['nn_eval_frac'],4), 'tflops', round(j['nn_exec_tflops'],2), 'exact', j['nn_exact'])"; }
for w in C3 C2 C4 C1; do
run "axis on " $w X=1
run "axis off" $w DCB200_AXIS_PRUNE=2
done
