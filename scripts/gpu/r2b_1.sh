#!/bin/bash
# round 2, session 2, call 1: FFMA2 micro-benchmark, full GPU suite after the register diet, C3 scans, shards of 8 on one GPU
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 scripts/micro/ffma2.cu && /tmp/ffma2 2>&1 | tee gpurun_out/ffma2.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python scripts/profile_kernels.py C3 2>&1 | tail -n 1 | tee gpurun_out/r2b_c3_scans.json | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('C3 pops', round(j['pops_ms'],2), 'eval', round(j['pops_eval_frac'],4), 'tflops', round(j['pops_exec_tflops'],2), '| nn', round(j['nn_ms'],2), 'eval', round(j['nn_eval_frac'],4), 'tflops', round(j['nn_exec_tflops'],2))"
echo "== shards of 8, adaptive items"; timeout 300 python scripts/shard_timing.py C3 8 0 3 2>&1 | tail -n 1 | tee gpurun_out/r2b_shards8.json
echo "== shards of 8, 16 column items"; DCB200_ITEMS_PER_CTA=1 timeout 300 python scripts/shard_timing.py C3 8 0 3 2>&1 | tail -n 1 | tee gpurun_out/r2b_shards8_old.json
