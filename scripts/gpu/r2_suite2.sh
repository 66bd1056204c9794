#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; tail -c 600 gpurun_out/bench_c3_1gpu.err; head -c 3000 gpurun_out/bench_c3_1gpu.json
