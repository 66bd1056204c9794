#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; tail -c 1500 gpurun_out/bench_c3_2gpu.err; head -c 1200 gpurun_out/bench_c3_2gpu.json
