set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r76_bench_c2_8gpu.json 2> gpurun_out/r76_c2.err; tail -n 2 gpurun_out/r76_c2.err; cat gpurun_out/r76_bench_c2_8gpu.json | cut -c1-300
