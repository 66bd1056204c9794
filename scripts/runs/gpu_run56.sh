set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r56_bench_c2_2gpu.json 2> gpurun_out/r56_c2.err; tail -n 2 gpurun_out/r56_c2.err; cat gpurun_out/r56_bench_c2_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload C5 > gpurun_out/r56_bench_c5_2gpu.json 2> gpurun_out/r56_c5.err; tail -n 2 gpurun_out/r56_c5.err; cat gpurun_out/r56_bench_c5_2gpu.json
