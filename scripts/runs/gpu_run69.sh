set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r69_bench_c2_4gpu.json 2> gpurun_out/r69_c2.err; tail -n 2 gpurun_out/r69_c2.err; cat gpurun_out/r69_bench_c2_4gpu.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
