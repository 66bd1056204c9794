nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -5 gpurun_out/bench_c2_n2.err; cat gpurun_out/bench_c2_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; tail -3 gpurun_out/bench_c2_n1.err; cat gpurun_out/bench_c2_n1.json
