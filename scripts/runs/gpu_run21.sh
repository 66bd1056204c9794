set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r21_bench_n2.json 2> gpurun_out/r21_bench_n2.err; tail -5 gpurun_out/r21_bench_n2.err; cat gpurun_out/r21_bench_n2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r21_bench_ref.json 2> gpurun_out/r21_bench_ref.err; tail -3 gpurun_out/r21_bench_ref.err; cat gpurun_out/r21_bench_ref.json
