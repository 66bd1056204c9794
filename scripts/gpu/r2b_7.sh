#!/bin/bash
# round 2, session 2, call 7 (2 GPUs): multi-GPU tests (in-process NCCL path, host assembly, torchrun driver) and the C3 bench line at N = 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c3_2gpu.json 2> gpurun_out/bench_r02_c3_2gpu.err
tail -c 300 gpurun_out/bench_r02_c3_2gpu.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_r02_c3_2gpu.json') if l.startswith('{')][-1])
print("N=2", j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], json.dumps(j['stages_ms']), json.dumps(j.get('cxx_inprocess')), j['parity']['sharded_equals_unsharded'])
PY
