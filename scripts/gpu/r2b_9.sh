#!/bin/bash
# round 2, session 2, call 9 (8 GPUs): the C3 bench line at N = 8 with the final kernels
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c3_8gpu.json 2> gpurun_out/bench_r02_c3_8gpu.err
tail -c 400 gpurun_out/bench_r02_c3_8gpu.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_r02_c3_8gpu.json') if l.startswith('{')][-1])
print("N=8", j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], json.dumps(j['stages_ms']), json.dumps(j.get('cxx_inprocess')), j['parity']['sharded_equals_unsharded'])
print([ (round(g['pops_ms'],1), round(g['nn_ms'],1)) for g in j['roofline']['per_gpu']])
PY
