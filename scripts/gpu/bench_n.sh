#!/bin/bash
# the C3 bench line under torchrun at N GPUs (gpurun --gpus N -- bash scripts/gpu/bench_n.sh N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err
tail -c 400 gpurun_out/bench_c3_${N}gpu.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/bench_c3_${N}gpu.json') if l.startswith('{')][-1])
print("N=${N}", j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], json.dumps(j['stages_ms']), json.dumps(j.get('cxx_inprocess')), j['parity']['sharded_equals_unsharded'])
print([(round(g['pops_ms'], 1), round(g['nn_ms'], 1)) for g in j['roofline']['per_gpu']])
PY
