#!/bin/bash
# 8-GPU evidence: multi-GPU parity tests, bench.py under torchrun at N = 8 and N = 4, the `clustering` binary on 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -4
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err
  tail -c 400 gpurun_out/bench_c3_${N}gpu.err
  python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/bench_c3_${N}gpu.json') if l.startswith('{')][-1])
print("N=${N}", j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], json.dumps(j['stages_ms']), json.dumps(j.get('cxx_inprocess')), j['parity']['sharded_equals_unsharded'])
PY
done
# the command-line binary on all 8 GPUs: C3 as a multi-radius run (-R) and as a single-radius run with neighbours (-r -b)
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from clustering_b200.synth import config_data
x = config_data("C3")
np.savetxt("/tmp/c3.coords", x, fmt="%.6f")
PY
ls -la /tmp/c3.coords
R="0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 1.1 1.2 1.3 1.4 1.5 1.6 1.7 1.8 1.9 2.0"
for rep in 1 2; do
  /usr/bin/time -f "cli -R (20 radii, pops + fe files) wall %e s" env DCB200_TRACE=1 ./clustering_b200/clustering density -f /tmp/c3.coords -R $R -p /tmp/pop -d /tmp/fe 2>&1 | grep -E "total|wall" | tail -3
  /usr/bin/time -f "cli -r 1.0 -p -d -b wall %e s" env DCB200_TRACE=1 ./clustering_b200/clustering density -f /tmp/c3.coords -r 1.0 -p /tmp/pop1 -d /tmp/fe1 -b /tmp/nn1 2>&1 | grep -E "total|wall" | tail -3
done
CUDA_VISIBLE_DEVICES=0 /usr/bin/time -f "cli -R on ONE gpu wall %e s" env DCB200_TRACE=1 ./clustering_b200/clustering density -f /tmp/c3.coords -R $R -p /tmp/pop_1g -d /tmp/fe_1g 2>&1 | grep -E "total|wall" | tail -3
cmp /tmp/pop_2.000000 /tmp/pop_1g_2.000000 && cmp /tmp/fe_1.000000 /tmp/fe_1g_1.000000 && echo "8-GPU files == 1-GPU files"
