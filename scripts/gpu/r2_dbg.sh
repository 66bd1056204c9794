#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu/debug_gemm17.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -8
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_gemm.py 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; tail -c 300 gpurun_out/bench_c3_1gpu.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_c3_1gpu.json') if l.startswith('{')][-1])
print(j['ms_per_step'], j['e2e']['ms_per_step'], json.dumps(j['stages_ms']), j['roofline']['frac'], j['roofline']['kernel_ms'])
PY
