#!/bin/bash
# round 2, session 2, call 4: neighbour filter state parked in shared memory (no spills in nn_kernel<10>): parity, item-count sweep,
# ncu --set full of both C3 scan kernels on a 400k-frame sample (39 replay passes each: the full size takes minutes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/sweep_knobs.py C3 DCB200_ITEMS_PER_CTA=192 DCB200_ITEMS_PER_CTA=384 DCB200_ITEMS_PER_CTA=768 2>&1 | grep "^{" | tee -a gpurun_out/r2b_4_sweep.jsonl
timeout 300 python scripts/sweep_knobs.py C2 DCB200_ITEMS_PER_CTA=192 DCB200_ITEMS_PER_CTA=384 2>&1 | grep "^{" | tee -a gpurun_out/r2b_4_sweep.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pops_bin|nn_kernel" -c 3 -f -o gpurun_out/prof_r02_c3_400k python scripts/profile_kernels.py C3 400000 1 > gpurun_out/ncu_r02_c3.log 2>&1
tail -2 gpurun_out/ncu_r02_c3.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep | tail -2
