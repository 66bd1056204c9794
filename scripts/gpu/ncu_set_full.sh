#!/bin/bash
# ncu --set full with source counters of both C3 scan kernels on a 400k-frame sample (39 replay passes each under source-level
# instrumentation: the full size takes minutes per kernel); read with ncu -i ... --page raw|source --csv, scripts/ncu_summary.py
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pops_bin|nn_kernel" -c 3 -f -o gpurun_out/prof_c3_400k python scripts/profile_kernels.py C3 400000 1 > gpurun_out/ncu_c3_400k.log 2>&1
tail -2 gpurun_out/ncu_c3_400k.log | cut -c1-300
