#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== C4 screening loop (edge list)"; DCB200_TRACE=1 timeout 900 python scripts/screening_timing.py C4 2>&1 | grep -v "upload+layout\|populations\|scan+download\|density_run: total" | tail -24
echo "== C3 ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pops_bin -c 1 -f -o gpurun_out/prof_bin_c3_v2 python scripts/profile_kernels.py C3 1000000 1 > gpurun_out/ncu_bin2.log 2>&1; tail -2 gpurun_out/ncu_bin2.log | cut -c1-200
