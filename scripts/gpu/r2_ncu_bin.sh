#!/bin/bash
# ncu capture of the multi-radius population kernel on C3 (1M x 10, 20 radii) + row-interleave A/B
mkdir -p gpurun_out
echo "== interleave"; DCB200_BIN_INTERLEAVE=1 timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1 | cut -c1-400
echo "== coherent"; timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1 | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pops_bin -c 1 -f -o gpurun_out/prof_bin_c3 python scripts/profile_kernels.py C3 1000000 1 > gpurun_out/ncu_bin.log 2>&1
tail -3 gpurun_out/ncu_bin.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
