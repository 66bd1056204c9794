set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
timeout 300 python scripts/profile_kernels.py C5 100000 2 2>&1 | tail -n 2
DCB200_GEMM=0 timeout 300 python scripts/profile_kernels.py C5 100000 2 2>&1 | tail -n 1
timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 2
ncu --set full --clock-control none --import-source on -k regex:'gscan' -c 3 -o gpurun_out/prof_r01_c5 -f python scripts/profile_kernels.py C5 500000 1 > gpurun_out/prof_c5.log 2>&1
tail -n 2 gpurun_out/prof_c5.log
