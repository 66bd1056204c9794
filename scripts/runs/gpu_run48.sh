ncu --set full --clock-control none --import-source on -k regex:'gscan_pops' -c 1 -o gpurun_out/prof_gpops3 -f python scripts/profile_kernels.py C5 200000 1 > gpurun_out/prof_gpops3.log 2>&1
tail -n 2 gpurun_out/prof_gpops3.log
