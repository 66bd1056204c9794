#!/bin/bash
mkdir -p gpurun_out
make -C clustering_b200/csrc -j16 > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "cyclic or screening or density_run" 2>&1 | tail -4
echo "== C4 screening loop"; DCB200_TRACE=0 timeout 900 python scripts/screening_timing.py C4 2>&1 | tail -16
echo "== C4 ncu"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 40 -c 2 -f -o gpurun_out/prof_screen_c4 python scripts/screening_timing.py C4 5000000 60 > gpurun_out/ncu_screen.log 2>&1; tail -2 gpurun_out/ncu_screen.log | cut -c1-200
echo "== C4 launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"screen_kernel|uf_" -c 400 --csv --log-file gpurun_out/launches_c4_screen.csv python scripts/screening_timing.py C4 > /dev/null 2>&1; wc -l gpurun_out/launches_c4_screen.csv
echo "== bench C4"; timeout 900 python bench.py --workload C4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.err; head -c 600 gpurun_out/bench_c4.json
