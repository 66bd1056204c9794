#!/bin/bash
# bench lines of the round (C3 default with both baselines, reference arm, C2 / C4 / C5), the ncu launch
# list of the default bench command and the speed-of-light sections of the two C3 scan kernels at full size
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -c 300 gpurun_out/bench_r02_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_c3_reference.json 2> gpurun_out/bench_r02_c3_reference.err
for w in C2 C4 C5; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; tail -c 200 gpurun_out/bench_r02_$w.err; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_r02.json 2> gpurun_out/bench_under_ncu_r02.err
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none -k regex:"pops_bin|nn_kernel" -c 3 -f -o gpurun_out/prof_r02_c3_full_sol python scripts/profile_kernels.py C3 1000000 1 > gpurun_out/ncu_r02_c3_full.log 2>&1
tail -2 gpurun_out/ncu_r02_c3_full.log | cut -c1-200
python - <<'PY'
import json
for f in ["bench_r02_c3", "bench_r02_c3_reference", "bench_r02_C2", "bench_r02_C4", "bench_r02_C5"]:
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms", round(j["ms_per_step"], 2), "e2e", round(j["e2e"].get("ms_per_step", 0), 2), "value", round(j["value"], 1),
              "frac", round(j.get("roofline", {}).get("frac", 0), 3), j.get("roofline", {}).get("kernel"), json.dumps(j.get("stages_ms"))[:300],
              json.dumps(j.get("cpu_baseline"))[:200], json.dumps(j.get("reference_cuda"))[:300], json.dumps(j.get("screening")))
    except Exception as ex:
        print(f, "failed", ex)
PY
