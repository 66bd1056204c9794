#!/bin/bash
# compute-sanitizer (memcheck, synccheck, racecheck) over a small pass through every FFMA scan kernel
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 420 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
