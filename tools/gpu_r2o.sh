#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_smoke.py 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2o_sanitizer_$tool.log 2>&1; echo "$tool rc $?"; tail -6 gpurun_out/r2o_sanitizer_$tool.log | cut -c1-200
done
