#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py (every deposit layout, the fused run, the LB kernels); logs -> gpurun_out/
O=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > $O/r02c_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; tail -4 $O/r02c_sanitizer_$tool.log
done
