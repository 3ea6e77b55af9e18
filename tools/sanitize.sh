#!/bin/bash
# compute-sanitizer over a small invocation of every kernel (tests/sanitize_small.py)
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  grep -E "Race reported|Invalid|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -10
done
