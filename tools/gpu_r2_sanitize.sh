#!/bin/bash
# compute-sanitizer over every kernel family on small shapes (memcheck, then racecheck and synccheck)
python tools/sanitize_run.py > gpurun_out/r2_sanitize_plain.log 2>&1 || { tail -20 gpurun_out/r2_sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitize_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok" gpurun_out/r2_sanitize_$tool.log >> gpurun_out/r2_sanitize_summary.log
done
cat gpurun_out/r2_sanitize_summary.log
